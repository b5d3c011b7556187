"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the host mirror raises the reference's errors, and the `DCNv3` drop-in module resolves.
No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from givepose_b200 import _lib
    header = open(os.path.join(ROOT, "include", "givepose_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/givepose_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    assert _lib.lib.gp_abi_version() == 1
    # dcnv3_cuda.cu:40-45
    assert _lib.lib.gp_dcnv3_out_size(64, 3, 2, 1, 1) == 32 and _lib.lib.gp_dcnv3_out_size(64, 3, 1, 1, 1) == 64
    assert b"remove_center" in _lib.lib.gp_error_string(-6)


def test_desc_struct_layout_matches_header():
    from givepose_b200._lib import DCNv3Desc
    assert ctypes.sizeof(DCNv3Desc) == 17 * 4


def test_cpu_tensors_are_refused_like_the_reference():
    """src/dcnv3.h:37 -> AT_ERROR("Not implemented on the CPU"); there is no CPU fallback here either."""
    import givepose_b200.functions as F
    x = torch.zeros(2, 8, 8, 64)
    off = torch.zeros(2, 8, 8, 72)
    m = torch.zeros(2, 8, 8, 36)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        F.dcnv3_forward(x, off, m, 3, 3, 1, 1, 1, 1, 1, 1, 4, 16, 1.0, 256, 0)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        F.dcnv3_backward(x, off, m, 3, 3, 1, 1, 1, 1, 1, 1, 4, 16, 1.0, x, 256, 0)
    with pytest.raises(RuntimeError):
        F.DCNv3Function.apply(x, off, m, 3, 3, 1, 1, 1, 1, 1, 1, 4, 16, 1.0, 256, 0)


def test_dropin_module_and_distribution():
    """functions/dcnv3_func.py:16-19 needs `import DCNv3` and a DCNv3 distribution whose version is > 1.0."""
    from givepose_b200 import dropin
    dropin.install()
    import importlib
    mod = importlib.import_module("DCNv3")
    assert callable(mod.dcnv3_forward) and callable(mod.dcnv3_backward)
    import importlib.metadata as md
    assert float(md.version("DCNv3")) > 1.0


def test_product_never_imports_oracle():
    """The oracle and the reference-arm plumbing are test infrastructure: nothing under givepose_b200/ may reference them."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "givepose_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|baseline)\b", src, flags=re.M), os.path.join(dirpath, f)
