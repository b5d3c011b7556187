"""Stage + import the reference's own files (``/root/reference`` exists only in the build container).

* ``stage()`` -- called by ``__graft_entry__.build()``: copies the ``*.py`` files of the reference's ``network/``, ``config/``,
  ``losses/`` and ``tools/`` trees, unmodified, into ``baseline/_ref/GIVEPose/`` (git-ignored, travels with gpurun) and writes a
  SHA-256 manifest.  A no-op where ``/root/reference`` is absent (the GPU box uses what was staged).
* ``load_dcnv3_func(dcnv3_module=None)`` -- imports the reference's ``network/ops_dcnv3/functions/dcnv3_func.py``
  (``DCNv3Function`` :22-106, ``dcnv3_core_pytorch`` :172-220).  The file does ``import DCNv3`` and reads
  ``pkg_resources.get_distribution('DCNv3').version`` at module scope (:16-19): ``dcnv3_module`` is what that import resolves to
  (``None`` = an empty placeholder, enough for the pure-PyTorch ``dcnv3_core_pytorch``; pass
  ``givepose_b200.dropin`` 's stub to run the reference's own autograd Function on our kernels).
* ``load_posenet(dcnv3_forward, dcnv3_backward=None)`` -- imports the reference's ``network/PoseNet.py`` unchanged behind leaf
  stubs for the third-party packages this image lacks (``baseline/stubs.py``) and returns the module.
* ``reference_posenet_cpu()`` -- the reference ``PoseNet`` (eval, CPU) the bench's reference arm times: compiled extension replaced
  by the reference's own ``dcnv3_core_pytorch`` behind the flat-slice adapter, ``convnext_backbone`` (timm +
  pretrained download) replaced by the reference's own ``network/resnet.py`` ResNet-34 trunk + 1x1 neck, like the golden vectors.
"""
from __future__ import annotations

import hashlib
import importlib
import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref", "GIVEPose")
TREES = ("network", "config", "losses", "tools")


def stage(force: bool = False) -> str | None:
    """Copy the reference's Python files into the git-ignored staging area.  Returns the staged root or None."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    lines = []
    for tree in TREES:
        for d, _, files in os.walk(os.path.join(SRC, tree)):
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                src = os.path.join(d, f)
                dst = os.path.join(DST, os.path.relpath(src, SRC))
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                    shutil.copy2(src, dst)
                lines.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {os.path.relpath(dst, DST)}")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return DST


def root() -> str | None:
    """Where the reference's files can be imported from: the staged copy, else the original tree, else None."""
    if os.path.isfile(os.path.join(DST, "network", "PoseNet.py")):
        return DST
    if os.path.isfile(os.path.join(SRC, "network", "PoseNet.py")):
        return SRC
    return None


def _fake_dcnv3_distribution():
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pkg_resources
    if getattr(pkg_resources.get_distribution, "_gp_patched", False):
        return
    real = pkg_resources.get_distribution

    def fake(name):
        if name == "DCNv3":
            return types.SimpleNamespace(version="1.1")   # network/ops_dcnv3/setup.py:63-64
        return real(name)

    fake._gp_patched = True
    pkg_resources.get_distribution = fake


def load_dcnv3_func(dcnv3_module=None):
    r = root()
    if r is None:
        raise FileNotFoundError("reference files are neither staged under baseline/_ref/ nor present at /root/reference")
    if dcnv3_module is not None:
        sys.modules["DCNv3"] = dcnv3_module
    else:
        sys.modules.setdefault("DCNv3", types.ModuleType("DCNv3"))
        _fake_dcnv3_distribution()
    import warnings
    path = os.path.join(r, "network", "ops_dcnv3", "functions", "dcnv3_func.py")
    spec = importlib.util.spec_from_file_location("gp_ref_dcnv3_func", path)
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


def load_posenet(dcnv3_forward, dcnv3_backward=None):
    from . import stubs
    r = root()
    if r is None:
        raise FileNotFoundError("reference files are neither staged under baseline/_ref/ nor present at /root/reference")
    stubs.install(dcnv3_forward, dcnv3_backward)
    if r not in sys.path:
        sys.path.insert(0, r)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import config.config  # noqa: F401
        import absl.flags as flags
        if not flags.FLAGS.is_parsed():
            flags.FLAGS(["x"])
        import network.PoseNet as PN
    return PN


def reference_posenet_cpu():
    """(module, PoseNet instance) of the reference on the CPU, DCNv3 core = its own dcnv3_core_pytorch + flat-slice adapter."""
    import torch.nn as nn
    from oracle.dcnv3 import flat_slice, out_size   # shape arithmetic of the adapter only (test infrastructure)
    core = {}

    def dcnv3_forward_stub(input, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, group, gc, scale, im2col_step, rc=0):
        N, H, W, _ = input.shape
        Ho, Wo = out_size(H, kh, sh, ph, dh), out_size(W, kw, sw, pw, dw)
        return core["fn"](input, flat_slice(offset, N, Ho, Wo), flat_slice(mask, N, Ho, Wo), kh, kw, sh, sw, ph, pw, dh, dw,
                          group, gc, scale, rc)

    PN = load_posenet(dcnv3_forward_stub)
    from network.ops_dcnv3.functions import dcnv3_func
    core["fn"] = dcnv3_func.dcnv3_core_pytorch
    from network.resnet import resnet34

    class RefBackbone(nn.Module):   # reference ResNet-34 trunk (network/resnet.py) + neck, emitting [B,1024,8,8]
        def __init__(self):
            super().__init__()
            r = resnet34()
            del r.fc, r.avgpool
            self.trunk, self.neck = r, nn.Conv2d(512, 1024, 1)

        def forward(self, x):
            t = self.trunk
            x = t.maxpool(t.relu(t.bn1(t.conv1(x))))
            return [self.neck(t.layer4(t.layer3(t.layer2(t.layer1(x)))))]

    PN.convnext_backbone = lambda: RefBackbone()
    return PN, PN.PoseNet().eval()


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
