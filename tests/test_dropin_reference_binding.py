"""INTEGRATION.md section 2, executed: the reference's OWN binding files -- ``network/ops_dcnv3/functions/dcnv3_func.py``
(``import DCNv3``, ``pkg_resources.get_distribution('DCNv3')``, ``DCNv3Function`` :16-19,22-106) and
``network/ops_dcnv3/modules/dcnv3.py`` (the ``DCNv3`` module :221-356) -- imported UNCHANGED on top of
``givepose_b200.dropin`` (our ``DCNv3`` stub + ``DCNv3-1.1.dist-info``).

The files come from ``baseline/_ref/GIVEPose`` (staged by ``__graft_entry__.build()``, git-ignored, travels to the GPU box) or
from ``/root/reference`` where it exists.  CPU part: the call reaches our stub (a CPU tensor is refused exactly like
``src/dcnv3.h:37``).  GPU part: the reference module's forward + backward on our kernels against the same module on the
reference's own CUDA extension (``oracle/_ref/DCNv3_ref.so``)."""
import importlib
import os
import sys

import pytest
import torch

from baseline import reference as R


@pytest.fixture(scope="module")
def ref_pkg():
    root = R.root()
    if root is None:
        pytest.skip("reference files neither staged (baseline/_ref) nor present (/root/reference)")
    from givepose_b200 import dropin
    here = dropin.install()
    for name in [m for m in sys.modules if m == "DCNv3" or m.startswith("network.ops_dcnv3")]:
        del sys.modules[name]                      # nothing cached from another test's stub
    import warnings
    assert os.path.isdir(os.path.join(here, "DCNv3-1.1.dist-info"))   # what pkg_resources.get_distribution('DCNv3') finds (:17-19)
    if root not in sys.path:
        sys.path.insert(0, root)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        func = importlib.import_module("network.ops_dcnv3.functions.dcnv3_func")
        mods = importlib.import_module("network.ops_dcnv3.modules.dcnv3")
    assert os.path.realpath(func.__file__).startswith(os.path.realpath(root))
    return func, mods


def test_reference_dcnv3_func_binds_to_the_dropin_stub(ref_pkg):
    func, mods = ref_pkg
    import givepose_b200.functions as ours
    assert func.DCNv3.__name__ == "DCNv3" and "givepose_b200" in os.path.realpath(func.DCNv3.__file__)
    assert func.DCNv3.dcnv3_forward is ours.dcnv3_forward and func.DCNv3.dcnv3_backward is ours.dcnv3_backward
    assert func.dcn_version == 1.1                                     # read from DCNv3-1.1.dist-info by pkg_resources
    x, off, m = torch.zeros(2, 8, 8, 64), torch.zeros(2, 8, 8, 72), torch.zeros(2, 8, 8, 36)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):   # the reference Function reached OUR entry point
        func.DCNv3Function.apply(x, off, m, 3, 3, 1, 1, 1, 1, 1, 1, 4, 16, 1.0, 256, False)
    # the reference nn.Module on top of it (modules/dcnv3.py:318-345)
    layer = mods.DCNv3(channels=64, kernel_size=3, stride=1, pad=1, dilation=1, group=4, offset_scale=1.0)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        layer(torch.zeros(2, 8, 8, 64))


@pytest.mark.gpu
@pytest.mark.parametrize("stride", [1, 2])
def test_reference_module_on_our_kernels_equals_reference_module_on_its_own_extension(ref_pkg, stride):
    func, mods = ref_pkg
    from oracle import build_ref_ext
    if not os.path.exists(build_ref_ext.OUT):
        pytest.skip("oracle/_ref/DCNv3_ref.so not built")
    ref_ext = build_ref_ext.load()
    torch.manual_seed(0)
    layer = mods.DCNv3(channels=256, kernel_size=3, stride=stride, pad=1, dilation=1, group=4, offset_scale=1.0).cuda()
    with torch.no_grad():   # DCNv3._reset_parameters zeroes offset / mask (:308-316): give the sampler something to do
        for lin in (layer.offset, layer.mask):
            lin.weight.normal_(std=0.05)
            lin.bias.normal_(std=0.5)
    x = torch.randn(8, 32, 32, 256, device="cuda")
    results = []
    ours_mod = func.DCNv3
    try:
        for ext in (ours_mod, ref_ext):
            func.DCNv3 = ext                       # what `import DCNv3` resolved to inside the reference file
            xi = x.clone().requires_grad_(True)
            layer.zero_grad()
            out = layer(xi)
            out.square().mean().backward()
            results.append((out.detach(), xi.grad.clone(), layer.offset.weight.grad.clone(), layer.mask.weight.grad.clone(),
                            layer.input_proj.weight.grad.clone()))
    finally:
        func.DCNv3 = ours_mod
    rel = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()
    assert results[0][0].shape == (8, 32 // stride, 32 // stride, 256)
    assert rel(results[0][0], results[1][0]) < 1e-5
    for a, b in zip(results[0][1:], results[1][1:]):
        assert rel(a, b) < 1e-4
