"""GPU parity: our sm_100a kernels (through the C ABI) against the REFERENCE'S OWN CUDA kernels on the same B200.

``oracle/_ref/DCNv3_ref.so`` is the reference extension (network/ops_dcnv3/src: vision.cpp, cpu/dcnv3_cpu.cpp,
cuda/dcnv3_cuda.cu + dcnv3_im2col_cuda.cuh) compiled unmodified from where it lies by ``oracle/build_ref_ext.py`` in
the build container; the prebuilt module travels to the GPU box.  It pins what no reference test pins:
the stride-2 flat-offset addressing (cuh:229,243-244), half storage, the golden cases run through the real kernels,
``im2col_step`` chunking (dcnv3_cuda.cu:59-83) and the batch-coupling quirk.

Bars: fp32 forward 1e-5 max-norm relative (same arithmetic, different summation order), grads 1e-4 (both sides
accumulate with atomics in undefined order); half forward 2e-3 / grads 5e-3 (both round to half once at the end; the
reference's half atomics path accumulates grad_offset/grad_mask in fp32 like ours, dcnv3_cuda.cu:126-173).
"""
import os

import pytest
import torch

from golden_util import CASES, CASE_IDS

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    from oracle import build_ref_ext
    if not os.path.exists(build_ref_ext.OUT):
        pytest.skip("oracle/_ref/DCNv3_ref.so was not built (needs /root/reference at build time)")
    return build_ref_ext.load()


@pytest.fixture(scope="module")
def ops():
    import givepose_b200.functions as F
    return F


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _inputs(N, H, W, G, gc, P, Ho, Wo, dist, dtype, seed=3, full_res=False):
    """dist T = ops_dcnv3/test.py:36-40, dist M = model-like.  full_res: offset/mask at input resolution
    (what modules/dcnv3.py:330-334 hands the stride-2 kernel)."""
    g = torch.Generator().manual_seed(seed)
    oh, ow = (H, W) if full_res else (Ho, Wo)
    if dist == "T":
        inp = torch.rand(N, H, W, G * gc, generator=g) * 0.01
        off = torch.rand(N, oh, ow, G * P * 2, generator=g) * 10
        m = torch.rand(N, oh, ow, G, P, generator=g) + 1e-5
        m = (m / m.sum(-1, keepdim=True)).reshape(N, oh, ow, G * P)
    else:
        inp = torch.randn(N, H, W, G * gc, generator=g)
        off = torch.randn(N, oh, ow, G * P * 2, generator=g)
        m = torch.softmax(torch.randn(N, oh, ow, G, P, generator=g), -1).reshape(N, oh, ow, G * P)
    gout = torch.randn(N, Ho, Wo, G * gc, generator=g)
    return [t.to("cuda", dtype).contiguous() for t in (inp, off, m, gout)]


def _both(ops, ref, inp, off, m, gout, args, step=256, rc=0):
    out = ops.dcnv3_forward(inp, off, m, *args, step, rc)
    rout = ref.dcnv3_forward(inp, off, m, *args, step, rc)
    grads = ops.dcnv3_backward(inp, off, m, *args, gout, step, rc)
    rgrads = ref.dcnv3_backward(inp, off, m, *args, gout, step, rc)
    torch.cuda.synchronize()
    return out, rout, grads, rgrads


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_golden_cases_through_the_reference_kernels_f32(ops, ref, case):
    dev = "cuda"
    inp, off, m, gout = (case.t(k).to(dev, torch.float32).contiguous() for k in ("input", "offset", "mask", "grad_output"))
    out, rout, grads, rgrads = _both(ops, ref, inp, off, m, gout, case.args, 256, case.rc)
    assert out.shape == rout.shape and _rel(out, rout) < 1e-5
    for got, want, name in zip(grads, rgrads, ("grad_input", "grad_offset", "grad_mask")):
        assert got.shape == want.shape and got.dtype == want.dtype, name
        assert _rel(got, want) < 1e-4, name


@pytest.mark.parametrize("dist", ["T", "M"])
@pytest.mark.parametrize("dtype,ftol,gtol", [(torch.float32, 1e-5, 1e-4), (torch.float16, 2e-3, 5e-3)], ids=["f32", "f16"])
def test_config_K_vs_reference_kernels(ops, ref, dist, dtype, ftol, gtol):
    """BASELINE configs[0]/[1] shape (N=8): 64x64x256, G=8, 3x3 s1 p1."""
    inp, off, m, gout = _inputs(8, 64, 64, 8, 32, 9, 64, 64, dist, dtype)
    args = (3, 3, 1, 1, 1, 1, 1, 1, 8, 32, 1.0)
    out, rout, grads, rgrads = _both(ops, ref, inp, off, m, gout, args)
    assert _rel(out, rout) < ftol
    for got, want, name in zip(grads, rgrads, ("grad_input", "grad_offset", "grad_mask")):
        assert _rel(got, want) < gtol, name


@pytest.mark.parametrize("H", [64, 32, 16])
@pytest.mark.parametrize("dtype,ftol,gtol", [(torch.float32, 1e-5, 1e-4), (torch.float16, 2e-3, 5e-3)], ids=["f32", "f16"])
def test_in_model_stride2_flat_offset_vs_reference_kernels(ops, ref, H, dtype, ftol, gtol):
    """The three MAPEncoder calls (G=4, gc=64, s2, offsets/mask at INPUT resolution, conv_pnp_net.py:264-272): the kernels
    read the flat prefix of the offset/mask buffers (cuh:229,243-244); grads of the unread rows are zero on both sides."""
    inp, off, m, gout = _inputs(8, H, H, 4, 64, 9, H // 2, H // 2, "M", dtype, full_res=True)
    args = (3, 3, 2, 2, 1, 1, 1, 1, 4, 64, 1.0)
    out, rout, grads, rgrads = _both(ops, ref, inp, off, m, gout, args)
    assert out.shape == (8, H // 2, H // 2, 256) and _rel(out, rout) < ftol
    for got, want, name in zip(grads, rgrads, ("grad_input", "grad_offset", "grad_mask")):
        assert got.shape == want.shape, name
        assert _rel(got, want) < gtol, name
    n_read = 8 * (H // 2) ** 2 * 4 * 9
    assert torch.count_nonzero(grads[2].flatten()[n_read:]) == 0 and torch.count_nonzero(rgrads[2].flatten()[n_read:]) == 0


def test_batch_coupling_quirk_matches_the_reference_kernels(ops, ref):
    """at stride 2 RoI b reads the offset rows of RoI b//4 -- perturbing RoI 7's offsets changes
    nothing, perturbing RoI 0's changes outputs 0..3, in the reference kernels and in ours alike."""
    inp, off, m, gout = _inputs(8, 32, 32, 4, 64, 9, 16, 16, "M", torch.float32, full_res=True)
    args = (3, 3, 2, 2, 1, 1, 1, 1, 4, 64, 1.0)
    for impl in (ops, ref):
        base = impl.dcnv3_forward(inp, off, m, *args, 256, 0)
        o7 = off.clone(); o7[7] += 1.5
        assert torch.equal(impl.dcnv3_forward(inp, o7, m, *args, 256, 0), base)
        o0 = off.clone(); o0[0] += 1.5
        changed = (impl.dcnv3_forward(inp, o0, m, *args, 256, 0) != base).flatten(1).any(1).cpu().tolist()
        assert changed == [True] * 4 + [False] * 4


@pytest.mark.parametrize("gc", [16, 30, 71])
def test_odd_group_channels_and_remove_center_vs_reference_kernels(ops, ref, gc):
    G = 3
    inp, off, m, gout = _inputs(2, 20, 17, G, gc, 8, 20, 17, "T", torch.float32)
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 0.7)
    out, rout, grads, rgrads = _both(ops, ref, inp, off, m, gout, args, rc=1)
    assert _rel(out, rout) < 1e-5
    for got, want, name in zip(grads, rgrads, ("grad_input", "grad_offset", "grad_mask")):
        assert _rel(got, want) < 1e-4, name


def test_im2col_step_chunking_vs_reference_kernels(ops, ref):
    """N=512 with im2col_step=256: the reference loops over two chunks (dcnv3_cuda.cu:59-83), we address flat."""
    inp, off, m, gout = _inputs(512, 16, 16, 4, 32, 9, 16, 16, "T", torch.float32)
    args = (3, 3, 1, 1, 1, 1, 1, 1, 4, 32, 1.0)
    out, rout, grads, rgrads = _both(ops, ref, inp, off, m, gout, args, step=256)
    assert _rel(out, rout) < 1e-5
    for got, want, name in zip(grads, rgrads, ("grad_input", "grad_offset", "grad_mask")):
        assert _rel(got, want) < 1e-4, name


def test_error_behaviour_matches_the_reference_extension(ops, ref):
    """src/dcnv3.h:37 (CPU tensors), dcnv3_cuda.cu:29-53 (contiguity, batch % im2col_step, C == G*gc): both raise RuntimeError."""
    inp, off, m, gout = _inputs(6, 8, 8, 2, 16, 9, 8, 8, "T", torch.float32)
    args = (3, 3, 1, 1, 1, 1, 1, 1, 2, 16, 1.0)
    for impl in (ops, ref):
        with pytest.raises(RuntimeError):
            impl.dcnv3_forward(inp.cpu(), off.cpu(), m.cpu(), *args, 256, 0)
        with pytest.raises(RuntimeError):
            impl.dcnv3_forward(inp.transpose(1, 2), off, m, *args, 256, 0)
        with pytest.raises(RuntimeError):
            impl.dcnv3_forward(inp, off, m, *args, 4, 0)            # 6 % min(6, 4) != 0
        with pytest.raises(RuntimeError):
            impl.dcnv3_forward(inp, off, m, *args[:8], 2, 15, 1.0, 256, 0)   # C != G*gc


SWEEP = [
    # (N, H, W, G, gc, k, stride, pad, dil, scale, remove_center, dist)
    (2, 17, 23, 2, 32, 3, 1, 1, 1, 1.0, 0, "T"),
    (2, 17, 23, 2, 32, 3, 1, 1, 2, 1.0, 0, "T"),      # dilation 2
    (2, 16, 16, 4, 16, 5, 1, 2, 1, 0.5, 0, "T"),      # 5x5, offset_scale 0.5
    (2, 16, 16, 4, 16, 5, 1, 2, 1, 2.0, 1, "M"),      # 5x5 without its centre
    (3, 21, 13, 1, 64, 3, 2, 1, 1, 1.0, 0, "M"),      # stride 2, odd sizes, one group
    (4, 12, 12, 8, 8, 3, 1, 0, 1, 1.0, 0, "T"),       # no padding: out 10x10
    (2, 9, 31, 3, 24, 3, 3, 1, 1, 1.3, 1, "T"),       # stride 3, gc 24 (generic kernels)
    (1, 64, 64, 8, 32, 3, 1, 1, 1, 4.0, 0, "T"),      # offsets up to 40 px: most samples leave the image
    (5, 8, 8, 2, 128, 3, 1, 1, 1, 1.0, 0, "M"),       # wide groups
    (2, 33, 33, 4, 64, 1, 1, 0, 1, 1.0, 0, "M"),      # 1x1 kernel: one sampling point
]


@pytest.mark.parametrize("cfg", SWEEP, ids=[f"N{c[0]}_{c[1]}x{c[2]}_G{c[3]}x{c[4]}_k{c[5]}s{c[6]}p{c[7]}d{c[8]}_sc{c[9]}_rc{c[10]}_{c[11]}" for c in SWEEP])
@pytest.mark.parametrize("dtype,ftol,gtol", [(torch.float64, 1e-12, 1e-10), (torch.float32, 1e-5, 1e-4), (torch.float16, 2e-3, 5e-3)],
                         ids=["f64", "f32", "f16"])
def test_geometry_sweep_vs_reference_kernels(ops, ref, cfg, dtype, ftol, gtol):
    """Kernel size / stride / padding / dilation / offset_scale / remove_center / group shapes the golden file does not hold,
    in every dtype the reference dispatches (dcnv3_cuda.cu:69: double, float, half), against the reference's own kernels."""
    N, H, W, G, gc, k, s, pad, dil, scale, rc, dist = cfg
    if rc and k == 1:
        pytest.skip("remove_center needs more than one point")
    P = k * k - rc
    Ho = (H + 2 * pad - (dil * (k - 1) + 1)) // s + 1
    Wo = (W + 2 * pad - (dil * (k - 1) + 1)) // s + 1
    inp, off, m, gout = _inputs(N, H, W, G, gc, P, Ho, Wo, dist, dtype, seed=11)
    args = (k, k, s, s, pad, pad, dil, dil, G, gc, scale)
    out, rout, grads, rgrads = _both(ops, ref, inp, off, m, gout, args, rc=rc)
    assert out.shape == rout.shape == (N, Ho, Wo, G * gc) and _rel(out, rout) < ftol
    for got, want, name in zip(grads, rgrads, ("grad_input", "grad_offset", "grad_mask")):
        assert got.shape == want.shape and _rel(got, want) < gtol, name
