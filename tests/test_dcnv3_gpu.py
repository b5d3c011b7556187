"""GPU parity tests for the DCNv3 core: CUDA kernels (through the C ABI) vs the CPU oracle and the golden
vectors produced by the reference's dcnv3_core_pytorch.

Tolerances (north_star: 1e-4 relative in fp32, indices/bounds bit-exact, bf16 stated here):
  * f64: 2e-6 max-norm relative vs golden (limited by the reference PyTorch path's f32 reference points,
    see tests/test_oracle_golden.py), 1e-12 vs the f64 C oracle;
  * f32: 1e-5 vs the f32 C oracle (same arithmetic, only summation-order / FMA-contraction differences) and
    1e-4 vs golden f64; grads 1e-4 (atomic accumulation order is not defined, as in the reference);
  * bf16 / f16 storage, fp32 accumulate: compared with the f32 oracle evaluated on the SAME rounded inputs;
    BF16_TOL = 8e-3 (two bf16 ulps of the largest magnitude) for outputs and grads, F16_TOL = 1e-3.
"""
import pytest
import torch

from golden_util import CASES, CASE_IDS

pytestmark = pytest.mark.gpu

BF16_TOL = 8e-3
F16_TOL = 1e-3


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    import givepose_b200.functions as F
    return F


@pytest.fixture(scope="module")
def O():
    from oracle import dcnv3
    dcnv3.build()
    return dcnv3


def _fwd(ops, case, dtype, inp=None, off=None, msk=None):
    dev = "cuda"
    inp = case.t("input") if inp is None else inp
    off = case.t("offset") if off is None else off
    msk = case.t("mask") if msk is None else msk
    return ops.dcnv3_forward(inp.to(dev, dtype).contiguous(), off.to(dev, dtype).contiguous(),
                             msk.to(dev, dtype).contiguous(), *case.args, 256, case.rc)


def _bwd(ops, case, dtype):
    dev = "cuda"
    return ops.dcnv3_backward(case.t("input").to(dev, dtype), case.t("offset").to(dev, dtype),
                              case.t("mask").to(dev, dtype), *case.args, case.t("grad_output").to(dev, dtype), 256,
                              case.rc)


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_forward_f64_vs_golden(ops, case):
    assert _rel(_fwd(ops, case, torch.float64), case.t("out_f64")) < 2e-6


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_forward_f32_vs_golden(ops, case):
    out = _fwd(ops, case, torch.float32)
    assert out.shape == case.t("out_f64").shape and out.dtype == torch.float32
    assert _rel(out, case.t("out_f64")) < 1e-4


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_backward_f64_vs_golden(ops, case):
    gi, go, gm = _bwd(ops, case, torch.float64)
    for got, key in ((gi, "grad_input_f64"), (go, "grad_offset_f64"), (gm, "grad_mask_f64")):
        assert got.shape == case.t(key).shape, key
        assert _rel(got, case.t(key)) < 2e-6, key


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_backward_f32_vs_golden(ops, case):
    gi, go, gm = _bwd(ops, case, torch.float32)
    for got, key in ((gi, "grad_input_f64"), (go, "grad_offset_f64"), (gm, "grad_mask_f64")):
        assert _rel(got, case.t(key)) < 1e-4, key


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_index_and_bounds_bit_exact(ops, O, case):
    """floor()ed corners and in-range / per-corner flags: kernel device function vs C restatement, bit-exact."""
    off = case.t("offset")
    for dtype in (torch.float32, torch.float64):
        hw_o, fl_o = O.index(off.to(dtype), case.N, case.H, case.W, *case.args[:8], case.G, case.scale, case.rc)
        hw_g, fl_g = ops.dcnv3_sample_index(off.to("cuda", dtype), case.N, case.H, case.W, *case.args[:8], case.G,
                                            case.scale, case.rc)
        assert torch.equal(hw_g.cpu(), hw_o) and torch.equal(fl_g.cpu(), fl_o)


def _rand_case(gen, N, H, W, G, gc, k, s, pad, dil, rc, dist, full_res):
    P = k * k - rc
    f = lambda n: (n + 2 * pad - (dil * (k - 1) + 1)) // s + 1
    Ho, Wo = f(H), f(W)
    Hm, Wm = (H, W) if full_res else (Ho, Wo)
    if dist == "T":
        inp = torch.rand(N, H, W, G * gc, generator=gen) * 0.01
        off = torch.rand(N, Hm, Wm, G * P * 2, generator=gen) * 10
        m = torch.rand(N, Hm, Wm, G, P, generator=gen) + 1e-5
        m = m / m.sum(-1, keepdim=True)
    else:
        inp = torch.randn(N, H, W, G * gc, generator=gen)
        off = torch.randn(N, Hm, Wm, G * P * 2, generator=gen)
        m = torch.softmax(torch.randn(N, Hm, Wm, G, P, generator=gen), -1)
    return inp, off, m.reshape(N, Hm, Wm, G * P), torch.randn(N, Ho, Wo, G * gc, generator=gen)


# (name, N, H, W, G, gc, k, s, pad, dil, scale, rc, dist, full_res)
ORACLE_CASES = [
    ("config1_K_N8_distT", 8, 64, 64, 8, 32, 3, 1, 1, 1, 1.0, 0, "T", False),
    ("config1_K_N8_distM", 8, 64, 64, 8, 32, 3, 1, 1, 1, 1.0, 0, "M", False),
    ("model_M_64to32", 8, 64, 64, 4, 64, 3, 2, 1, 1, 1.0, 0, "M", True),
    ("model_M_32to16", 8, 32, 32, 4, 64, 3, 2, 1, 1, 1.0, 0, "M", True),
    ("model_M_16to8", 8, 16, 16, 4, 64, 3, 2, 1, 1, 1.0, 0, "M", True),
    ("ragged_tile_edges", 3, 13, 19, 2, 16, 3, 1, 1, 1, 1.3, 0, "T", False),
    ("gc8_L2", 2, 9, 9, 4, 8, 3, 1, 1, 1, 1.0, 0, "M", False),
    ("gc4_L1", 2, 9, 9, 4, 4, 3, 1, 1, 1, 1.0, 0, "M", False),
    ("gc128_L32", 1, 9, 9, 2, 128, 3, 1, 1, 1, 1.0, 0, "M", False),
    ("gc1025_generic", 1, 6, 6, 2, 1025, 3, 1, 1, 1, 2.0, 0, "T", False),
    ("k5_rc_generic_loops", 2, 11, 11, 2, 16, 5, 1, 2, 1, 1.0, 1, "M", False),
    ("dil3_s2", 2, 17, 15, 2, 32, 3, 2, 3, 3, 0.5, 0, "M", False),
]


@pytest.mark.parametrize("spec", ORACLE_CASES, ids=[c[0] for c in ORACLE_CASES])
def test_f32_vs_c_oracle_fwd_bwd(ops, O, spec):
    name, N, H, W, G, gc, k, s, pad, dil, scale, rc, dist, full = spec
    gen = torch.Generator().manual_seed(3)
    inp, off, m, gout = _rand_case(gen, N, H, W, G, gc, k, s, pad, dil, rc, dist, full)
    args = (k, k, s, s, pad, pad, dil, dil, G, gc, scale)
    ref = O.forward(inp, off, m, *args, rc)
    out = ops.dcnv3_forward(inp.cuda(), off.cuda(), m.cuda(), *args, 256, rc)
    assert _rel(out, ref) < 1e-5
    rgi, rgo, rgm = O.backward(inp, off, m, gout, *args, rc)
    gi, go, gm = ops.dcnv3_backward(inp.cuda(), off.cuda(), m.cuda(), *args, gout.cuda(), 256, rc)
    assert gi.shape == inp.shape and go.shape == off.shape and gm.shape == m.shape
    assert _rel(gi, rgi) < 1e-4 and _rel(go, rgo) < 1e-4 and _rel(gm, rgm) < 1e-4
    if full:   # stride-2 quirk: rows beyond the flat prefix are never read, their grads are exactly zero
        rows = N * ref.shape[1] * ref.shape[2]
        assert go.view(-1, off.shape[-1])[rows:].abs().max().item() == 0
        assert gm.view(-1, m.shape[-1])[rows:].abs().max().item() == 0
    hw_o, fl_o = O.index(off, N, H, W, *args[:8], G, scale, rc)
    hw_g, fl_g = ops.dcnv3_sample_index(off.cuda(), N, H, W, *args[:8], G, scale, rc)
    assert torch.equal(hw_g.cpu(), hw_o) and torch.equal(fl_g.cpu(), fl_o)


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, BF16_TOL), (torch.float16, F16_TOL)])
@pytest.mark.parametrize("spec", ORACLE_CASES[:6] + ORACLE_CASES[9:10], ids=[c[0] for c in ORACLE_CASES[:6] + ORACLE_CASES[9:10]])
def test_16bit_storage_vs_oracle_on_rounded_inputs(ops, O, spec, dtype, tol):
    name, N, H, W, G, gc, k, s, pad, dil, scale, rc, dist, full = spec
    gen = torch.Generator().manual_seed(5)
    inp, off, m, gout = _rand_case(gen, N, H, W, G, gc, k, s, pad, dil, rc, dist, full)
    inp, off, m, gout = (t.to(dtype) for t in (inp, off, m, gout))          # the rounded values are the inputs
    args = (k, k, s, s, pad, pad, dil, dil, G, gc, scale)
    ref = O.forward(inp.float(), off.float(), m.float(), *args, rc)
    out = ops.dcnv3_forward(inp.cuda(), off.cuda(), m.cuda(), *args, 256, rc)
    assert out.dtype == dtype
    assert _rel(out, ref) < tol
    rgi, rgo, rgm = O.backward(inp.float(), off.float(), m.float(), gout.float(), *args, rc)
    gi, go, gm = ops.dcnv3_backward(inp.cuda(), off.cuda(), m.cuda(), *args, gout.cuda(), 256, rc)
    assert gi.dtype == dtype and go.dtype == dtype and gm.dtype == dtype
    assert _rel(gi, rgi) < tol and _rel(go, rgo) < tol and _rel(gm, rgm) < tol
    # coordinates are computed in fp32 from the same 16-bit offsets: indices / bounds stay bit-exact
    hw_o, fl_o = O.index(off.float(), N, H, W, *args[:8], G, scale, rc)
    hw_g, fl_g = ops.dcnv3_sample_index(off.cuda(), N, H, W, *args[:8], G, scale, rc)
    assert torch.equal(hw_g.cpu(), hw_o) and torch.equal(fl_g.cpu(), fl_o)


def test_index_edge_cases_bit_exact(ops, O):
    """Offsets that put samples exactly on integers, in (-1,0), in (H-1,H), at <= -1 and >= H, huge, and NaN."""
    N, H, W, G = 1, 7, 9, 2
    k, s, pad, dil = 3, 1, 1, 1
    P = 9
    base = torch.zeros(N, H, W, G * P * 2)
    specials = torch.tensor([0.0, -0.5, 0.5, -1.0, 1.0, -1.0000001, 0.99999994, 1e-7, -1e-7, float(H), float(-H),
                             H - 1.0, W - 1.0, 7.5, -7.5, 1e9, -1e9, float("nan"), float("inf"), -float("inf"),
                             2.5, 3.0000002, 5.9999995])
    gen = torch.Generator().manual_seed(11)
    idx = torch.randint(0, len(specials), base.shape, generator=gen)
    off = specials[idx]
    for scale in (1.0, 2.0, 0.7):
        hw_o, fl_o = O.index(off, N, H, W, k, k, s, s, pad, pad, dil, dil, G, scale, 0)
        hw_g, fl_g = ops.dcnv3_sample_index(off.cuda(), N, H, W, k, k, s, s, pad, pad, dil, dil, G, scale, 0)
        assert torch.equal(hw_g.cpu(), hw_o) and torch.equal(fl_g.cpu(), fl_o)


def test_softmax_fused_forward_matches_unfused(ops):
    gen = torch.Generator().manual_seed(2)
    N, H, W, G, gc = 4, 16, 16, 4, 64
    inp = torch.randn(N, H, W, G * gc, generator=gen).cuda()
    off = torch.randn(N, H, W, G * 18, generator=gen).cuda()
    logits = (torch.randn(N, H, W, G, 9, generator=gen) * 3).cuda()
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    ref = ops.dcnv3_forward(inp, off, torch.softmax(logits, -1).reshape(N, H, W, -1).contiguous(), *args, 256, 0)
    out = ops.dcnv3_forward(inp, off, logits.reshape(N, H, W, -1).contiguous(), *args, 256, 0, mask_is_logits=True)
    assert _rel(out, ref) < 1e-5


def test_im2col_step_semantics(ops):
    """dcnv3_cuda.cu:46-49: batch % min(batch, im2col_step) must be 0; otherwise chunking never changes results."""
    gen = torch.Generator().manual_seed(4)
    inp, off, m, _ = _rand_case(gen, 8, 8, 8, 2, 16, 3, 1, 1, 1, 0, "T", False)
    args = (3, 3, 1, 1, 1, 1, 1, 1, 2, 16, 1.0)
    a = ops.dcnv3_forward(inp.cuda(), off.cuda(), m.cuda(), *args, 256, 0)
    b = ops.dcnv3_forward(inp.cuda(), off.cuda(), m.cuda(), *args, 2, 0)
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError, match="must divide im2col_step"):
        ops.dcnv3_forward(inp.cuda(), off.cuda(), m.cuda(), *args, 3, 0)


def test_autograd_function_matches_reference_test_protocol(ops, O):
    """network/ops_dcnv3/test.py:157-217 protocol: out.sum().backward() through DCNv3Function vs the oracle."""
    case = next(c for c in CASES if c.name == "test_py_fixture")
    x = case.t("input").cuda().requires_grad_(True)
    o = case.t("offset").cuda().requires_grad_(True)
    m = case.t("mask").cuda().requires_grad_(True)
    out = ops.DCNv3Function.apply(x, o, m, *case.args, 2, case.rc)
    out.sum().backward()
    ones = torch.ones(out.shape)
    rgi, rgo, rgm = O.backward(case.t("input"), case.t("offset"), case.t("mask"), ones, *case.args, case.rc)
    # the reference's own tolerance for this check is rtol=1e-2, atol=1e-3 (test.py:198-217); we hold 1e-4
    assert _rel(x.grad, rgi) < 1e-4 and _rel(o.grad, rgo) < 1e-4 and _rel(m.grad, rgm) < 1e-4


# ---- BASELINE config 2 at full size: size-independent properties ------------------------------------------

@pytest.fixture(scope="module")
def full_size():
    gen = torch.Generator(device="cuda").manual_seed(3)
    N, H, W, G, gc = 64, 64, 64, 8, 32
    inp = torch.rand(N, H, W, G * gc, generator=gen, device="cuda") * 0.01
    off = torch.rand(N, H, W, G * 18, generator=gen, device="cuda") * 10
    m = torch.rand(N, H, W, G, 9, generator=gen, device="cuda") + 1e-5
    m = (m / m.sum(-1, keepdim=True)).reshape(N, H, W, G * 9).contiguous()
    gout = torch.randn(N, H, W, G * gc, generator=gen, device="cuda")
    return inp, off, m, gout, (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)


def test_full_size_identity_sampling_is_exact(ops, full_size):
    """zero offsets + one-hot mask on the centre point => out == input bit-exactly (w1 == 1, other corners x0)."""
    inp, off, m, gout, args = full_size
    N, H, W, _ = inp.shape
    onehot = torch.zeros(N, H, W, 8, 9, device="cuda")
    onehot[..., 4] = 1.0   # p = i*kh + j with i = j = 1
    out = ops.dcnv3_forward(inp, torch.zeros_like(off), onehot.reshape(N, H, W, 72), *args, 256, 0)
    assert torch.equal(out, inp)


def test_full_size_adjoint_identities(ops, full_size):
    """The op is linear in `input` and in `mask`:  <f(x,m), g> == <x, grad_input(g)> == <m, grad_mask(g)>."""
    inp, off, m, gout, args = full_size
    out = ops.dcnv3_forward(inp, off, m, *args, 256, 0)
    gi, go, gm = ops.dcnv3_backward(inp, off, m, *args, gout, 256, 0)
    lhs = (out.double() * gout.double()).sum().item()
    assert abs((inp.double() * gi.double()).sum().item() - lhs) < 1e-5 * abs(lhs) + 1e-6
    assert abs((m.double() * gm.double()).sum().item() - lhs) < 1e-5 * abs(lhs) + 1e-6
    # linearity in the input, checked on the outputs themselves
    out2 = ops.dcnv3_forward(inp * 2, off, m, *args, 256, 0)
    assert torch.equal(out2, out * 2)      # scaling by 2 is exact in binary floating point


def test_full_size_vs_c_oracle_sampled_images(ops, O, full_size):
    """Full-size launch, oracle on a strided subset of the 64 images (each image is independent at stride 1)."""
    inp, off, m, gout, args = full_size
    out = ops.dcnv3_forward(inp, off, m, *args, 256, 0)
    gi, go, gm = ops.dcnv3_backward(inp, off, m, *args, gout, 256, 0)
    sel = [0, 21, 42, 63]
    ci, co, cm, cg = (t[sel].cpu().contiguous() for t in (inp, off, m, gout))
    ref = O.forward(ci, co, cm, *args, 0)
    assert _rel(out[sel], ref) < 1e-5
    rgi, rgo, rgm = O.backward(ci, co, cm, cg, *args, 0)
    assert _rel(gi[sel], rgi) < 1e-4 and _rel(go[sel], rgo) < 1e-4 and _rel(gm[sel], rgm) < 1e-4


def test_host_buffer_entry_points(ops, O):
    """gp_dcnv3_forward_host / gp_dcnv3_backward_host: the end-to-end (H2D + kernels + D2H) C-ABI calls."""
    import ctypes
    from givepose_b200 import _lib
    gen = torch.Generator().manual_seed(9)
    N, H, W, G, gc = 4, 16, 16, 4, 64
    inp, off, m, gout = _rand_case(gen, N, H, W, G, gc, 3, 2, 1, 1, 0, "M", True)
    args = (3, 3, 2, 2, 1, 1, 1, 1, G, gc, 1.0)
    Ho = Wo = 8
    d = _lib.DCNv3Desc(N, H, W, G, gc, 3, 3, 2, 2, 1, 1, 1, 1, 0, Ho, Wo, 1.0)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    out = torch.empty(N, Ho, Wo, G * gc)
    _lib.check(_lib.lib.gp_dcnv3_forward_host(vp(inp), vp(off), vp(m), vp(out), off.numel(), m.numel(),
                                              ctypes.byref(d), _lib.GP_F32, 0), "fwd_host")
    assert _rel(out, O.forward(inp, off, m, *args, 0)) < 1e-5
    gi, go, gm = torch.empty_like(inp), torch.empty_like(off), torch.empty_like(m)
    _lib.check(_lib.lib.gp_dcnv3_backward_host(vp(inp), vp(off), vp(m), vp(gout), vp(gi), vp(go), vp(gm), off.numel(),
                                               m.numel(), ctypes.byref(d), _lib.GP_F32, 0), "bwd_host")
    rgi, rgo, rgm = O.backward(inp, off, m, gout, *args, 0)
    assert _rel(gi, rgi) < 1e-4 and _rel(go, rgo) < 1e-4 and _rel(gm, rgm) < 1e-4
    # pipelined forward+backward over RoI chunks (stride 2, full-resolution offset tensors: flat addressing is chunk-exact)
    for chunks, dtype, code, tol in ((3, torch.float32, _lib.GP_F32, 1e-4), (4, torch.bfloat16, _lib.GP_BF16, BF16_TOL), (1, torch.float32, _lib.GP_F32, 1e-4)):
        hi, ho_, hm, hg = (t.to(dtype).pin_memory() for t in (inp, off, m, gout))
        o2, gi2, go2, gm2 = (torch.full_like(t, 7).pin_memory() for t in (out.to(dtype), hi, ho_, hm))
        _lib.check(_lib.lib.gp_dcnv3_forward_backward_host(vp(hi), vp(ho_), vp(hm), vp(hg), vp(o2), vp(gi2), vp(go2), vp(gm2),
                                                           ho_.numel(), hm.numel(), ctypes.byref(d), code, 0, chunks), "fwd_bwd_host")
        ro = O.forward(hi.float(), ho_.float(), hm.float(), *args, 0)
        r1, r2, r3 = O.backward(hi.float(), ho_.float(), hm.float(), hg.float(), *args, 0)
        assert _rel(o2, ro) < tol and _rel(gi2, r1) < tol and _rel(go2, r2) < tol and _rel(gm2, r3) < tol, (chunks, dtype)
    _lib.lib.gp_host_cache_release()


# ---- backward variants: one-pass scatter kernel (mode 0) and split backward with the binned grad_input kernel (mode 1) ----

OPT_BWD_MODE, OPT_GIN_TH, OPT_GIN_TW, OPT_GIN_NT = 0, 1, 2, 3


@pytest.fixture
def bwd_options():
    from givepose_b200._lib import lib
    saved = [lib.gp_get_option(k) for k in range(5)]
    yield lib
    for k, v in enumerate(saved):
        lib.gp_set_option(k, v)


@pytest.mark.parametrize("mode,th,tw,nt", [(0, 8, 8, 192), (1, 8, 8, 192), (1, 8, 8, 128), (1, 16, 16, 256), (1, 4, 8, 192),
                                           (1, 16, 8, 256), (1, 2, 2, 192), (2, 8, 8, 192), (2, 4, 8, 192), (2, 16, 16, 192),
                                           (2, 2, 2, 192)])
@pytest.mark.parametrize("spec", ORACLE_CASES, ids=[c[0] for c in ORACLE_CASES])
def test_backward_variants_vs_c_oracle(ops, O, bwd_options, spec, mode, th, tw, nt):
    """Every backward variant the tuning knobs can select against the C oracle: f32 1e-4, bf16 8e-3 on rounded inputs."""
    name, N, H, W, G, gc, k, s, pad, dil, scale, rc, dist, full = spec
    lib = bwd_options
    lib.gp_set_option(OPT_BWD_MODE, mode)
    lib.gp_set_option(OPT_GIN_TH, th)
    lib.gp_set_option(OPT_GIN_TW, tw)
    lib.gp_set_option(OPT_GIN_NT, nt)
    gen = torch.Generator().manual_seed(7)
    inp, off, m, gout = _rand_case(gen, N, H, W, G, gc, k, s, pad, dil, rc, dist, full)
    args = (k, k, s, s, pad, pad, dil, dil, G, gc, scale)
    rgi, rgo, rgm = O.backward(inp, off, m, gout, *args, rc)
    gi, go, gm = ops.dcnv3_backward(inp.cuda(), off.cuda(), m.cuda(), *args, gout.cuda(), 256, rc)
    assert _rel(gi, rgi) < 1e-4 and _rel(go, rgo) < 1e-4 and _rel(gm, rgm) < 1e-4
    if gc <= 128:
        b = [t.to(torch.bfloat16) for t in (inp, off, m, gout)]
        rgi, rgo, rgm = O.backward(*(t.float() for t in b[:3]), b[3].float(), *args, rc)
        gi, go, gm = ops.dcnv3_backward(b[0].cuda(), b[1].cuda(), b[2].cuda(), *args, b[3].cuda(), 256, rc)
        assert _rel(gi, rgi) < BF16_TOL and _rel(go, rgo) < BF16_TOL and _rel(gm, rgm) < BF16_TOL


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("std", [3.0, 30.0, 300.0])
def test_binned_grad_input_window_overflow(ops, O, bwd_options, std, mode):
    """Offsets far beyond the 32-cell window of the aggregating kernels: the per-sample fallback scatter must give the same sums."""
    lib = bwd_options
    lib.gp_set_option(OPT_BWD_MODE, mode)
    gen = torch.Generator().manual_seed(13)
    N, H, W, G, gc = 2, 80, 72, 2, 32
    inp = torch.randn(N, H, W, G * gc, generator=gen)
    off = torch.randn(N, H, W, G * 18, generator=gen) * std
    m = torch.softmax(torch.randn(N, H, W, G, 9, generator=gen), -1).reshape(N, H, W, G * 9)
    gout = torch.randn(N, H, W, G * gc, generator=gen)
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    rgi, rgo, rgm = O.backward(inp, off, m, gout, *args, 0)
    gi, go, gm = ops.dcnv3_backward(inp.cuda(), off.cuda(), m.cuda(), *args, gout.cuda(), 256, 0)
    assert _rel(gi, rgi) < 1e-4 and _rel(go, rgo) < 1e-4 and _rel(gm, rgm) < 1e-4


def test_backward_elementwise_with_absolute_floor(ops, O, bwd_options):
    """Element-wise check (not max-norm): |got - ref| <= 1e-4 * |ref| + 1e-5 * max|ref| for every element of the three
    gradients at config K (N = 2), so a localised error in small-magnitude outputs cannot hide behind the largest one."""
    gen = torch.Generator().manual_seed(21)
    N, H, W, G, gc = 2, 64, 64, 8, 32
    for dist in ("T", "M"):
        inp, off, m, gout = _rand_case(gen, N, H, W, G, gc, 3, 1, 1, 1, 0, dist, False)
        args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
        ref_out = O.forward(inp.double(), off.double(), m.double(), *args, 0)
        out = ops.dcnv3_forward(inp.cuda(), off.cuda(), m.cuda(), *args, 256, 0).cpu().double()
        assert ((out - ref_out).abs() <= 1e-4 * ref_out.abs() + 1e-5 * ref_out.abs().max()).all()
        refs = O.backward(inp.double(), off.double(), m.double(), gout.double(), *args, 0)
        for mode in (0, 1, 2):
            bwd_options.gp_set_option(OPT_BWD_MODE, mode)
            got = ops.dcnv3_backward(inp.cuda(), off.cuda(), m.cuda(), *args, gout.cuda(), 256, 0)
            for g_, r_, nm in zip(got, refs, ("grad_input", "grad_offset", "grad_mask")):
                d = (g_.cpu().double() - r_).abs()
                # grad_offset is discontinuous where a sample sits on an integer (floor flips between f32 and f64): skip those
                tol = 1e-4 * r_.abs() + 1e-5 * r_.abs().max()
                bad = (d > tol)
                assert bad.float().mean().item() < (1e-5 if nm != "grad_input" else 1e-6), (dist, mode, nm, bad.sum().item())


def test_full_size_bf16_vs_c_oracle_sampled_images(ops, O, full_size):
    """BASELINE config 2 in bf16 (N = 64, the size bench.py --dtype bf16 times), forward + backward, C oracle (fp32 arithmetic
    on the same bf16-rounded inputs) on a strided subset of the images; indices / bounds stay bit-exact."""
    inp, off, m, gout, args = (t.to(torch.bfloat16) if torch.is_tensor(t) else t for t in full_size)
    out = ops.dcnv3_forward(inp, off, m, *args, 256, 0)
    gi, go, gm = ops.dcnv3_backward(inp, off, m, *args, gout, 256, 0)
    assert out.dtype == torch.bfloat16 and gi.dtype == torch.bfloat16
    sel = [0, 21, 42, 63]
    ci, co, cm, cg = (t[sel].float().cpu().contiguous() for t in (inp, off, m, gout))
    assert _rel(out[sel], O.forward(ci, co, cm, *args, 0)) < BF16_TOL
    rgi, rgo, rgm = O.backward(ci, co, cm, cg, *args, 0)
    assert _rel(gi[sel], rgi) < BF16_TOL and _rel(go[sel], rgo) < BF16_TOL and _rel(gm[sel], rgm) < BF16_TOL
    hw_o, fl_o = O.index(co, len(sel), 64, 64, *args[:8], 8, 1.0, 0)
    hw_g, fl_g = ops.dcnv3_sample_index(off[sel].contiguous(), len(sel), 64, 64, *args[:8], 8, 1.0, 0)
    assert torch.equal(hw_g.cpu(), hw_o) and torch.equal(fl_g.cpu(), fl_o)


def test_forward_inf_next_to_the_border_documented_deviation(ops, O):
    """for a footprint corner OUTSIDE the image the forward reads a pixel of the same footprint that lies
    inside, with weight 0.  For finite inputs that is exactly the reference's "outside counts as 0" (cuh:55-75); an Inf in that
    borrowed pixel gives 0 * Inf = NaN where the reference keeps the value it computes without it.  This test pins both
    halves: finite inputs agree everywhere; with Inf border pixels every output the reference keeps finite is either equal
    or NaN here (never a different finite number), and outputs that do not touch the border are untouched."""
    gen = torch.Generator().manual_seed(17)
    N, H, W, G, gc = 1, 12, 12, 2, 16
    inp = torch.randn(N, H, W, G * gc, generator=gen)
    off = (torch.rand(N, H, W, G * 18, generator=gen) - 0.5) * 3
    m = torch.softmax(torch.randn(N, H, W, G, 9, generator=gen), -1).reshape(N, H, W, G * 9)
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    assert _rel(ops.dcnv3_forward(inp.cuda(), off.cuda(), m.cuda(), *args, 256, 0), O.forward(inp, off, m, *args, 0)) < 1e-5
    bad = inp.clone()
    bad[:, 0, :, :] = float("inf")            # the whole first image row
    ref = O.forward(bad, off, m, *args, 0)
    got = ops.dcnv3_forward(bad.cuda(), off.cuda(), m.cuda(), *args, 256, 0).cpu()
    fin = torch.isfinite(ref)
    same = torch.isclose(got, ref, rtol=1e-5, atol=1e-6)
    assert (same | torch.isnan(got))[fin].all()            # never a different finite number
    assert (~torch.isfinite(got[~fin])).all()              # where the reference overflows, so do we
    # rows far from the poisoned border (no footprint can reach image row 0 with |offset| < 1.5 + kernel reach 1) are exact
    assert torch.isfinite(got[:, 5:]).all() and _rel(got[:, 5:], ref[:, 5:]) < 1e-5


# ---- forward variants: GP_OPT_FWD_MODE 0 = dcnv3_fwd_tile (per-thread row reads, 24-byte records), 1 = dcnv3_fwd_rows (3x3: offset /
# mask rows staged by TMA, 16-byte rank-one records; shapes outside its rules fall back to mode 0 inside the library) ----------------

OPT_FWD_MODE = 4


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("spec", ORACLE_CASES, ids=[c[0] for c in ORACLE_CASES])
def test_forward_modes_vs_c_oracle(ops, O, bwd_options, spec, mode):
    """Both forward kernels against the C oracle: fp32 1e-5, bf16 / f16 on rounded inputs, the fused softmax, and the packed
    offset||logits rows of the fused offset/mask Linear (pitch padded to a multiple of 8 elements)."""
    name, N, H, W, G, gc, k, s, pad, dil, scale, rc, dist, full = spec
    bwd_options.gp_set_option(OPT_FWD_MODE, mode)
    gen = torch.Generator().manual_seed(11)
    inp, off, m, _ = _rand_case(gen, N, H, W, G, gc, k, s, pad, dil, rc, dist, full)
    args = (k, k, s, s, pad, pad, dil, dil, G, gc, scale)
    ref = O.forward(inp, off, m, *args, rc)
    assert _rel(ops.dcnv3_forward(inp.cuda(), off.cuda(), m.cuda(), *args, 256, rc), ref) < 1e-5
    P = k * k - rc
    logits = torch.randn(*m.shape[:-1], G, P, generator=gen) * 2
    sm = torch.softmax(logits, -1).reshape(m.shape)
    ref_sm = O.forward(inp, off, sm, *args, rc)
    out_sm = ops.dcnv3_forward(inp.cuda(), off.cuda(), logits.reshape(m.shape).contiguous().cuda(), *args, 256, rc, mask_is_logits=True)
    assert _rel(out_sm, ref_sm) < 1e-5
    if gc <= 128:
        from givepose_b200.functions import dcnv3_forward_packed
        pitch = (G * P * 3 + 7) // 8 * 8
        packed = torch.zeros(*off.shape[:-1], pitch)
        packed[..., :G * P * 2] = off
        packed[..., G * P * 2:G * P * 3] = logits.reshape(m.shape)
        assert _rel(dcnv3_forward_packed(inp.cuda(), packed.cuda(), *args, 256, rc), ref_sm) < 1e-5
        for dtype, tol in ((torch.bfloat16, BF16_TOL), (torch.float16, F16_TOL)):
            i16, o16, m16, p16 = (t.to(dtype) for t in (inp, off, m, packed))
            r16 = O.forward(i16.float(), o16.float(), m16.float(), *args, rc)
            got = ops.dcnv3_forward(i16.cuda(), o16.cuda(), m16.cuda(), *args, 256, rc)
            assert got.dtype == dtype and _rel(got, r16) < tol
            l16 = p16[..., G * P * 2:G * P * 3].float().reshape(*m.shape[:-1], G, P)
            r16p = O.forward(i16.float(), p16[..., :G * P * 2].float().contiguous(), torch.softmax(l16, -1).reshape(m.shape), *args, rc)
            assert _rel(dcnv3_forward_packed(i16.cuda(), p16.cuda(), *args, 256, rc), r16p) < tol


def test_forward_modes_agree_at_full_size(ops, bwd_options, full_size):
    """Config 2 (N = 64, 64x64x256, G = 8): the TMA-row kernel against the per-thread-row kernel on the same inputs -- the same
    sums up to the association of the mask-folded bilinear weights (2e-6 of the output's max)."""
    inp, off, m = full_size[:3]
    args = (3, 3, 1, 1, 1, 1, 1, 1, 8, 32, 1.0)
    outs = []
    for mode in (0, 1):
        bwd_options.gp_set_option(OPT_FWD_MODE, mode)
        outs.append(ops.dcnv3_forward(inp, off, m, *args, 256, 0))
    assert _rel(outs[1], outs[0]) < 2e-6


def test_forward_kernel_selection_is_observable(ops, bwd_options):
    """GP_OPT_LAST_FWD_KERNEL reports which kernel a forward call launched: the TMA-row kernel for the benchmark configuration
    (fp32 and bf16), the per-thread-row kernel when a row pitch is not a multiple of 16 bytes (bf16, G = 2: 72-byte mask rows) or
    when mode 0 is selected, the generic kernel for odd group widths -- so a silent fall-back cannot hide behind a label."""
    lib = bwd_options
    gen = torch.Generator().manual_seed(1)

    def run(N, H, W, G, gc, dtype):
        inp = torch.randn(N, H, W, G * gc, generator=gen).to("cuda", dtype)
        off = torch.randn(N, H, W, G * 18, generator=gen).to("cuda", dtype)
        m = torch.softmax(torch.randn(N, H, W, G, 9, generator=gen), -1).reshape(N, H, W, G * 9).to("cuda", dtype)
        ops.dcnv3_forward(inp, off, m, 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0, 256, 0)
        return lib.gp_get_option(5)

    lib.gp_set_option(OPT_FWD_MODE, 1)
    assert run(2, 64, 64, 8, 32, torch.float32) == 1
    assert run(2, 64, 64, 8, 32, torch.bfloat16) == 1
    assert run(2, 16, 16, 2, 32, torch.bfloat16) == 0      # mask rows of 2 * 9 * 2 = 36 bytes: no TMA
    assert run(1, 9, 9, 2, 30, torch.float32) == 2          # gc = 30: generic kernel
    lib.gp_set_option(OPT_FWD_MODE, 0)
    assert run(2, 64, 64, 8, 32, torch.float32) == 0
