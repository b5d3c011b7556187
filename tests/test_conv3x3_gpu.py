"""Decoder 3x3 convolution as a hand-written tcgen05 implicit GEMM (``gp_conv3x3_gn_bf16``, ``conv3x3_tc.cu``; reference
``network/xyz_head.py:195-366`` -> ``ConvModule`` = Conv2d(256, 256, 3, padding=1, bias=False) -> GroupNorm(32) -> GELU) against a
plain PyTorch fp32 reference of the same op on the same bf16-rounded operands (TF32 off).

Tolerances: the result is rounded to bf16 once (2^-9 relative) on top of fp32 accumulation in a different order: 1e-2 relative to
the output's max magnitude (measured ~4e-3), plus an element-wise bar |got - ref| <= 2^-7 |ref| + 1e-3 max|ref|.  The GroupNorm
statistics come from the fp32 accumulators: mean within 1e-3 of the group's std, rstd within 1e-3 relative."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[0, 1], ids=["one_cta", "cta_pair"])
def _variant(request):
    """Every test runs on both kernel variants: one CTA per tile, and the tcgen05 cta_group::2 CTA pair."""
    from givepose_b200._lib import lib
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    old = lib.gp_conv3x3_set_pair(request.param)
    yield request.param
    lib.gp_conv3x3_set_pair(old)
    torch.backends.cudnn.allow_tf32 = old_tf32


def _case(N, H, W, Cin, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(N, H, W, Cin, generator=g) * scale).bfloat16().cuda()
    w = (torch.randn(256, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).bfloat16().cuda()
    return x, w


def _ref(x, w):
    return F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, 1, 1).permute(0, 2, 3, 1).contiguous()


# (N, H, W, Cin): the three decoder resolutions, more tiles than SMs (persistent loop, accumulator hand-back), other Cin
SHAPES = [(1, 16, 16, 256), (3, 16, 16, 256), (2, 32, 32, 256), (2, 64, 64, 256), (40, 64, 64, 256), (5, 32, 32, 64),
          (2, 16, 16, 1024), (1, 4, 64, 128), (2, 32, 8, 64), (148 * 3 + 5, 16, 16, 64)]


@pytest.mark.parametrize("N,H,W,Cin", SHAPES)
def test_conv3x3_matches_torch_fp32(N, H, W, Cin):
    from givepose_b200 import ops
    x, w = _case(N, H, W, Cin, N * 1000 + H + Cin)
    ref = _ref(x, w)
    y, stats = ops.conv3x3_gn_bf16(x, ops.pack_conv3x3_weight(w))
    torch.cuda.synchronize()
    assert y.shape == (N, H, W, 256) and y.dtype == torch.bfloat16
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-2, err
    # element-wise: a localised error (one wrong tap / border pixel / channel block) cannot hide behind the max norm
    bad = (y.float() - ref).abs() > ref.abs() * 2 ** -7 + 1e-3 * ref.abs().max()
    assert not bad.any(), (int(bad.sum()), bad.nonzero()[:5].tolist())
    # GroupNorm(32) statistics of the fp32 result, produced in the convolution's epilogue
    r = ref.view(N, H * W, 32, 8).permute(0, 2, 1, 3).reshape(N, 32, -1).double()
    mean, var = r.mean(-1), r.var(-1, unbiased=False)
    st = stats.view(N, 32, 2).double()
    assert ((st[..., 0] - mean).abs() / var.sqrt()).max().item() < 1e-3
    assert ((st[..., 1] - (var + 1e-5).rsqrt()).abs() / (var + 1e-5).rsqrt()).max().item() < 1e-3


def test_conv3x3_borders_are_zero_padded_exactly():
    """An all-ones image and a single-tap weight: every output equals 1 inside and 0 where the tap falls outside the image --
    the zero padding comes from TMA's out-of-bounds fill, checked tap by tap and exactly."""
    from givepose_b200 import ops
    N, H, W, Cin = 2, 16, 16, 64
    x = torch.ones(N, H, W, Cin, dtype=torch.bfloat16).cuda()
    for ky in range(3):
        for kx in range(3):
            w = torch.zeros(256, Cin, 3, 3)
            w[:, 0, ky, kx] = 1.0
            y, _ = ops.conv3x3_gn_bf16(x, ops.pack_conv3x3_weight(w.cuda()), stats=False)
            want = torch.ones(H, W)
            if ky == 0: want[0, :] = 0
            if ky == 2: want[H - 1, :] = 0
            if kx == 0: want[:, 0] = 0
            if kx == 2: want[:, W - 1] = 0
            assert torch.equal(y.float().cpu(), want[None, :, :, None].expand(N, H, W, 256)), (ky, kx)


def test_conv3x3_is_deterministic_and_stats_optional():
    from givepose_b200 import ops
    x, w = _case(4, 32, 32, 256, 7)
    wp = ops.pack_conv3x3_weight(w)
    y1, s1 = ops.conv3x3_gn_bf16(x, wp)
    y2, s2 = ops.conv3x3_gn_bf16(x, wp)
    y3, s3 = ops.conv3x3_gn_bf16(x, wp, stats=False)
    assert torch.equal(y1, y2) and torch.equal(s1, s2) and torch.equal(y1, y3) and s3 is None


def test_conv3x3_refuses_unsupported():
    from givepose_b200 import ops
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        ops.conv3x3_gn_bf16(torch.zeros(1, 16, 16, 256, dtype=torch.bfloat16), torch.zeros(256, 9 * 256, dtype=torch.bfloat16))
    x = torch.zeros(1, 8, 8, 256, dtype=torch.bfloat16).cuda()   # 64 pixels: not a whole 256-pixel tile
    with pytest.raises(RuntimeError, match="unsupported shapes"):
        ops.conv3x3_gn_bf16(x, torch.zeros(256, 9 * 256, dtype=torch.bfloat16).cuda())
    assert not ops.conv3x3_gn_supported(x, 256) and not ops.conv3x3_gn_supported(x.float(), 256)


@pytest.mark.parametrize("res", [16, 32, 64])
def test_conv_module_tc_equals_cudnn_path(res):
    """ConvModule (conv -> GN -> GELU [-> bilinear x2]) at bf16 inference: hand-written convolution + epilogue statistics against
    the cuDNN convolution + gn_stats pass.  Both round the conv output to bf16 once; the statistics differ only in summing
    fp32 accumulators instead of rounded values -> results agree to bf16 resolution of the activation (3e-2 of its max)."""
    import givepose_b200.posenet as P
    torch.manual_seed(res)
    m = P.ConvModule(256, 256).cuda().eval()
    with torch.no_grad():
        m.conv.weight.mul_(30.0)   # kaiming fan_out init is tiny: lift the activations to O(1)
        m.norm.weight.uniform_(0.5, 1.5)
        m.norm.bias.uniform_(-0.5, 0.5)
        x = torch.randn(3, res, res, 256, device="cuda").bfloat16()
        for up in (False, True):
            P.DECODER_CONV = "tc"
            assert m._tc(x)
            a = m.forward_nhwc(x, upsample2x=up)
            P.DECODER_CONV = "cudnn"
            assert not m._tc(x)
            b = m.forward_nhwc(x, upsample2x=up)
            P.DECODER_CONV = "tc"
            ref = F.gelu(F.group_norm(F.conv2d(x.float().permute(0, 3, 1, 2), m.conv.weight, None, 1, 1), 32, m.norm.weight, m.norm.bias, 1e-5))
            if up:
                ref = F.interpolate(ref, scale_factor=2, mode="bilinear", align_corners=True)
            ref = ref.permute(0, 2, 3, 1)
            ea = ((a.float() - ref).abs().max() / ref.abs().max()).item()
            eb = ((b.float() - ref).abs().max() / ref.abs().max()).item()
            assert a.shape == b.shape == ref.shape and ea < 3e-2 and eb < 3e-2, (ea, eb)
            assert ea < 1.5 * eb + 2e-3, (ea, eb)   # not worse than the library path against the fp32 reference


@pytest.mark.parametrize("N,H,W", [(3, 16, 16), (2, 32, 32), (2, 64, 64), (150, 16, 16), (37, 32, 32)])
def test_fused_input_norm_is_bit_identical_to_apply_then_conv(N, H, W, _variant):
    """ConvModule -> ConvModule: the second convolution applies the first module's GroupNorm + GELU to its operand on the way to
    the tensor core (transform warps rewrite each TMA slab in place, zero padding left alone).  The operand bits are those
    gn_apply would have written, the MMAs run in the same order: output and statistics are BIT-IDENTICAL to apply-then-conv."""
    from givepose_b200 import ops
    if not _variant:
        x = torch.zeros(1, 16, 16, 256, dtype=torch.bfloat16).cuda()
        assert not ops.conv3x3_fused_in_supported(x)      # one-CTA kernel: the fused input path is refused
        with pytest.raises(RuntimeError):
            ops.conv3x3_gn_bf16(x, torch.zeros(256, 9 * 256, dtype=torch.bfloat16).cuda(),
                                in_norm=(torch.zeros(64).cuda(), torch.ones(256).cuda(), torch.zeros(256).cuda()))
        return
    x, w0 = _case(N, H, W, 256, 11 + N)
    g = torch.Generator().manual_seed(N + H)
    w1 = (torch.randn(256, 256, 3, 3, generator=g) / 48).bfloat16().cuda()
    gamma = (torch.rand(256, generator=g) + 0.5).cuda()
    beta = (torch.randn(256, generator=g) * 0.3).cuda()
    y0, s0 = ops.conv3x3_gn_bf16(x, ops.pack_conv3x3_weight(w0))
    assert ops.conv3x3_fused_in_supported(y0)
    z = ops.groupnorm_apply(y0, s0, gamma, beta, 32, 1e-5, "gelu")
    ref, sref = ops.conv3x3_gn_bf16(z, ops.pack_conv3x3_weight(w1))
    got, sgot = ops.conv3x3_gn_bf16(y0, ops.pack_conv3x3_weight(w1), in_norm=(s0, gamma, beta))
    torch.cuda.synchronize()
    assert torch.equal(got, ref), ((got.float() - ref.float()).abs().max().item(), int((got != ref).sum()))
    assert torch.equal(sgot, sref)
    # and against the fp32 composition of the same ops
    want = F.conv2d(F.gelu(F.group_norm(y0.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)), w1.float(), None, 1, 1).permute(0, 2, 3, 1)
    assert ((got.float() - want).abs().max() / want.abs().max()).item() < 2e-2
