"""tcgen05 dense layer (``gp_linear_bf16``: TMA operand loads, TMEM accumulator, bias + activation epilogue) vs a plain
PyTorch fp32 reference of the same op on the same bf16-rounded operands.  Tolerance: the result is rounded to bf16 once
(2^-8 relative) on top of fp32 accumulation in a different order -> 1e-2 relative to the output's max magnitude is a loose
bar, measured ~4e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


SHAPES = [(128, 128, 64), (256, 256, 256), (677, 108, 256), (300, 2048, 1024), (128, 136, 8192), (5, 8, 72), (1000, 256, 1024),
          (4096, 2048, 512), (1300, 1032, 1088), (257, 264, 72)]


@pytest.fixture(params=[0, 1], ids=["one_cta", "cta_pair"])
def pair_mode(request):
    """0: always one CTA per tile; 1: the tcgen05 cta_group::2 CTA-pair kernel whenever the shape is legal for it."""
    from givepose_b200._lib import lib
    old = lib.gp_linear_set_pair(request.param)
    yield request.param
    lib.gp_linear_set_pair(old)


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("act", ["none", "lrelu", "relu"])
def test_linear_bf16_matches_torch(M, N, K, act, pair_mode):
    from givepose_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g)
    ref = F.linear(x.float(), w.float(), b)
    ref = F.leaky_relu(ref, 0.1) if act == "lrelu" else F.relu(ref) if act == "relu" else ref
    got = ops.linear_bf16(x.cuda(), w.cuda(), b.cuda(), act, 0.1)
    torch.cuda.synchronize()
    assert got.shape == (M, N) and got.dtype == torch.bfloat16
    assert rel(got.float(), ref) < 1e-2, rel(got.float(), ref)


def test_linear_bf16_batched_leading_dims_and_no_bias():
    from givepose_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 17, 256, generator=g).bfloat16()
    w = (torch.randn(72, 256, generator=g) / 16).bfloat16()
    got = ops.linear_bf16(x.cuda(), w.cuda())
    assert got.shape == (3, 17, 72) and rel(got.float(), F.linear(x.float(), w.float())) < 1e-2


def test_linear_bf16_refuses_cpu_and_bad_k():
    from givepose_b200 import ops
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        ops.linear_bf16(torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(8, 64, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError, match="unsupported shapes"):
        ops.linear_bf16(torch.zeros(4, 60, dtype=torch.bfloat16).cuda(), torch.zeros(8, 60, dtype=torch.bfloat16).cuda())


@pytest.mark.parametrize("N", [1, 3, 40])
def test_stem_implicit_gemm_matches_cudnn(N):
    """gp_stem_s2d_gemm (tcgen05 implicit GEMM, TMA im2col through an overlapping-row tensor map, ring of four input rows) against
    the same 4x4 convolution through cuDNN on the same packed operand: bf16 outputs, fp32 accumulation on both sides."""
    import torch.nn.functional as F
    from givepose_b200 import ops
    g = torch.Generator().manual_seed(5)
    img = torch.randn(N, 3, 256, 256, generator=g).cuda()
    packed = ops.stem_s2d_pack(img, torch.bfloat16)                                  # (N, 131, 131, 16)
    w = (torch.randn(64, 16, 4, 4, generator=g) * 0.1).bfloat16().cuda().contiguous(memory_format=torch.channels_last)
    b = torch.randn(64, generator=g).cuda()
    got = ops.stem_s2d_gemm(packed, w.permute(0, 2, 3, 1).reshape(64, 256), b)
    want = F.relu(F.conv2d(packed.permute(0, 3, 1, 2).float(), w.float(), b)).permute(0, 2, 3, 1)
    assert got.shape == (N, 128, 128, 64) and got.dtype == torch.bfloat16
    err = ((got.float() - want).abs().max() / want.abs().max()).item()
    assert err < 1e-2, err
    # every output row / image is written exactly once: no stale memory
    assert torch.isfinite(got.float()).all()
    # ReLU + MaxPool2d(3, 2, 1) folded into the epilogue: equals pooling the kernel's own un-pooled output exactly
    pooled = ops.stem_s2d_gemm(packed, w.permute(0, 2, 3, 1).reshape(64, 256), b, pool=True)
    ref = F.max_pool2d(got.permute(0, 3, 1, 2).float(), 3, 2, 1).permute(0, 2, 3, 1)
    assert pooled.shape == (N, 64, 64, 64) and torch.equal(pooled.float(), ref)
