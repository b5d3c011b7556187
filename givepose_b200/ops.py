"""Host wrappers of the fused PoseNet glue kernels (``include/givepose_b200.h``, "PoseNet forward" block).

Each wrapper validates like the DCNv3 boundary does (CUDA, contiguous, supported dtype), allocates its output on the
input's device and enqueues on the current torch stream.  No CPU / PyTorch fallback: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import check, lib

_DTYPES = {torch.float32: _lib.GP_F32, torch.bfloat16: _lib.GP_BF16, torch.float16: _lib.GP_F16}
ACT = {"none": 0, "relu": 1, "gelu": 2}


def _vp(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _need_cuda(name, t, dtype=None):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: Not implemented on the CPU")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} tensor has to be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")


def dwconv3x3_ln_gelu(x, w_t, bias, ln_w, ln_b, rows=None, eps=1e-6):
    """``GELU(LayerNorm(DWConv3x3(x)))`` of the DCNv3 module (``modules/dcnv3.py:269-283,329``) for the first ``rows``
    pixels of channel-last ``x`` (N,H,W,C); returns ``(rows, C)``.  ``w_t`` is the depthwise weight as ``[9, C]`` fp32."""
    _need_cuda("input", x)
    for n, t in (("dw weight", w_t), ("dw bias", bias), ("ln weight", ln_w), ("ln bias", ln_b)):
        _need_cuda(n, t, torch.float32)
    dt = _DTYPES.get(x.dtype)
    if dt is None or x.dim() != 4:
        raise RuntimeError(f"dwconv3x3_ln_gelu: unsupported input {x.dtype} {tuple(x.shape)}")
    N, H, W, C = x.shape
    rows = N * H * W if rows is None else int(rows)
    out = torch.empty((rows, C), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gp_dwconv3x3_ln_gelu(_vp(x), _vp(w_t), _vp(bias), _vp(ln_w), _vp(ln_b), _vp(out), N, H, W, C, rows,
                                       float(eps), dt, _stream(x)), "dwconv3x3_ln_gelu")
    return out


def small_k_linear(x, w_t, bias):
    """``x (..., K) @ w_t (K, C) + bias`` for K = 3 input channels (the composed ``input_proj(conv1x1(x))`` of the first
    MAPEncoder layer); returns ``(..., C)`` in ``x.dtype``."""
    _need_cuda("input", x)
    _need_cuda("weight", w_t, torch.float32)
    _need_cuda("bias", bias, torch.float32)
    dt = _DTYPES.get(x.dtype)
    K, C = w_t.shape
    if dt is None or x.shape[-1] != K or bias.shape != (C,):
        raise RuntimeError(f"small_k_linear: unsupported input {x.dtype} {tuple(x.shape)} / weight {tuple(w_t.shape)}")
    out = torch.empty((*x.shape[:-1], C), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gp_small_k_linear(_vp(x), _vp(w_t), _vp(bias), _vp(out), x.numel() // K, K, C, dt, _stream(x)), "small_k_linear")
    return out


def smallk_dwconv3x3_ln_gelu(x, w_eff, bias, ln_w, ln_b, rows=None, eps=1e-6):
    """``GELU(LayerNorm(DWConv3x3(Conv1x1_{K->C}(x))))`` for the first ``rows`` pixels of channel-last ``x`` (N,H,W,K), K = 3,
    without materialising the C-channel convolution output.  ``w_eff`` (9, K+1, C) fp32, see ``include/givepose_b200.h``."""
    _need_cuda("input", x)
    for n, t in (("w_eff", w_eff), ("dw bias", bias), ("ln weight", ln_w), ("ln bias", ln_b)):
        _need_cuda(n, t, torch.float32)
    dt = _DTYPES.get(x.dtype)
    if dt is None or x.dim() != 4 or w_eff.dim() != 3 or w_eff.shape[1] != x.shape[-1] + 1:
        raise RuntimeError(f"smallk_dwconv3x3_ln_gelu: unsupported input {x.dtype} {tuple(x.shape)}")
    N, H, W, K = x.shape
    C = w_eff.shape[-1]
    rows = N * H * W if rows is None else int(rows)
    out = torch.empty((rows, C), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gp_smallk_dwconv3x3_ln_gelu(_vp(x), _vp(w_eff), _vp(bias), _vp(ln_w), _vp(ln_b), _vp(out), N, H, W, K, C, rows,
                                              float(eps), dt, _stream(x)), "smallk_dwconv3x3_ln_gelu")
    return out


def dcnv3_smallk_fused(x, offset, mask_logits, w2, bias, geom, remove_center=0):
    """The whole first-layer DCNv3 module (``include/givepose_b200.h``, ``gp_dcnv3_smallk_fused``): ``x`` (N,H,W,3) channel-last,
    ``offset`` / ``mask_logits`` flat-prefix buffers, ``w2`` (G*4, 256) / ``bias`` (256,) fp32 composed weights, ``geom`` the
    11-tuple ``(kh, kw, sh, sw, ph, pw, dh, dw, group, group_channels, offset_scale)``; returns (N,Ho,Wo,256)."""
    from ._lib import DCNv3Desc
    _need_cuda("input", x)
    _need_cuda("offset", offset, x.dtype)
    _need_cuda("mask", mask_logits, x.dtype)
    _need_cuda("w2", w2, torch.float32)
    _need_cuda("bias", bias, torch.float32)
    dt = _DTYPES.get(x.dtype)
    if dt is None or x.dim() != 4:
        raise RuntimeError(f"dcnv3_smallk_fused: unsupported input {x.dtype} {tuple(x.shape)}")
    kh, kw, sh, sw, ph, pw, dh, dw, G, gc, scale = geom
    N, H, W, K = x.shape
    Ho = lib.gp_dcnv3_out_size(H, kh, sh, ph, dh)
    Wo = lib.gp_dcnv3_out_size(W, kw, sw, pw, dw)
    C = bias.numel()
    if tuple(w2.shape) != (G * (K + 1), C):
        raise RuntimeError(f"dcnv3_smallk_fused: w2 must be ({G * (K + 1)}, {C}), got {tuple(w2.shape)}")
    d = DCNv3Desc(N, H, W, G, gc, kh, kw, sh, sw, ph, pw, dh, dw, int(bool(remove_center)), Ho, Wo, float(scale))
    out = torch.empty((N, Ho, Wo, C), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gp_dcnv3_smallk_fused(_vp(x), _vp(offset), _vp(mask_logits), _vp(w2), _vp(bias), _vp(out), offset.numel(), mask_logits.numel(),
                                        ctypes.byref(d), K, C, dt, _stream(x)), "dcnv3_smallk_fused")
    return out


def _nhwc(name, x):
    _need_cuda(name, x)
    dt = _DTYPES.get(x.dtype)
    if dt is None or x.dim() != 4:
        raise RuntimeError(f"{name}: unsupported input {x.dtype} {tuple(x.shape)}")
    return dt


def groupnorm_act(x, gamma, beta, groups=32, eps=1e-5, act="none", upsample2x=False, return_stats=False):
    """``act(GroupNorm(x))`` on channel-last ``x`` (N,H,W,C); ``upsample2x`` appends the align_corners=True bilinear x2
    upsampling of ``TopDownXyzHead`` (a second pass: see posenet_kernels.cuh).  ``gamma`` / ``beta`` are fp32."""
    dt = _nhwc("input", x)
    _need_cuda("gn weight", gamma, torch.float32)
    _need_cuda("gn bias", beta, torch.float32)
    N, H, W, C = x.shape
    y = torch.empty_like(x)
    stats = torch.empty(lib.gp_groupnorm_workspace_floats(N, H, W, int(groups)), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gp_groupnorm_act(_vp(x), _vp(y), _vp(stats), stats.numel(), _vp(gamma), _vp(beta), N, H, W, C, int(groups), float(eps),
                                   ACT[act], dt, _stream(x)), "groupnorm_act")
    y = upsample_bilinear2x(y) if upsample2x else y
    return (y, stats[:N * int(groups) * 2]) if return_stats else y


def groupnorm_act_backward(x, dy, stats, gamma, beta, groups=32, act="none"):
    """Gradients of ``groupnorm_act`` (without the upsampling): ``(dx, dgamma, dbeta)``; ``stats`` is what the forward returned
    with ``return_stats=True``.  ``dx`` has ``x``'s dtype, ``dgamma`` / ``dbeta`` are fp32."""
    dt = _nhwc("input", x)
    _need_cuda("grad_output", dy, x.dtype)
    for n, t in (("stats", stats), ("gn weight", gamma), ("gn bias", beta)):
        _need_cuda(n, t, torch.float32)
    N, H, W, C = x.shape
    if dy.shape != x.shape or stats.numel() < N * int(groups) * 2:
        raise RuntimeError("groupnorm_act_backward: inconsistent shapes")
    dx = torch.empty_like(x)
    dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
    ws = torch.empty(lib.gp_groupnorm_backward_workspace_floats(N, H, W, C, int(groups)), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gp_groupnorm_act_backward(_vp(x), _vp(dy), _vp(stats), _vp(gamma), _vp(beta), _vp(dx), _vp(dgamma), _vp(dbeta), _vp(ws),
                                            ws.numel(), N, H, W, C, int(groups), ACT[act], dt, _stream(x)), "groupnorm_act_backward")
    return dx, dgamma, dbeta


class GroupNormAct(torch.autograd.Function):
    """``act(GroupNorm(x))`` on channel-last activations as ONE autograd node for the training step: forward = the inference
    kernels, backward = ``groupnorm_act_backward``.  Only ``x`` (storage dtype) and the (mean, rstd) pairs are saved; torch's
    eager path keeps the fp32 GroupNorm output and the activation input alive and, under autocast, wraps both in dtype copies."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, act):
        xc = x.contiguous()
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        y, stats = groupnorm_act(xc, g32, b32, groups, eps, act, return_stats=True)
        ctx.save_for_backward(xc, stats, g32, b32)
        ctx.groups, ctx.act, ctx.pdtype = groups, act, gamma.dtype
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, stats, g32, b32 = ctx.saved_tensors
        dx, dg, db = groupnorm_act_backward(x, dy.to(x.dtype).contiguous(), stats, g32, b32, ctx.groups, ctx.act)
        return dx, dg.to(ctx.pdtype), db.to(ctx.pdtype), None, None, None


def groupnorm_act_conv1x1(x, gamma, beta, weight, bias, groups=32, eps=1e-5, act="gelu"):
    """``Conv1x1(act(GroupNorm(x))) + bias`` on channel-last ``x`` (N,H,W,256) -> (N,H,W,3): the decoder's last GN/GELU
    fused with its ``out_layer`` (``xyz_head.py:349-366``).  ``weight`` (3,256) / ``bias`` (3,) / ``gamma`` / ``beta`` fp32."""
    dt = _nhwc("input", x)
    for n, t in (("gn weight", gamma), ("gn bias", beta), ("out_layer weight", weight), ("out_layer bias", bias)):
        _need_cuda(n, t, torch.float32)
    N, H, W, C = x.shape
    OC = weight.shape[0]
    if weight.shape != (OC, C) or bias.shape != (OC,):
        raise RuntimeError("groupnorm_act_conv1x1: inconsistent weight / bias shapes")
    y = torch.empty((N, H, W, OC), dtype=x.dtype, device=x.device)
    stats = torch.empty(lib.gp_groupnorm_workspace_floats(N, H, W, int(groups)), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gp_groupnorm_act_conv1x1(_vp(x), _vp(y), _vp(stats), stats.numel(), _vp(gamma), _vp(beta), _vp(weight), _vp(bias), N, H, W, C,
                                           int(groups), float(eps), ACT[act], OC, dt, _stream(x)), "groupnorm_act_conv1x1")
    return y


def pack_conv3x3_weight(weight):
    """``Conv2d(Cin, 256, 3)`` weight (Cout, Cin, 3, 3) -> the K-major bf16 matrix (Cout, 9*Cin), K = (ky, kx, cin), that
    ``conv3x3_gn_bf16`` reads."""
    return weight.detach().permute(0, 2, 3, 1).reshape(weight.shape[0], -1).contiguous().to(torch.bfloat16)


def conv3x3_gn_supported(x, cout):
    """Shapes the tcgen05 implicit-GEMM convolution takes: bf16 channel-last, Cout 256, Cin % 64 == 0, W in {8, 16, 32, 64}, H*W % 256 == 0."""
    if x.dim() != 4 or x.dtype != torch.bfloat16 or not x.is_cuda:
        return False
    N, H, W, C = x.shape
    return cout == 256 and C % 64 == 0 and 8 <= W <= 64 and 128 % W == 0 and (H * W) % 256 == 0 and H % (256 // W) == 0


def conv3x3_fused_in_supported(x):
    """Can ``conv3x3_gn_bf16(..., in_norm=...)`` apply the producer layer's GroupNorm(32) + GELU inside the convolution?  (CTA-pair
    kernel, input channels a multiple of 256: a 16-byte operand slot then holds channels of one group.)"""
    return x.dim() == 4 and x.shape[-1] % 256 == 0 and lib.gp_conv3x3_gn_slabs(x.shape[1], x.shape[2]) == 8 * (x.shape[1] * x.shape[2] // 256)


def conv3x3_gn_bf16(x, w_packed, groups=32, eps=1e-5, stats=True, in_norm=None):
    """``Conv2d(Cin, 256, 3, padding=1, bias=False)`` on channel-last bf16 ``x`` (N,H,W,Cin) as a hand-written tcgen05 implicit
    GEMM (``conv3x3_tc.cu``; ``xyz_head.py:195-366`` / ``conv_module.py:57-234``).  Returns ``(y, stats)``: ``y`` (N,H,W,256)
    bf16 and, if ``stats``, the GroupNorm(32) ``(mean, rstd)`` pairs [N*32*2] of the fp32 accumulators (for ``groupnorm_apply``).

    ``in_norm = (in_stats, gamma, beta)``: ``x`` is the RAW output of the previous ConvModule's convolution and that module's
    GroupNorm + GELU is applied to the operand inside this kernel (bit-identical to ``groupnorm_apply(..., "gelu")`` followed by
    the plain call, without the pass over the activation); see ``conv3x3_fused_in_supported``."""
    _need_cuda("input", x, torch.bfloat16)
    _need_cuda("weight", w_packed, torch.bfloat16)
    N, H, W, C = x.shape
    Cout = w_packed.shape[0]
    if tuple(w_packed.shape) != (Cout, 9 * C) or not conv3x3_gn_supported(x, Cout) or int(groups) != 32:
        raise RuntimeError(f"conv3x3_gn_bf16: unsupported shapes {tuple(x.shape)} x {tuple(w_packed.shape)} (groups {groups})")
    y = torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device=x.device)
    slabs = lib.gp_conv3x3_gn_slabs(H, W)          # 4 (one CTA per tile) or 8 (CTA pair) partial rows per 256-pixel tile
    room = max(slabs, 8 * (H * W // 256))          # sized for either kernel variant, whatever gp_conv3x3_set_pair says later
    ws = torch.empty(N * 32 * 2 * (1 + room), dtype=torch.float32, device=x.device) if stats else None
    partial = ws[N * 32 * 2:] if stats else None
    with torch.cuda.device(x.device):
        if in_norm is not None:
            in_stats, gamma, beta = in_norm
            for n, t in (("in_stats", in_stats), ("gn weight", gamma), ("gn bias", beta)):
                _need_cuda(n, t, torch.float32)
            if in_stats.numel() < N * 64 or gamma.numel() != C or beta.numel() != C:
                raise RuntimeError("conv3x3_gn_bf16: in_norm does not match the input")
            check(lib.gp_conv3x3_gn_bf16_fused_in(_vp(x), _vp(in_stats), _vp(gamma), _vp(beta), _vp(w_packed), _vp(y),
                                                  _vp(partial) if stats else None, N, H, W, C, Cout, _stream(x)), "conv3x3_gn_bf16_fused_in")
        else:
            check(lib.gp_conv3x3_gn_bf16(_vp(x), _vp(w_packed), _vp(y), _vp(partial) if stats else None, N, H, W, C, Cout, _stream(x)),
                  "conv3x3_gn_bf16")
        if stats:
            check(lib.gp_groupnorm_finalize(_vp(partial), _vp(ws), N, 32, slabs, H * W * (Cout // 32), float(eps), _stream(x)),
                  "groupnorm_finalize")
    return y, (ws[:N * 32 * 2] if stats else None)


def groupnorm_apply(x, stats, gamma, beta, groups=32, eps=1e-5, act="none", upsample2x=False):
    """Apply pass of ``groupnorm_act`` with precomputed ``(mean, rstd)`` pairs (``conv3x3_gn_bf16``)."""
    dt = _nhwc("input", x)
    for n, t in (("stats", stats), ("gn weight", gamma), ("gn bias", beta)):
        _need_cuda(n, t, torch.float32)
    N, H, W, C = x.shape
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.gp_groupnorm_apply(_vp(x), _vp(y), _vp(stats), stats.numel(), _vp(gamma), _vp(beta), N, H, W, C, int(groups), float(eps),
                                     ACT[act], dt, _stream(x)), "groupnorm_apply")
    return upsample_bilinear2x(y) if upsample2x else y


def groupnorm_apply_conv1x1(x, stats, gamma, beta, weight, bias, groups=32, eps=1e-5, act="gelu"):
    """``groupnorm_act_conv1x1`` with precomputed ``(mean, rstd)`` pairs."""
    dt = _nhwc("input", x)
    for n, t in (("stats", stats), ("gn weight", gamma), ("gn bias", beta), ("out_layer weight", weight), ("out_layer bias", bias)):
        _need_cuda(n, t, torch.float32)
    N, H, W, C = x.shape
    OC = weight.shape[0]
    if weight.shape != (OC, C) or bias.shape != (OC,):
        raise RuntimeError("groupnorm_apply_conv1x1: inconsistent weight / bias shapes")
    y = torch.empty((N, H, W, OC), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gp_groupnorm_apply_conv1x1(_vp(x), _vp(y), _vp(stats), stats.numel(), _vp(gamma), _vp(beta), _vp(weight), _vp(bias), N, H, W,
                                             C, int(groups), float(eps), ACT[act], OC, dt, _stream(x)), "groupnorm_apply_conv1x1")
    return y


LIN_ACT = {"none": 0, "lrelu": 1, "relu": 2}


def linear_bf16(x, weight, bias=None, act="none", slope=0.1):
    """``act(x @ weight.T + bias)`` on the tcgen05 tensor cores: ``x`` (..., K) and ``weight`` (N, K) bf16, ``bias`` fp32 (N,) or
    None; returns (..., N) bf16.  ``act``: 'none' | 'lrelu' (``slope``) | 'relu'."""
    _need_cuda("input", x, torch.bfloat16)
    _need_cuda("weight", weight, torch.bfloat16)
    N, K = weight.shape
    if x.shape[-1] != K or K % 8:
        raise RuntimeError(f"linear_bf16: unsupported shapes {tuple(x.shape)} x {tuple(weight.shape)}")
    if bias is None:
        bias = torch.zeros(N, dtype=torch.float32, device=x.device)
    _need_cuda("bias", bias, torch.float32)
    M = x.numel() // K
    y = torch.empty((*x.shape[:-1], N), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gp_linear_bf16(_vp(x), _vp(weight), _vp(bias), _vp(y), M, N, K, LIN_ACT[act], float(slope), _stream(x)), "linear_bf16")
    return y


def mhsa_tokens(qkv, num_heads):
    """Self-attention over the 64 patch tokens of ``MAPTransformerEncoer``: ``qkv`` (B, 64, 3*C) straight from the qkv
    Linear (timm layout ``(B, N, 3, heads, C/heads)``) -> (B, 64, C)."""
    _need_cuda("qkv", qkv)
    dt = _DTYPES.get(qkv.dtype)
    if dt is None or qkv.dim() != 3 or qkv.shape[2] % (3 * num_heads):
        raise RuntimeError(f"mhsa_tokens: unsupported input {qkv.dtype} {tuple(qkv.shape)}")
    B, NT, C3 = qkv.shape
    C = C3 // 3
    hd = C // num_heads
    out = torch.empty((B, NT, C), dtype=qkv.dtype, device=qkv.device)
    with torch.cuda.device(qkv.device):
        check(lib.gp_mhsa_tokens(_vp(qkv), _vp(out), B, NT, num_heads, hd, float(hd) ** -0.5, dt, _stream(qkv)), "mhsa_tokens")
    return out


def stem_s2d_pack(img, dtype):
    """fp32 NCHW RoI crops (N,3,H,W) -> (N, H/2+3, W/2+3, 16) channel-last ``dtype``: the 2x2 space-to-depth operand of a
    7x7/2 stem convolution run as a 4x4/1 convolution (see ``posenet._stem``)."""
    _need_cuda("roi_img", img, torch.float32)
    dt = _DTYPES.get(dtype)
    if dt is None or img.dim() != 4 or img.shape[1] != 3 or img.shape[2] % 2 or img.shape[3] % 2:
        raise RuntimeError(f"stem_s2d_pack: unsupported input {dtype} {tuple(img.shape)}")
    N, _, H, W = img.shape
    out = torch.empty((N, H // 2 + 3, W // 2 + 3, 16), dtype=dtype, device=img.device)
    with torch.cuda.device(img.device):
        check(lib.gp_stem_s2d_pack(_vp(img), _vp(out), N, H, W, dt, _stream(img)), "stem_s2d_pack")
    return out


def stem_s2d_gemm(packed, w2d, bias, pool=False):
    """relu(conv4x4/1(packed) + bias) on the tcgen05 tensor cores (``gp_stem_s2d_gemm``): ``packed`` (N,Hp,Wp,16) bf16 from
    ``stem_s2d_pack``, ``w2d`` (64, 256) bf16 tap-major, ``bias`` fp32 (64,) -> (N,Hp-3,Wp-3,64) bf16; ``pool=True`` applies
    the backbone's ``MaxPool2d(3, 2, 1)`` in the epilogue and returns the pooled (N,64,64,64) tensor.  Raises for unsupported
    shapes (the caller falls back to cuDNN)."""
    _need_cuda("packed", packed, torch.bfloat16)
    _need_cuda("weight", w2d, torch.bfloat16)
    _need_cuda("bias", bias, torch.float32)
    N, Hp, Wp, C = packed.shape
    if C != 16 or tuple(w2d.shape) != (64, 256) or bias.numel() != 64:
        raise RuntimeError("stem_s2d_gemm: unsupported shapes")
    Ho, Wo = Hp - 3, Wp - 3
    shape = (N, (Ho - 1) // 2 + 1, (Wo - 1) // 2 + 1, 64) if pool else (N, Ho, Wo, 64)
    y = torch.empty(shape, dtype=torch.bfloat16, device=packed.device)
    with torch.cuda.device(packed.device):
        check(lib.gp_stem_s2d_gemm(_vp(packed), _vp(w2d), _vp(bias), _vp(y), N, Hp, Wp, int(bool(pool)), _stream(packed)), "stem_s2d_gemm")
    return y


def upsample_bilinear2x(x):
    """``nn.UpsamplingBilinear2d(scale_factor=2)`` (align_corners=True) on channel-last ``x``."""
    dt = _nhwc("input", x)
    N, H, W, C = x.shape
    y = torch.empty((N, 2 * H, 2 * W, C), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gp_upsample_bilinear2x(_vp(x), _vp(y), N, H, W, C, dt, _stream(x)), "upsample_bilinear2x")
    return y


def upsample_bilinear2x_backward(dy):
    """Gradient of ``upsample_bilinear2x`` w.r.t. its input: ``dy`` (N,2H,2W,C) -> (N,H,W,C); a gather, no atomics."""
    dt = _nhwc("grad_output", dy)
    N, Ho, Wo, C = dy.shape
    if Ho % 2 or Wo % 2:
        raise RuntimeError("upsample_bilinear2x_backward: odd output size")
    dx = torch.empty((N, Ho // 2, Wo // 2, C), dtype=dy.dtype, device=dy.device)
    with torch.cuda.device(dy.device):
        check(lib.gp_upsample_bilinear2x_backward(_vp(dy), _vp(dx), N, Ho // 2, Wo // 2, C, dt, _stream(dy)), "upsample_bilinear2x_backward")
    return dx


class UpsampleBilinear2x(torch.autograd.Function):
    """``nn.UpsamplingBilinear2d(2)`` on channel-last activations as one autograd node (training step)."""

    @staticmethod
    def forward(ctx, x):
        return upsample_bilinear2x(x.contiguous())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        return upsample_bilinear2x_backward(dy.contiguous())


def bias_add_relu_(y, residual, bias):
    """In place ``y = relu(y + bias + residual)`` on channel-last ``y`` (..., C); ``bias`` fp32 (C,).  Returns ``y``."""
    _need_cuda("y", y)
    _need_cuda("residual", residual, y.dtype)
    _need_cuda("bias", bias, torch.float32)
    dt = _DTYPES.get(y.dtype)
    C = y.shape[-1]
    if dt is None or residual.shape != y.shape or bias.numel() != C:
        raise RuntimeError("bias_add_relu_: inconsistent shapes / dtype")
    with torch.cuda.device(y.device):
        check(lib.gp_bias_add_relu(_vp(y), _vp(residual), _vp(bias), y.numel() // C, C, dt, _stream(y)), "bias_add_relu")
    return y


def maxpool3x3s2(x, relu=False):
    """``MaxPool2d(3, 2, 1)`` on channel-last ``x`` (N,H,W,C); ``relu=True`` computes ``maxpool(relu(x))`` in the same pass."""
    dt = _nhwc("input", x)
    N, H, W, C = x.shape
    y = torch.empty((N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gp_maxpool3x3s2(_vp(x), _vp(y), N, H, W, C, int(bool(relu)), dt, _stream(x)), "maxpool3x3s2")
    return y


def pose_decode(rot6, t, cam_K, centers, whs, ratios, is_allo=True, z_calib=1.0):
    """rot6d -> R, back-projection and allo->ego on the device (replaces the host loop of
    ``pose_from_predictions_test``, ``pose_from_pred_centroid_z.py:139-157``).  Returns ``(rot (B,3,3), trans (B,3))``."""
    args = [("rot6", rot6), ("t", t), ("cam_K", cam_K), ("bbox_center", centers), ("roi_wh", whs), ("resize_ratio", ratios)]
    args = [(n, a.float().contiguous()) for n, a in args]
    for n, a in args:
        _need_cuda(n, a, torch.float32)
    rot6, t, cam_K, centers, whs, ratios = (a for _, a in args)
    B = rot6.shape[0]
    if rot6.shape != (B, 6) or t.shape != (B, 3) or centers.shape != (B, 2) or whs.shape != (B, 2) or ratios.numel() != B:
        raise RuntimeError("pose_decode: inconsistent shapes")
    if cam_K.numel() not in (9, 9 * B):
        raise RuntimeError("pose_decode: cam_K must be (3,3), (1,3,3) or (B,3,3)")
    batched = cam_K.dim() == 3 and cam_K.shape[0] == B   # (1,3,3) with B == 1 reads the same 9 floats either way
    rot = torch.empty((B, 3, 3), dtype=torch.float32, device=rot6.device)
    trans = torch.empty((B, 3), dtype=torch.float32, device=rot6.device)
    with torch.cuda.device(rot6.device):
        check(lib.gp_pose_decode(_vp(rot6), _vp(t), _vp(cam_K), int(batched), _vp(centers), _vp(whs), _vp(ratios),
                                 _vp(rot), _vp(trans), B, int(bool(is_allo)), float(z_calib), _stream(rot6)), "pose_decode")
    return rot, trans
