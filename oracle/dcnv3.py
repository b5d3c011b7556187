"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DCNv3 core (see ``oracle/__init__.py``).

Two independent restatements of the reference:

* ``forward`` / ``backward`` / ``index``: ctypes calls into ``oracle/_build/libgp_oracle.so``
  (``dcnv3_oracle.c``), the C restatement of the reference *CUDA* arithmetic
  (``network/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:216-282, :386-487``) -- including the flat
  ``(q*G+g)*P`` addressing of ``offset``/``mask`` that makes the stride-2 in-model calls read only the
  first ``N*Ho*Wo`` rows of the full-resolution tensors (``cuda/dcnv3_cuda.cu:59-83``).
* ``dcnv3_core_torch``: restatement of the reference's *PyTorch* path ``dcnv3_core_pytorch``
  (``network/ops_dcnv3/functions/dcnv3_func.py:172-220``) on top of ``F.grid_sample``; together with
  ``flat_slice`` it covers stride 2 as well.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgp_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (``make -C oracle``).  Building the checker is not using it."""
    src_mtime = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("dcnv3_oracle.c", "dcnv3_oracle_impl.h"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_mtime:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def out_size(size: int, k: int, s: int, p: int, d: int) -> int:
    """``dcnv3_cuda.cu:40-45``."""
    return (size + 2 * p - (d * (k - 1) + 1)) // s + 1


def _ptr(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


def _suffix(dtype):
    if dtype == torch.float32:
        return "f32", ctypes.c_float
    if dtype == torch.float64:
        return "f64", ctypes.c_double
    raise TypeError(f"oracle computes in float32 or float64, got {dtype}")


def _geom(input, kh, kw, sh, sw, ph, pw, dh, dw):
    N, H, W, C = input.shape
    return N, H, W, C, out_size(H, kh, sh, ph, dh), out_size(W, kw, sw, pw, dw)


def forward(input, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, group, group_channels, offset_scale,
            remove_center=0):
    """C-oracle forward.  ``offset``/``mask`` are read through their flat prefix, like the CUDA kernel."""
    lib = _load()
    sfx, creal = _suffix(input.dtype)
    N, H, W, C, Ho, Wo = _geom(input, kh, kw, sh, sw, ph, pw, dh, dw)
    assert C == group * group_channels
    P = kh * kw - int(remove_center)
    assert offset.numel() >= N * Ho * Wo * group * P * 2 and mask.numel() >= N * Ho * Wo * group * P
    input, offset, mask = input.contiguous(), offset.contiguous().to(input.dtype), mask.contiguous().to(input.dtype)
    out = torch.empty((N, Ho, Wo, C), dtype=input.dtype)
    getattr(lib, f"gpo_dcnv3_forward_{sfx}")(
        _ptr(input), _ptr(offset), _ptr(mask), _ptr(out), N, H, W, group, group_channels, kh, kw, sh, sw, ph, pw,
        dh, dw, creal(offset_scale), int(remove_center), Ho, Wo)
    return out


def backward(input, offset, mask, grad_output, kh, kw, sh, sw, ph, pw, dh, dw, group, group_channels,
             offset_scale, remove_center=0):
    """C-oracle backward: returns (grad_input, grad_offset, grad_mask) with the FULL shapes of
    input/offset/mask (rows beyond the flat prefix stay zero, ``dcnv3_cuda.cu:131-133``)."""
    lib = _load()
    sfx, creal = _suffix(input.dtype)
    N, H, W, C, Ho, Wo = _geom(input, kh, kw, sh, sw, ph, pw, dh, dw)
    input, offset, mask = input.contiguous(), offset.contiguous().to(input.dtype), mask.contiguous().to(input.dtype)
    grad_output = grad_output.contiguous().to(input.dtype)
    assert grad_output.shape == (N, Ho, Wo, C)
    gi, go, gm = torch.zeros_like(input), torch.zeros_like(offset), torch.zeros_like(mask)
    getattr(lib, f"gpo_dcnv3_backward_{sfx}")(
        _ptr(input), _ptr(offset), _ptr(mask), _ptr(grad_output), _ptr(gi), _ptr(go), _ptr(gm), N, H, W, group,
        group_channels, kh, kw, sh, sw, ph, pw, dh, dw, creal(offset_scale), int(remove_center), Ho, Wo)
    return gi, go, gm


def index(offset, N, H, W, kh, kw, sh, sw, ph, pw, dh, dw, group, offset_scale, remove_center=0):
    """Integer part of the contract: per (q, g, p) the floor()ed corner (h_low, w_low) and the
    in-range / four per-corner bounds flags (bit0 in_range, bit1..4 corners 1..4)."""
    lib = _load()
    sfx, creal = _suffix(offset.dtype)
    Ho, Wo = out_size(H, kh, sh, ph, dh), out_size(W, kw, sw, pw, dw)
    P = kh * kw - int(remove_center)
    n = N * Ho * Wo * group * P
    offset = offset.contiguous()
    assert offset.numel() >= 2 * n
    hw = torch.zeros((n, 2), dtype=torch.int32)
    flags = torch.zeros((n,), dtype=torch.uint8)
    getattr(lib, f"gpo_dcnv3_index_{sfx}")(
        _ptr(offset), _ptr(hw), _ptr(flags), N, H, W, group, kh, kw, sh, sw, ph, pw, dh, dw,
        creal(offset_scale), int(remove_center), Ho, Wo)
    return hw, flags


class CoreFunction(torch.autograd.Function):
    """Autograd wrapper around the C oracle (forward :216-282, backward :386-487 of dcnv3_im2col_cuda.cuh): the CPU twin of
    the reference's ``DCNv3Function`` (functions/dcnv3_func.py:22-77), same flat offset / mask addressing."""

    @staticmethod
    def forward(ctx, input, offset, mask, geom, remove_center):
        ctx.geom, ctx.rc = geom, remove_center
        ctx.save_for_backward(input, offset, mask)
        return forward(input, offset, mask, *geom, remove_center)

    @staticmethod
    def backward(ctx, grad_output):
        input, offset, mask = ctx.saved_tensors
        gi, go, gm = backward(input, offset, mask, grad_output.contiguous(), *ctx.geom, ctx.rc)
        return gi, go, gm, None, None


def flat_slice(t: torch.Tensor, N: int, Ho: int, Wo: int) -> torch.Tensor:
    """The stride-2 adapter of SURVEY.md section 0.1: the CUDA kernel reads ``offset``/``mask`` as a flat
    ``[N*Ho*Wo, G*P*(2)]`` matrix (``dcnv3_im2col_cuda.cuh:229,243-244``), i.e. only the first
    ``N*Ho*Wo`` rows of a full-resolution ``(N, H, W, G*P*(2))`` tensor."""
    last = t.shape[-1]
    return t.reshape(-1)[: N * Ho * Wo * last].view(N, Ho, Wo, last)


def dcnv3_core_torch(input, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, group, group_channels,
                     offset_scale, remove_center=0):
    """grid_sample restatement of ``dcnv3_core_pytorch`` (``functions/dcnv3_func.py:172-220``).

    Same quantities in the same float order: normalised reference point (``:109-137``) + normalised
    kernel grid * scale (``:140-162``) + offset * scale / (W_in, H_in) (``:196-200``); ``2x-1``;
    ``F.grid_sample(bilinear, zeros, align_corners=False)`` per group (``:205-212``); mask-weighted sum
    over the P points (``:215-218``).  ``H_out, W_out`` come from ``offset.shape`` (``:187``).
    """
    if remove_center and (kh % 2 == 0 or kw % 2 == 0 or kw != kh):
        raise ValueError("remove_center is only compatible with square odd kernel size.")
    # :183-185 -- the reference passes (pad_h, pad_h) for the W axis and (pad_w, pad_w) for the H axis.
    x = F.pad(input, [0, 0, ph, ph, pw, pw])
    N, Hp, Wp, C = x.shape
    _, Ho, Wo, _ = offset.shape
    dev, f32 = x.device, torch.float32
    P = kh * kw - int(remove_center)

    # reference points, pixel centres of the padded image (:113-137); linspace in float32 like the reference
    y0, x0 = (dh * (kh - 1)) // 2 + 0.5, (dw * (kw - 1)) // 2 + 0.5
    Hq = (Hp - (dh * (kh - 1) + 1)) // sh + 1
    Wq = (Wp - (dw * (kw - 1) + 1)) // sw + 1
    ry = torch.linspace(y0, y0 + (Hq - 1) * sh, Hq, dtype=f32, device=dev) / Hp
    rx = torch.linspace(x0, x0 + (Wq - 1) * sw, Wq, dtype=f32, device=dev) / Wp
    ref = torch.stack((rx[None, :].expand(Hq, Wq), ry[:, None].expand(Hq, Wq)), -1)          # (Hq, Wq, 2) = (x, y)

    # kernel grid, x (width) is the slow axis: p = i*kh + j (:143-160)
    gx = torch.linspace(-((dw * (kw - 1)) // 2), -((dw * (kw - 1)) // 2) + (kw - 1) * dw, kw, dtype=f32, device=dev)
    gy = torch.linspace(-((dh * (kh - 1)) // 2), -((dh * (kh - 1)) // 2) + (kh - 1) * dh, kh, dtype=f32, device=dev)
    grid = torch.stack((gx[:, None].expand(kw, kh) / Wp, gy[None, :].expand(kw, kh) / Hp), -1).reshape(kw * kh, 2)
    if remove_center:                                                                        # :165-170
        keep = [p for p in range(kw * kh) if p != (kw * kh - 1) // 2]
        grid = grid[keep]

    loc = ref[None, :, :, None, None, :] + (grid * offset_scale)[None, None, None, None, :, :]  # (1,Hq,Wq,1,P,2)
    norm = torch.tensor([Wp, Hp], device=dev)
    loc = loc + (offset.view(N, Ho, Wo, group, P, 2) * offset_scale / norm)                   # :196-200
    g = 2 * loc - 1                                                                           # :203
    g = g.permute(0, 3, 1, 2, 4, 5).reshape(N * group, Ho * Wo, P, 2)
    xg = x.view(N, Hp * Wp, group, group_channels).permute(0, 2, 3, 1).reshape(N * group, group_channels, Hp, Wp)
    samp = F.grid_sample(xg, g, mode="bilinear", padding_mode="zeros", align_corners=False)   # (N*G, gc, Ho*Wo, P)
    m = mask.view(N, Ho * Wo, group, P).permute(0, 2, 1, 3).reshape(N * group, 1, Ho * Wo, P)
    out = (samp * m).sum(-1).view(N, group * group_channels, Ho * Wo)
    return out.transpose(1, 2).reshape(N, Ho, Wo, -1).contiguous()
