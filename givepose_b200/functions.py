"""Operator boundary: ``dcnv3_forward`` / ``dcnv3_backward`` / ``DCNv3Function``.

Mirrors the reference's compiled module ``DCNv3`` (``network/ops_dcnv3/src/vision.cpp:14-17``,
``src/dcnv3.h:20-59``, ``src/cuda/dcnv3_cuda.cu:21-174``) and its autograd wrapper
(``network/ops_dcnv3/functions/dcnv3_func.py:22-106``): same names, argument order, output ownership
(the op allocates outputs on the inputs' device / dtype; grads have the full shapes of
input/offset/mask), and error behaviour (RuntimeError for CPU / non-contiguous tensors, bad batch vs
im2col_step, C != group*group_channels).  Differences, all documented in DESIGN.md: bf16 is accepted;
kernel-launch errors are raised instead of printed; the im2col_step chunk loop is gone (the flat
offset/mask addressing makes it a no-op, the argument is still validated).
"""
from __future__ import annotations

import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import DCNv3Desc, check, lib

dcn_version = 1.1   # network/ops_dcnv3/setup.py:63-64

_DTYPES = {torch.float32: _lib.GP_F32, torch.bfloat16: _lib.GP_BF16, torch.float16: _lib.GP_F16,
           torch.float64: _lib.GP_F64}


def _vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _validate(named_tensors, input, group, group_channels, im2col_step):
    # dcnv3_cuda.cu:29-35 / dcnv3.h:37
    for name, t in named_tensors:
        if not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU" if name == "input" else f"{name} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
    if input.dim() != 4:
        raise RuntimeError("input must be (N, H, W, C) channel-last")
    dt = _DTYPES.get(input.dtype)
    if dt is None:
        raise RuntimeError(f"dcnv3: unsupported dtype {input.dtype}")
    for name, t in named_tensors:
        if t.dtype != input.dtype or t.device != input.device:
            raise RuntimeError(f"{name} must have the dtype and device of input ({input.dtype}, {input.device})")
    batch, _, _, channels = input.shape
    step = min(batch, int(im2col_step))
    if step <= 0 or batch % step != 0:   # dcnv3_cuda.cu:46-49
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")
    if channels != group * group_channels:   # dcnv3_cuda.cu:50-53
        raise RuntimeError(
            f"Input channels and group times group channels wont match: ({channels} vs {group * group_channels}).")
    return dt


def _desc(input, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
          group_channels, offset_scale, remove_center):
    N, H, W, _ = input.shape
    Ho = lib.gp_dcnv3_out_size(H, kernel_h, stride_h, pad_h, dilation_h)
    Wo = lib.gp_dcnv3_out_size(W, kernel_w, stride_w, pad_w, dilation_w)
    d = DCNv3Desc(N, H, W, group, group_channels, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                  dilation_w, int(bool(remove_center)), Ho, Wo, float(offset_scale))
    P = kernel_h * kernel_w - int(bool(remove_center))
    return d, Ho, Wo, P


def _check_prefix(offset, mask, N, Ho, Wo, group, P):
    # The reference does not validate offset/mask shapes (flat addressing); reading past the end of
    # a too-small buffer would be a fault there.  We refuse instead.
    need = N * Ho * Wo * group * P
    if offset.numel() < 2 * need or mask.numel() < need:
        raise RuntimeError(f"offset/mask too small for the flat [N*Ho*Wo*G*P] addressing: need {2 * need}/{need} "
                           f"elements, got {offset.numel()}/{mask.numel()}")


def dcnv3_forward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                  group, group_channels, offset_scale, im2col_step, remove_center=0, *, mask_is_logits=False):
    """``DCNv3.dcnv3_forward`` (``src/dcnv3.h:20-38``).  ``mask_is_logits=True`` fuses the softmax over the P
    points (``modules/dcnv3.py:332-334``) into the sampler."""
    dt = _validate((("input", input), ("offset", offset), ("mask", mask)), input, group, group_channels, im2col_step)
    d, Ho, Wo, P = _desc(input, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                         group_channels, offset_scale, remove_center)
    _check_prefix(offset, mask, input.shape[0], Ho, Wo, group, P)
    out = torch.empty((input.shape[0], Ho, Wo, group * group_channels), dtype=input.dtype, device=input.device)
    fn = lib.gp_dcnv3_forward_softmax if mask_is_logits else lib.gp_dcnv3_forward
    with torch.cuda.device(input.device):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(fn(_vp(input), _vp(offset), _vp(mask), _vp(out), ctypes.byref(d), dt, stream), "dcnv3_forward")
    return out


def dcnv3_forward_packed(input, offset_mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                         group, group_channels, offset_scale, im2col_step, remove_center=0):
    """``dcnv3_forward(..., mask_is_logits=True)`` with offsets and mask logits packed in the rows of ONE tensor
    ``offset_mask`` (rows, pitch): ``[G*P*2 offsets | G*P logits | padding]`` -- the output of the fused offset||mask Linear
    (``modules/dcnv3.py:330-334``), consumed without splitting it into two contiguous tensors."""
    dt = _validate((("input", input), ("offset_mask", offset_mask)), input, group, group_channels, im2col_step)
    d, Ho, Wo, P = _desc(input, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                         group_channels, offset_scale, remove_center)
    pitch = offset_mask.shape[-1]
    rows = offset_mask.numel() // pitch
    if offset_mask.dim() < 2 or pitch < group * P * 3 or pitch % 2 or rows < input.shape[0] * Ho * Wo:
        raise RuntimeError(f"offset_mask must hold >= N*Ho*Wo rows of >= G*P*3 (even) elements, got {tuple(offset_mask.shape)}")
    out = torch.empty((input.shape[0], Ho, Wo, group * group_channels), dtype=input.dtype, device=input.device)
    with torch.cuda.device(input.device):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib.gp_dcnv3_forward_softmax_packed(_vp(input), _vp(offset_mask), _vp(out), pitch, ctypes.byref(d), dt, stream),
              "dcnv3_forward_packed")
    return out


def dcnv3_backward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                   group, group_channels, offset_scale, grad_output, im2col_step, remove_center=0):
    """``DCNv3.dcnv3_backward`` (``src/dcnv3.h:40-59``); note ``grad_output`` sits before ``im2col_step``."""
    dt = _validate((("input", input), ("offset", offset), ("mask", mask), ("grad_output", grad_output)), input, group,
                   group_channels, im2col_step)
    d, Ho, Wo, P = _desc(input, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                         group_channels, offset_scale, remove_center)
    _check_prefix(offset, mask, input.shape[0], Ho, Wo, group, P)
    if tuple(grad_output.shape) != (input.shape[0], Ho, Wo, group * group_channels):
        raise RuntimeError(f"grad_output has shape {tuple(grad_output.shape)}, expected "
                           f"{(input.shape[0], Ho, Wo, group * group_channels)}")
    grad_input = torch.empty_like(input)
    grad_offset = torch.empty_like(offset)
    grad_mask = torch.empty_like(mask)
    ws_bytes = lib.gp_dcnv3_backward_workspace(ctypes.byref(d), dt)
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=input.device) if ws_bytes else None
    with torch.cuda.device(input.device):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib.gp_dcnv3_backward(_vp(input), _vp(offset), _vp(mask), _vp(grad_output), _vp(grad_input),
                                    _vp(grad_offset), _vp(grad_mask), grad_offset.numel(), grad_mask.numel(),
                                    _vp(ws), ws_bytes, ctypes.byref(d), dt, stream), "dcnv3_backward")
    return [grad_input, grad_offset, grad_mask]


def dcnv3_sample_index(offset, N, H, W, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                       group, offset_scale, remove_center=0):
    """Parity hook: the floor()ed corners and bounds flags the kernels use (``gp_dcnv3_sample_index``)."""
    if not offset.is_cuda or not offset.is_contiguous():
        raise RuntimeError("offset must be a contiguous CUDA tensor")
    dt = _DTYPES[offset.dtype]
    Ho = lib.gp_dcnv3_out_size(H, kernel_h, stride_h, pad_h, dilation_h)
    Wo = lib.gp_dcnv3_out_size(W, kernel_w, stride_w, pad_w, dilation_w)
    P = kernel_h * kernel_w - int(bool(remove_center))
    d = DCNv3Desc(N, H, W, group, 1, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                  int(bool(remove_center)), Ho, Wo, float(offset_scale))
    n = N * Ho * Wo * group * P
    if offset.numel() < 2 * n:
        raise RuntimeError("offset too small")
    hw = torch.empty((n, 2), dtype=torch.int32, device=offset.device)
    flags = torch.empty((n,), dtype=torch.uint8, device=offset.device)
    with torch.cuda.device(offset.device):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib.gp_dcnv3_sample_index(_vp(offset), _vp(hw), _vp(flags), ctypes.byref(d), dt, stream),
              "dcnv3_sample_index")
    return hw, flags


class DCNv3Function(Function):
    """Autograd boundary, same call signature as ``functions/dcnv3_func.py:22-77``."""

    @staticmethod
    def forward(ctx, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                dilation_w, group, group_channels, offset_scale, im2col_step, remove_center):
        ctx.geom = (kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                    group_channels, offset_scale)
        ctx.im2col_step = im2col_step
        ctx.remove_center = remove_center
        output = dcnv3_forward(input, offset, mask, *ctx.geom, im2col_step, remove_center)
        ctx.save_for_backward(input, offset, mask)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, mask = ctx.saved_tensors
        grad_input, grad_offset, grad_mask = dcnv3_backward(
            input, offset, mask, *ctx.geom, grad_output.contiguous(), ctx.im2col_step, ctx.remove_center)
        return (grad_input, grad_offset, grad_mask) + (None,) * 13
