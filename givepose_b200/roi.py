"""RoI input pipeline in front of ``PoseNet.forward``: detections + full frames in, the loader's
per-RoI tensors out -- on the device, for the whole batch, bit-exact with the reference's OpenCV host code.

Reference (one RoI at a time, NumPy + OpenCV on the host):
  ``evaluation/load_data_eval.py:256-289,329-333`` (test time) / ``datasets/load_data_nocs.py:270-305`` (training):
  bbox -> ``bbox_center`` / ``img_scale``; ``crop_resize_by_warp_affine`` (``tools/dataset_utils.py:101-157``,
  ``cv2.warpAffine(INTER_NEAREST)``) of the image at ``img_size``, of ``get_2d_coord_np`` (``:8-30``) at ``out_res`` and of the
  instance mask at ``img_size``; ``(roi / 255.0 - mean) / std``, HWC -> CHW.
Here: ``detection_geometry`` (host arithmetic, same expressions), ``gp_roi_affine_inverse`` (host, double) and ONE kernel launch
(``gp_roi_crop``) that writes ``roi_img``, ``roi_mask`` and ``roi_coord_2d`` in the layout ``PoseNet.forward`` consumes.  Only
the uint8 frames / masks cross PCIe (0.9 MB per 640x480 frame instead of 1 MB of fp32 crops per RoI).
No CPU fallback: host tensors for ``images`` / ``masks`` raise.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from ._lib import check, lib

IMG_MEAN = (0.485, 0.456, 0.406)   # load_data_eval.py / load_data_nocs.py:171-172
IMG_STD = (0.229, 0.224, 0.225)


def detection_geometry(bboxes, im_H: int, im_W: int, pad_scale: float = 1.5, out_res: int = 64) -> Dict[str, np.ndarray]:
    """``bboxes`` (B,4) as ``[y1, x1, y2, x2]`` (Mask-RCNN order, ``tools/eval_utils.py:185-187``) -> the geometry entries of the
    loader's dict (``load_data_eval.py:258-268,329-333``): ``bbox_center`` (B,2) = (cx, cy), ``img_scale`` (B,) =
    ``min(max(h, w) * DZI_PAD_SCALE, max(im_H, im_W))``, ``roi_wh`` (``get_real_hw``, ``eval_utils.py:243-249`` -- its default
    480 x 640 clamp is what the reference calls it with), ``resize_ratio = out_res / img_scale``.  float64 like the reference."""
    b = np.asarray(bboxes, dtype=np.float64).reshape(-1, 4)
    y1, x1, y2, x2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    center = np.stack([0.5 * (x1 + x2), 0.5 * (y1 + y2)], axis=1)
    scale = np.minimum(np.maximum(y2 - y1, x2 - x1) * pad_scale, float(max(im_H, im_W))) * 1.0
    wh = np.stack([np.minimum(640.0, x2) - np.maximum(0.0, x1), np.minimum(480.0, y2) - np.maximum(0.0, y1)], axis=1)
    return {"bbox_center": center, "img_scale": scale, "roi_wh": wh.astype(np.float32), "resize_ratio": (out_res / scale)}


def roi_affine_inverse(bbox_center, img_scale, out_size: int) -> np.ndarray:
    """(B,6) float64: for every RoI the dst->src matrix ``cv2.warpAffine`` samples with, i.e. ``get_affine_transform(center,
    (scale, scale), 0, (out, out))`` (``tools/dataset_utils.py:116-157``) inverted as in ``cv::warpAffine``.  Host, C ABI."""
    c = np.ascontiguousarray(np.asarray(bbox_center, dtype=np.float64).reshape(-1, 2))
    s = np.ascontiguousarray(np.asarray(img_scale, dtype=np.float64).reshape(-1))
    if c.shape[0] != s.shape[0]:
        raise RuntimeError("roi_affine_inverse: bbox_center and img_scale disagree on the number of RoIs")
    out = np.empty((c.shape[0], 6), dtype=np.float64)
    check(lib.gp_roi_affine_inverse(c.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p), c.shape[0], int(out_size),
                                    out.ctypes.data_as(ctypes.c_void_p)), "roi_affine_inverse")
    return out


def normalisation_table(mean=IMG_MEAN, std=IMG_STD) -> torch.Tensor:
    """[3][256] fp32: ``float32((v / 255.0 - mean[c]) / std[c])`` evaluated in float64 like the reference's NumPy expression."""
    v = np.arange(256, dtype=np.float64)
    return torch.from_numpy(np.stack([((v / 255.0 - m) / s) for m, s in zip(mean, std)]).astype(np.float32))


def _i32(x, B, dev, name, lo=None, hi=None):
    """Per-RoI int32 vector on the device; range-checked on the HOST before the upload (no device sync)."""
    t = torch.as_tensor(x, dtype=torch.int32, device="cpu" if not torch.is_tensor(x) else None).reshape(-1)
    if t.numel() == 1 and B != 1:
        t = t.expand(B)
    if t.numel() != B:
        raise RuntimeError(f"{name}: expected {B} entries, got {t.numel()}")
    if hi is not None and B > 0 and not t.is_cuda and (int(t.max()) >= hi or int(t.min()) < lo):
        raise RuntimeError(f"roi_crops: {name} out of range")
    return t.contiguous().to(dev, non_blocking=True)


def roi_crops(images: torch.Tensor, bbox_center, img_scale, image_index=0, masks: Optional[torch.Tensor] = None, mask_index=None,
              inst_id=-1, img_size: int = 256, out_res: int = 64, mean=IMG_MEAN, std=IMG_STD, want_coord_2d: bool = True,
              affines: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """The three cropped tensors of the loader's dict for B RoIs.

    ``images`` (M,H,W,3) or (H,W,3) uint8 RGB on the GPU; ``image_index`` (B,) which frame each RoI comes from.
    ``masks`` (K,H,W) or (H,W) uint8 on the GPU with ``mask_index`` (B,) (default: RoI i uses plane i if K == B, else plane
    ``image_index[i]``) and ``inst_id`` (B,): < 0 -> ``roi_mask = float(mask)`` (test time: one binary plane per detection,
    ``load_data_eval.py:284-288``), >= 0 -> ``roi_mask = (mask == inst_id)`` (training: instance-id map, ``load_data_nocs.py:284-291``).
    ``affines``: optional (2B,6) float64 CUDA tensor = ``roi_affine_inverse`` at ``img_size`` stacked on the one at ``out_res``
    (reuse across calls with the same RoIs; skips the host part).
    Returns ``roi_img`` (B,3,S,S), ``roi_mask`` (B,1,S,S) (if masks given), ``roi_coord_2d`` (B,2,R,R), all fp32 on the device."""
    if not images.is_cuda:
        raise RuntimeError("roi_crops: Not implemented on the CPU (images must be a CUDA tensor)")
    if images.dtype != torch.uint8 or images.dim() not in (3, 4) or images.shape[-1] != 3 or not images.is_contiguous():
        raise RuntimeError(f"roi_crops: images must be contiguous uint8 (M,H,W,3), got {images.dtype} {tuple(images.shape)}")
    dev = images.device
    M = 1 if images.dim() == 3 else images.shape[0]
    H, W = images.shape[-3], images.shape[-2]
    c = np.asarray(bbox_center, dtype=np.float64).reshape(-1, 2)
    B = c.shape[0]
    if affines is None:
        minv = torch.from_numpy(np.concatenate([roi_affine_inverse(c, img_scale, img_size), roi_affine_inverse(c, img_scale, out_res)]))
        minv = minv.to(dev, non_blocking=True)
    else:
        minv = affines
        if not minv.is_cuda or minv.dtype != torch.float64 or tuple(minv.shape) != (2 * B, 6) or not minv.is_contiguous():
            raise RuntimeError("roi_crops: affines must be a contiguous CUDA float64 (2B,6) tensor")
    if img_size % 4 or out_res % 4:
        raise RuntimeError("roi_crops: img_size and out_res must be multiples of 4")
    iidx = _i32(image_index, B, dev, "image_index", 0, M)
    lut = normalisation_table(mean, std).to(dev, non_blocking=True)
    roi_img = torch.empty((B, 3, img_size, img_size), dtype=torch.float32, device=dev)
    roi_coord = torch.empty((B, 2, out_res, out_res), dtype=torch.float32, device=dev) if want_coord_2d else None
    roi_mask = midx = iid = None
    K = 0
    if masks is not None:
        if not masks.is_cuda or masks.dtype != torch.uint8 or masks.dim() not in (2, 3) or not masks.is_contiguous() \
                or tuple(masks.shape[-2:]) != (H, W):
            raise RuntimeError(f"roi_crops: masks must be contiguous CUDA uint8 (K,{H},{W}), got {masks.dtype} {tuple(masks.shape)}")
        K = 1 if masks.dim() == 2 else masks.shape[0]
        if mask_index is None:
            mask_index = torch.arange(B, dtype=torch.int32) if K == B else iidx
        midx = _i32(mask_index, B, dev, "mask_index", 0, K)
        iid = _i32(inst_id, B, dev, "inst_id")
        roi_mask = torch.empty((B, 1, img_size, img_size), dtype=torch.float32, device=dev)
    vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for b0 in range(0, B, 65535):   # grid.y limit
            b1 = min(B, b0 + 65535)
            sl = lambda t: vp(t[b0:b1]) if t is not None else None
            check(lib.gp_roi_crop(vp(images), M, H, W, sl(iidx), vp(masks), K, sl(midx), sl(iid), vp(minv[b0:b1]), vp(minv[B + b0:B + b1]),
                                  vp(lut), sl(roi_img), sl(roi_mask), sl(roi_coord), b1 - b0, int(img_size), int(out_res), stream), "roi_crop")
    out = {"roi_img": roi_img}
    if roi_mask is not None:
        out["roi_mask"] = roi_mask
    if roi_coord is not None:
        out["roi_coord_2d"] = roi_coord
    return out


def full_image_tensor(images: torch.Tensor, image_index=None, resize=(256, 256), mean=IMG_MEAN, std=IMG_STD) -> torch.Tensor:
    """``full_img`` of the loaders (the input of ``Scale_net``, ``load_data_eval.py:336-338,348``): ``cv2.resize(frame, resize)``
    (INTER_LINEAR on uint8, ``FLAGS.resize_full`` -- the default) or the frame itself (``resize=None``), then
    ``(v / 255.0 - mean) / std`` -> CHW float32; ``image_index`` (B,) repeats each frame for its RoIs like
    ``np.array([full_img] * len(roi_imgs))``.  Bit-exact with the reference for frames at least as large as ``resize``
    (``resize`` is (width, height) like cv2's dsize); smaller frames raise (OpenCV upscales on a different path)."""
    if not images.is_cuda or images.dtype != torch.uint8 or images.shape[-1] != 3 or not images.is_contiguous():
        raise RuntimeError("full_image_tensor: images must be contiguous CUDA uint8 (M,H,W,3)")
    if images.dim() == 3:
        images = images[None]
    M, H, W, _ = images.shape
    dev = images.device
    lut = normalisation_table(mean, std).to(dev)
    if resize is None:
        idx = images.long()
        out = torch.stack([lut[c][idx[..., c]] for c in range(3)], dim=1)            # (M,3,H,W)
        return out if image_index is None else out[torch.as_tensor(image_index, dtype=torch.long, device=dev)]
    dw, dh = int(resize[0]), int(resize[1])
    B = M if image_index is None else len(image_index)
    iidx = None if image_index is None else _i32(image_index, B, dev, "image_index", 0, M)
    out = torch.empty((B, 3, dh, dw), dtype=torch.float32, device=dev)
    vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    with torch.cuda.device(dev):
        check(lib.gp_resize_linear_u8_normalize(vp(images), M, H, W, vp(iidx), vp(lut), vp(out), B, dh, dw,
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "resize_linear_u8_normalize")
    return out


def posenet_inputs_from_detections(images: torch.Tensor, bboxes, masks: torch.Tensor, cam_K, mean_size, image_index=0, mask_index=None,
                                   inst_id=-1, img_size: int = 256, out_res: int = 64, pad_scale: float = 1.5) -> Dict[str, torch.Tensor]:
    """Frames + detections -> the complete input dict of ``PoseNet.forward`` (keys / shapes / dtypes of
    ``load_data_eval.py:351-375``), crops on the device."""
    H, W = images.shape[-3], images.shape[-2]
    geo = detection_geometry(bboxes, H, W, pad_scale, out_res)
    data = roi_crops(images, geo["bbox_center"], geo["img_scale"], image_index, masks, mask_index, inst_id, img_size, out_res)
    dev = images.device
    f32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).to(dev, non_blocking=True)
    data.update({"bbox_center": f32(geo["bbox_center"]), "roi_wh": f32(geo["roi_wh"]), "resize_ratio": f32(geo["resize_ratio"]),
                 "cam_K": torch.as_tensor(cam_K, dtype=torch.float32).to(dev), "mean_size": torch.as_tensor(mean_size, dtype=torch.float32).to(dev)})
    return data
