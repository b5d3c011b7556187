#!/usr/bin/env python
"""ncu target for the kernels added in r06: gp_roi_crop, dcnv3_smallk_fused, the GroupNorm+act backward passes and the
bilinear x2 backward gather.  Each runs twice (the second launch is the one to read)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from givepose_b200 import ops, roi  # noqa: E402

g = torch.Generator().manual_seed(0)
dev = "cuda"
# RoI crops: 1024 RoIs from 128 frames
B, M = 1024, 128
frames = torch.randint(0, 256, (M, 480, 640, 3), dtype=torch.uint8, generator=g).to(dev)
inst = torch.randint(0, 9, (M, 480, 640), dtype=torch.uint8, generator=g).to(dev)
y1, x1 = torch.randint(0, 300, (B,), generator=g), torch.randint(0, 400, (B,), generator=g)
bboxes = torch.stack([y1, x1, y1 + torch.randint(40, 180, (B,), generator=g), x1 + torch.randint(40, 240, (B,), generator=g)], 1).numpy()
geo = roi.detection_geometry(bboxes, 480, 640)
iidx = (torch.arange(B) // 8).int()
for _ in range(2):
    roi.roi_crops(frames, geo["bbox_center"], geo["img_scale"], iidx, inst, iidx, (torch.arange(B) % 8 + 1).int())
# first-layer fused DCNv3 module: 256 RoIs
x3 = (torch.rand(256, 64, 64, 3, generator=g) - 0.5).bfloat16().to(dev)
off = torch.randn(256 * 32 * 32, 72, generator=g).bfloat16().to(dev)
msk = torch.randn(256 * 32 * 32, 36, generator=g).bfloat16().to(dev)
w2 = torch.randn(16, 256, generator=g).to(dev)
b2 = torch.randn(256, generator=g).to(dev)
for _ in range(2):
    ops.dcnv3_smallk_fused(x3, off, msk, w2, b2, (3, 3, 2, 2, 1, 1, 1, 1, 4, 64, 1.0))
# GroupNorm + GELU forward / backward and the upsampling backward at the training-step shape (48 RoIs, 64x64x256 bf16)
x = torch.randn(48, 64, 64, 256, generator=g).bfloat16().to(dev).requires_grad_(True)
gamma = torch.ones(256, device=dev, requires_grad=True)
beta = torch.zeros(256, device=dev, requires_grad=True)
dy = torch.randn(48, 64, 64, 256, generator=g).bfloat16().to(dev)
for _ in range(2):
    y = ops.GroupNormAct.apply(x, gamma, beta, 32, 1e-5, "gelu")
    y.backward(dy)
    ops.upsample_bilinear2x_backward(dy)
# tcgen05 stem (implicit GEMM + fused max-pool) at 256 RoIs
img = torch.randn(256, 3, 256, 256, generator=g).to(dev)
packed = ops.stem_s2d_pack(img, torch.bfloat16)
w2d = (torch.randn(64, 256, generator=g) * 0.1).bfloat16().to(dev)
for _ in range(2):
    ops.stem_s2d_gemm(packed, w2d, b2[:64].contiguous(), pool=True)
torch.cuda.synchronize()
print("done")
