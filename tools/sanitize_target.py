#!/usr/bin/env python
"""Small-shape tour of every hand-written kernel for compute-sanitizer (memcheck / racecheck / initcheck are 10-50x slow):
DCNv3 forward / backward (tiled + generic, border-heavy offsets, stride 2 flat prefix), the PoseNet glue kernels, the training-step
nodes, the tcgen05 dense layer, the RoI pipeline, pose decode, attention.  Prints 'tour done' when every launch returned."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import givepose_b200.functions as F  # noqa: E402
from givepose_b200 import ops, roi  # noqa: E402

g = torch.Generator().manual_seed(0)
dev = "cuda"
r = lambda *s: torch.randn(*s, generator=g)

# DCNv3: tiled fp32 / bf16, generic (gc 30), stride 2 with full-resolution offset buffers, offsets far outside the image
for dtype in (torch.float32, torch.bfloat16):
    for (N, H, W, G, gc, s, scale) in ((2, 17, 23, 2, 32, 1, 1.0), (4, 16, 16, 4, 64, 2, 1.0), (1, 9, 9, 8, 32, 1, 6.0), (2, 8, 11, 3, 30, 1, 1.0)):
        Ho, Wo = (H + 2 - 3) // s + 1, (W + 2 - 3) // s + 1
        inp = r(N, H, W, G * gc).to(dev, dtype)
        off = (r(N, H, W, G * 18) * 3).to(dev, dtype)
        m = torch.softmax(r(N, H, W, G, 9), -1).reshape(N, H, W, G * 9).to(dev, dtype)
        gout = r(N, Ho, Wo, G * gc).to(dev, dtype)
        a = (3, 3, s, s, 1, 1, 1, 1, G, gc, scale)
        F.dcnv3_forward(inp, off, m, *a, 256, 0)
        F.dcnv3_forward(inp, off, m, *a, 256, 0, mask_is_logits=True)
        F.dcnv3_backward(inp, off, m, *a, gout, 256, 0)
F.dcnv3_sample_index(off.float(), 2, 8, 11, 3, 3, 1, 1, 1, 1, 1, 1, 3, 1.0)
# both forward kernels (GP_OPT_FWD_MODE 0: per-thread row reads, 1: TMA-staged rows) and every backward variant (0 scatter,
# 1 split + binned grad_input, 2 fused in-SM aggregation) on a ragged, border-heavy shape; packed offset||logits rows
from givepose_b200._lib import lib  # noqa: E402
for dtype in (torch.float32, torch.bfloat16):
    N, H, W, G, gc = 2, 13, 19, 4, 32
    inp = r(N, H, W, G * gc).to(dev, dtype)
    off = (r(N, H, W, G * 18) * 4).to(dev, dtype)
    m = torch.softmax(r(N, H, W, G, 9), -1).reshape(N, H, W, G * 9).to(dev, dtype)
    packed = torch.cat([off, m, torch.zeros(N, H, W, 4, device=dev, dtype=dtype)], -1).contiguous()
    gout = r(N, H, W, G * gc).to(dev, dtype)
    a = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    for fm in (0, 1):
        lib.gp_set_option(4, fm)
        F.dcnv3_forward(inp, off, m, *a, 256, 0)
        F.dcnv3_forward_packed(inp, packed, *a, 256, 0)
    for bm in (0, 1, 2):
        lib.gp_set_option(0, bm)
        F.dcnv3_backward(inp, off, m, *a, gout, 256, 0)
    lib.gp_set_option(0, 0)
# tcgen05 implicit-GEMM 3x3 convolution + GroupNorm statistics in the epilogue, TMA bulk-store epilogue; apply passes on its statistics
for (N, H, W, Cin) in ((3, 16, 16, 64), (2, 32, 32, 128), (1, 64, 64, 64)):
    xc = r(N, H, W, Cin).to(dev).bfloat16()
    wp = ops.pack_conv3x3_weight((r(256, Cin, 3, 3) / (9 * Cin) ** 0.5).to(dev))
    yc, st = ops.conv3x3_gn_bf16(xc, wp)
    ops.groupnorm_apply(yc, st, torch.ones(256, device=dev), torch.zeros(256, device=dev), 32, 1e-5, "gelu", upsample2x=True)
    ops.groupnorm_apply_conv1x1(yc, st, torch.ones(256, device=dev), torch.zeros(256, device=dev), r(3, 256).to(dev), r(3).to(dev))
    ops.conv3x3_gn_bf16(xc, wp, stats=False)
# glue kernels
for dtype in (torch.float32, torch.bfloat16):
    x = r(3, 10, 12, 256).to(dev, dtype)
    gam, bet = (torch.rand(256, generator=g) + 0.5).to(dev), r(256).to(dev)
    for act in ("none", "relu", "gelu"):
        xa = x.clone().requires_grad_(True)
        y = ops.GroupNormAct.apply(xa, gam.clone().requires_grad_(True), bet.clone().requires_grad_(True), 32, 1e-5, act)
        y.backward(r(3, 10, 12, 256).to(dev, dtype))
    ops.groupnorm_act(x, gam, bet, 32, 1e-5, "gelu", upsample2x=True)
    ops.groupnorm_act_conv1x1(x, gam, bet, r(3, 256).to(dev), r(3).to(dev))
    ops.upsample_bilinear2x(r(2, 5, 9, 64).to(dev, dtype))                 # generic kernel
    ops.upsample_bilinear2x(x[:, :8, :8].contiguous())                      # strip kernel
    ops.upsample_bilinear2x_backward(r(2, 10, 18, 64).to(dev, dtype))
    ops.maxpool3x3s2(r(2, 9, 13, 64).to(dev, dtype), relu=True)
    ops.dwconv3x3_ln_gelu(x, r(9, 256).to(dev), r(256).to(dev), gam, bet, rows=3 * 5 * 6)
    x3 = r(4, 16, 16, 3).to(dev, dtype)
    ops.small_k_linear(x3, r(3, 256).to(dev), r(256).to(dev))
    ops.smallk_dwconv3x3_ln_gelu(x3, r(9, 4, 256).to(dev), r(256).to(dev), gam, bet, rows=4 * 8 * 8)
    ops.dcnv3_smallk_fused(x3, r(4 * 8 * 8, 72).to(dev, dtype), r(4 * 8 * 8, 36).to(dev, dtype), r(16, 256).to(dev), r(256).to(dev),
                           (3, 3, 2, 2, 1, 1, 1, 1, 4, 64, 1.0))
    ops.mhsa_tokens(r(2, 64, 3 * 256).to(dev, dtype), 8)
ops.stem_s2d_pack(r(2, 3, 32, 32).to(dev), torch.bfloat16)
pk = ops.stem_s2d_pack(r(3, 3, 256, 256).to(dev), torch.bfloat16)
for pool in (False, True):
    ops.stem_s2d_gemm(pk, (r(64, 256) * 0.1).to(dev).bfloat16(), r(64).to(dev), pool=pool)
ops.bias_add_relu_(r(2, 8, 8, 64).to(dev).bfloat16(), r(2, 8, 8, 64).to(dev).bfloat16(), r(64).to(dev))
ops.linear_bf16(r(300, 256).to(dev).bfloat16(), r(108, 256).to(dev).bfloat16(), r(108).to(dev), "lrelu", 0.1)
ops.pose_decode(r(5, 6).to(dev), r(5, 3).to(dev) + torch.tensor([0, 0, 2.0], device=dev), torch.tensor([[591.0, 0, 322], [0, 590, 244], [0, 0, 1]]).to(dev),
                torch.rand(5, 2, generator=g).to(dev) * 300, torch.rand(5, 2, generator=g).to(dev) * 100 + 50, torch.rand(5, generator=g).to(dev) + 0.2)
# RoI pipeline: RoIs inside, across the border and entirely outside
frames = torch.randint(0, 256, (2, 60, 80, 3), dtype=torch.uint8, generator=g).to(dev)
inst = torch.randint(0, 4, (2, 60, 80), dtype=torch.uint8, generator=g).to(dev)
roi.roi_crops(frames, np.array([[40.0, 30.0], [-20.0, 10.5], [500.0, 500.0]]), np.array([50.0, 90.0, 20.0]), [0, 1, 1], inst, [0, 1, 1], [1, -1, 2],
              img_size=64, out_res=16)
roi.full_image_tensor(frames, [1, 0, 1], resize=(32, 24))
torch.cuda.synchronize()
print("tour done")
