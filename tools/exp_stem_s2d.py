#!/usr/bin/env python
"""Experiment (GPU box): 7x7/2 stem conv (Cin=3) as a 4x4/1 conv on the 2x2 space-to-depth input (Cin=12 -> 16)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.randn(B, 3, 256, 256, device=dev)
w = torch.randn(64, 3, 7, 7, device=dev) * 0.05
b = torch.randn(64, device=dev)


def timeit(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / reps


def s2d_weight(w, cpad):
    w8 = F.pad(w, (1, 0, 1, 0))                                   # (64,3,8,8), zero tap at index 0
    w8 = w8.reshape(64, 3, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4)   # o, c, ry, rx, a, b
    w8 = w8.reshape(64, 12, 4, 4)
    return F.pad(w8, (0, 0, 0, 0, 0, cpad - 12))


def s2d_input(x, cpad):
    s = F.pixel_unshuffle(x, 2)                                    # (B,12,128,128), channel = c*4 + ry*2 + rx
    s = F.pad(s, (2, 1, 2, 1, 0, cpad - 12))                       # spatial pad 2 top/left, 1 bottom/right; channel pad
    return s.contiguous(memory_format=torch.channels_last)


with torch.no_grad():
    ref = F.conv2d(x, w, b, 2, 3)
    for cpad in (12, 16):
        y = F.conv2d(s2d_input(x, cpad), s2d_weight(w, cpad), b)
        print("fp32 C", cpad, "shape", tuple(y.shape), "max abs diff", (y - ref).abs().max().item(), "ref max", ref.abs().max().item())
    xb, wb, bb = x.bfloat16().contiguous(memory_format=torch.channels_last), w.bfloat16().contiguous(memory_format=torch.channels_last), b.bfloat16()
    print("bf16 direct 7x7/2 C=3 ms", round(timeit(lambda: F.conv2d(xb, wb, bb, 2, 3)), 3))
    for cpad in (12, 16, 32):
        xs, ws = s2d_input(x, cpad).bfloat16(), s2d_weight(w, cpad).bfloat16().contiguous(memory_format=torch.channels_last)
        for bm in (False, True):
            torch.backends.cudnn.benchmark = bm
            t = timeit(lambda: F.conv2d(xs, ws, bb))
            print(f"bf16 s2d C={cpad} benchmark={bm} ms", round(t, 3))
        if hasattr(torch, "cudnn_convolution_relu"):
            t = timeit(lambda: torch.cudnn_convolution_relu(xs, ws, bb, (1, 1), (0, 0), (1, 1), 1))
            print(f"bf16 s2d C={cpad} cudnn_convolution_relu ms", round(t, 3))
        y = F.conv2d(xs, ws, bb)
        print("   max abs diff vs fp32 ref", (y.float() - ref).abs().max().item())
