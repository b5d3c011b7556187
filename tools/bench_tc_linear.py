#!/usr/bin/env python
"""tcgen05 dense layer vs the library path (cuBLAS GEMM + separate activation) at the PoseNet head shapes, B = 1024 RoIs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from givepose_b200 import ops  # noqa: E402

SHAPES = [("dcnv3 input_proj L1 (32x32)", 1024 * 32 * 32, 256, 256, "none"), ("dcnv3 output_proj L0", 1024 * 32 * 32, 256, 256, "none"),
          ("dcnv3 offset||mask L0", 1024 * 32 * 32, 108, 256, "none"), ("feat_reducer", 1024 * 64, 256, 1024, "none"),
          ("fc1||fc1_z + lrelu", 1024, 2048, 8192, "lrelu"), ("fc1||fc1_z + lrelu B=4096", 4096, 2048, 8192, "lrelu"),
          ("fc2 + lrelu", 1024, 256, 1024, "lrelu"), ("M4096 N2048 K8192 (VERDICT r1 bar)", 4096, 2048, 8192, "none"),
          ("square 8192", 8192, 8192, 8192, "none")]


def timeit(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, M, N, K, act in SHAPES:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda")
    bb = b.bfloat16()
    lib = (lambda: F.leaky_relu(F.linear(x, w, bb), 0.1)) if act == "lrelu" else (lambda: F.linear(x, w, bb))
    t_lib = timeit(lib)
    from givepose_b200._lib import lib as _l
    _l.gp_linear_set_pair(0)
    t_one = timeit(lambda: ops.linear_bf16(x, w, b, act, 0.1))
    _l.gp_linear_set_pair(1)
    t_pair = timeit(lambda: ops.linear_bf16(x, w, b, act, 0.1)) if (N % 8 == 0 and M >= 128 and N >= 128) else float("nan")
    _l.gp_linear_set_pair(2)
    t_tc = timeit(lambda: ops.linear_bf16(x, w, b, act, 0.1))
    err = ((ops.linear_bf16(x, w, b, act, 0.1).float() - lib().float()).abs().max() / lib().float().abs().max()).item()
    gb = (M * K + N * K + M * N) * 2 / 1e9
    tf = 2.0 * M * N * K / 1e12
    print(f"{name:36s} M={M:8d} N={N:5d} K={K:5d}  library {t_lib * 1e3:8.1f} us ({tf / t_lib * 1e3:6.0f} TF)  one-CTA {t_one * 1e3:8.1f} us ({tf / t_one * 1e3:6.0f} TF)  "
          f"CTA-pair {t_pair * 1e3:8.1f} us ({tf / t_pair * 1e3:6.0f} TF)  default {t_tc * 1e3:8.1f} us ({gb / t_tc * 1e3:6.0f} GB/s, {tf / t_tc * 1e3:6.0f} TFLOP/s)  "
          f"max rel diff {err:.1e}", flush=True)
