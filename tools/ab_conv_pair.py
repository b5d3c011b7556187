"""Steady-state timing of the CTA-pair convolution (1024 x 64x64x256 and 1024 x 32x32x256), variants alternating with cuDNN."""
import statistics
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from givepose_b200 import ops  # noqa: E402
from givepose_b200._lib import lib  # noqa: E402

torch.backends.cudnn.benchmark = True


def t(fn, it=30):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it


for R in (64, 32):
    x = torch.randn(1024, R, R, 256, device="cuda").bfloat16()
    w = (torch.randn(256, 256, 3, 3, device="cuda") / 48).bfloat16()
    wp = ops.pack_conv3x3_weight(w)
    wcl = w.contiguous(memory_format=torch.channels_last)
    xn = x.permute(0, 3, 1, 2)
    res = {"pair": [], "cudnn": []}
    lib.gp_conv3x3_set_pair(1)
    for r in range(5):
        res["pair"].append(t(lambda: ops.conv3x3_gn_bf16(x, wp)))
        res["cudnn"].append(t(lambda: F.conv2d(xn, wcl, None, 1, 1)))
    fl = 2.0 * 1024 * R * R * 256 * 2304
    for k, v in res.items():
        print(sys.argv[1] if len(sys.argv) > 1 else "", R, k, [round(a, 3) for a in v], "median", round(statistics.median(v), 3), "ms",
              round(fl / statistics.median(v) / 1e9, 1), "TFLOP/s", flush=True)
