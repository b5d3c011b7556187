#!/usr/bin/env python
"""Experiment (GPU box): stand-in backbone stem variants -- cuDNN autotuning and input-channel padding of the 7x7/2 stem."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from bench import build_posenet, posenet_inputs  # noqa: E402
from givepose_b200 import posenet as PN  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
_, net = build_posenet("bf16", dev)
img = posenet_inputs(B, 0)["roi_img"].to(dev)


def timeit(fn, reps=5, warm=3):
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


t = net.backbone.trunk
x_cl = img.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w, b = PN._folded(t.conv1, t.bn1, torch.bfloat16)
with torch.no_grad():
    for bench_mode in (False, True):
        torch.backends.cudnn.benchmark = bench_mode
        print(f"cudnn.benchmark={bench_mode}")
        print("  prep (cast + channels_last) ms", round(timeit(lambda: img.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)), 3))
        print("  stem conv C=3 ms", round(timeit(lambda: F.conv2d(x_cl, w, b, 2, 3)), 3))
        for cp in (4, 8):
            xp = F.pad(x_cl, (0, 0, 0, 0, 0, cp - 3)).contiguous(memory_format=torch.channels_last)
            wp = F.pad(w, (0, 0, 0, 0, 0, cp - 3)).contiguous(memory_format=torch.channels_last)
            print(f"  stem conv C={cp} ms", round(timeit(lambda: F.conv2d(xp, wp, b, 2, 3)), 3),
                  "maxdiff", (F.conv2d(xp, wp, b, 2, 3).float() - F.conv2d(x_cl, w, b, 2, 3).float()).abs().max().item())
        print("  backbone ms", round(timeit(lambda: net.backbone(x_cl)), 3))
        data = {k: v.to(dev) for k, v in posenet_inputs(B, 0).items()}
        print("  full forward ms", round(timeit(lambda: net(data, dev)), 3))
