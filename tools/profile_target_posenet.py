#!/usr/bin/env python
"""Small fixed workload for ncu: two bf16 PoseNet inference forwards over B RoIs resident in HBM (default 256)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import build_posenet, posenet_inputs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
_, net = build_posenet("bf16", dev)
data = {k: v.to(dev) for k, v in posenet_inputs(B, 0).items()}
with torch.no_grad():
    for _ in range(2):
        out = net(data, dev)
torch.cuda.synchronize()
print("done", out["trans"].shape)
