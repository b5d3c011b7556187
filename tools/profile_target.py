#!/usr/bin/env python
"""Small fixed workload for ncu: BASELINE config 2 (N=64, 64x64x256, G=8) fwd + bwd, a few iterations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import givepose_b200.functions as F  # noqa: E402
from bench import ARGS, make_inputs  # noqa: E402

dtype = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == "bf16" else torch.float32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
inp, off, m, gout = make_inputs(64, "T", dtype, torch.device("cuda", 0))
for _ in range(iters):
    out = F.dcnv3_forward(inp, off, m, *ARGS, 256, 0)
    g = F.dcnv3_backward(inp, off, m, *ARGS, gout, 256, 0)
torch.cuda.synchronize()
print("done", out.shape)
