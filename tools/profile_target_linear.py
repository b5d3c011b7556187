#!/usr/bin/env python
"""ncu target: the tcgen05 dense layer at one memory-bound and one compute-bound PoseNet shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from givepose_b200 import ops  # noqa: E402

for M, N, K in ((262144, 256, 256), (4096, 2048, 8192)):
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda")
    for _ in range(2):
        y = ops.linear_bf16(x, w, b, "lrelu", 0.1)
torch.cuda.synchronize()
print("done")
