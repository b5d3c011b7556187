#!/usr/bin/env python
"""Where the training step's time goes (GPU box): wall vs device-busy time, top kernels.  python tools/profile_train.py [B]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import build_posenet, posenet_inputs  # noqa: E402
from givepose_b200.loss import PoseLoss, make_loss_inputs  # noqa: E402
from givepose_b200.train import GradBucket, train_step  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
dev = torch.device("cuda", 0)
_, net = build_posenet("bf16", dev)
data = {k: v.to(dev) for k, v in posenet_inputs(B, 100).items()}
tgt = {k: v.to(dev) for k, v in make_loss_inputs(B, 0).items()}
crit = PoseLoss().to(dev)
opt = torch.optim.SGD(net.parameters(), lr=1e-5, momentum=0.9)
bucket = GradBucket(net.parameters())
for _ in range(3):
    train_step(net, data, tgt, opt, bucket, dev, criterion=crit)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    train_step(net, data, tgt, opt, bucket, dev, criterion=crit)
torch.cuda.synchronize()
print(f"B={B}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms/step wall")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    train_step(net, data, tgt, opt, bucket, dev, criterion=crit)
    torch.cuda.synchronize()
ka = prof.key_averages()
dev_ms = sum(e.self_device_time_total for e in ka) / 1e3
n_k = sum(e.count for e in ka if e.self_device_time_total > 0)
print(f"device-busy {dev_ms:.2f} ms in {n_k} kernels/memcpys")
rows = sorted((e for e in ka if e.self_device_time_total > 0), key=lambda e: -e.self_device_time_total)
for e in rows[:45]:
    print(f"{e.self_device_time_total / 1e3:8.3f} ms {e.count:5d} x  {e.key[:150]}")
