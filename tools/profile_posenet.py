#!/usr/bin/env python
"""Per-kernel and per-stage time breakdown of the bf16 PoseNet inference forward (B RoIs resident in HBM).

    python tools/profile_posenet.py [B] [out.json]

Prints the torch-profiler kernel table (device time, top 45) and CUDA-event times of the model stages."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import build_posenet, posenet_inputs  # noqa: E402
from givepose_b200 import posenet as PN  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
_, net = build_posenet("bf16", dev)
data = {k: v.to(dev) for k, v in posenet_inputs(B, 0).items()}

stages = {}


def timed(name, fn):
    def wrap(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        stages.setdefault(name, []).append((e0, e1))
        return r
    return wrap


net.backbone.forward = timed("backbone", net.backbone.forward)
net.xyz_nocs_head.forward_nhwc = timed("xyz_nocs_head", net.xyz_nocs_head.forward_nhwc)
net.nocs_encoder.forward_nhwc = timed("nocs_encoder", net.nocs_encoder.forward_nhwc)
net.xyz_deform_head.forward_nhwc = timed("xyz_deform_head", net.xyz_deform_head.forward_nhwc)
net.pnp_net.forward_nhwc = timed("pnp_net", net.pnp_net.forward_nhwc)
net.size_head.forward = timed("size_head", net.size_head.forward)

with torch.no_grad():
    for _ in range(3):
        net(data, dev)
    torch.cuda.synchronize()
    stages.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        net(data, dev)
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / 3
    summary = {k: sum(a.elapsed_time(b) for a, b in v) / 3 for k, v in stages.items()}
    print(f"B={B} bf16: {total:.2f} ms/batch  {B / total * 1e3:.0f} RoIs/s")
    for k, v in summary.items():
        print(f"  {k:18s} {v:8.2f} ms  {100 * v / total:5.1f} %")
    print(f"  {'(other)':18s} {total - sum(summary.values()):8.2f} ms")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net(data, dev)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
if len(sys.argv) > 2:
    rows = [{"name": e.key, "calls": e.count, "device_ms": e.device_time_total / 1e3} for e in prof.key_averages()]
    rows.sort(key=lambda r: -r["device_ms"])
    json.dump({"B": B, "ms_per_batch": total, "stages_ms": summary, "kernels": rows[:60]}, open(sys.argv[2], "w"), indent=1)
