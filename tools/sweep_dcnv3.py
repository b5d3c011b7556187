#!/usr/bin/env python
"""Tuning sweep for the tiled DCNv3 kernels (run on the GPU box): tile shape x groups-per-CTA x dtype x
offset distribution, CUDA-event timing of forward and backward separately.  Writes gpurun_out/sweep_*.json."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import givepose_b200.functions as F  # noqa: E402
from givepose_b200._lib import lib  # noqa: E402
from bench import alg_bytes  # noqa: E402


def inputs(N, H, W, G, gc, s, dist, dtype, full_res):
    gen = torch.Generator(device="cuda").manual_seed(3)
    Ho, Wo = (H + 2 - 3) // s + 1, (W + 2 - 3) // s + 1
    Hm, Wm = (H, W) if full_res else (Ho, Wo)
    if dist == "T":
        inp = torch.rand(N, H, W, G * gc, generator=gen, device="cuda") * 0.01
        off = torch.rand(N, Hm, Wm, G * 18, generator=gen, device="cuda") * 10
        m = torch.rand(N, Hm, Wm, G, 9, generator=gen, device="cuda") + 1e-5
        m = m / m.sum(-1, keepdim=True)
    else:
        inp = torch.randn(N, H, W, G * gc, generator=gen, device="cuda")
        off = torch.randn(N, Hm, Wm, G * 18, generator=gen, device="cuda")
        m = torch.softmax(torch.randn(N, Hm, Wm, G, 9, generator=gen, device="cuda"), -1)
    gout = torch.randn(N, Ho, Wo, G * gc, generator=gen, device="cuda")
    return [t.to(dtype).contiguous() for t in (inp, off, m.reshape(N, Hm, Wm, G * 9), gout)], Ho, Wo


def timeit(fn, reps=10, warm=2):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_dcnv3.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--shape", default="", help="only the shape with this name (K_N64 | M_64to32_N256)")
    a = ap.parse_args()
    shapes = [("K_N64", 64, 64, 64, 8, 32, 1, False), ("M_64to32_N256", 256, 64, 64, 4, 64, 2, True)]
    tiles = [(8, 8, 1), (8, 8, 2), (8, 8, 4), (4, 16, 1), (8, 16, 1), (4, 8, 1), (4, 4, 1), (4, 4, 2), (4, 4, 4), (4, 4, 8),
             (2, 32, 1), (4, 8, 2), (4, 8, 4), (16, 8, 1), (16, 4, 2), (16, 4, 1), (8, 4, 4), (16, 2, 4)]
    if a.quick:
        tiles = tiles[:3]
    rows = []
    for name, N, H, W, G, gc, s, full in shapes:
        if a.shape and name != a.shape:
            continue
        for dtype, dn in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
            for dist in ("T", "M"):
                (inp, off, m, gout), Ho, Wo = inputs(N, H, W, G, gc, s, dist, dtype, full)
                args = (3, 3, s, s, 1, 1, 1, 1, G, gc, 1.0)
                fb, bb = alg_bytes(N, H, W, G * gc, G, 9, Ho, Wo, inp.element_size())
                for th, tw, gs in tiles:
                  for vec16 in ((8, 4) if dn == "bf16" else (8,)):
                    if G % gs:
                        continue
                    lib.gp_set_tuning(th, tw, gs, vec16)
                    tf = timeit(lambda: F.dcnv3_forward(inp, off, m, *args, 256, 0))
                    tb = timeit(lambda: F.dcnv3_backward(inp, off, m, *args, gout, 256, 0))
                    row = dict(shape=name, dtype=dn, dist=dist, tile=(th, tw, gs), vec16=vec16, fwd_ms=round(tf, 4), bwd_ms=round(tb, 4),
                               fwd_GBps=round(fb / tf / 1e6, 1), bwd_GBps=round(bb / tb / 1e6, 1),
                               fwdbwd_GBps=round((fb + bb) / (tf + tb) / 1e6, 1))
                    rows.append(row)
                    print(json.dumps(row), flush=True)
                del inp, off, m, gout
                torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
