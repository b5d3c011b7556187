#!/usr/bin/env python
"""A/B sweep of the DCNv3 backward on the GPU box: one-pass scatter kernel (GP_OPT_BWD_MODE 0) against the split backward
(grad_offset/grad_mask kernel + binned grad_input kernel, mode 1) over the binned kernel's tile / CTA size, both offset
distributions, fp32 and bf16, at BASELINE config 2 and at the largest in-model shape.  Every variant is also checked
against mode 0 on the spot (max-norm relative error of the three gradients).  Writes gpurun_out/<tag>_sweep_bwd.json."""
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
from tools.sweep_dcnv3 import inputs, timeit  # noqa: E402

OPT_BWD_MODE, OPT_GIN_TH, OPT_GIN_TW, OPT_GIN_NT, OPT_FWD_MODE = 0, 1, 2, 3, 4


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_bwd.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--fwd-only", action="store_true", help="only the forward-mode A/B (mode 0 vs 1) and the mode-0 backward time")
    a = ap.parse_args()
    shapes = [("K_N64", 64, 64, 64, 8, 32, 1, False), ("M_64to32_N256", 256, 64, 64, 4, 64, 2, True)]
    variants = [(8, 8, 192), (8, 8, 128)]
    if a.quick:
        variants = variants[:1]
    rows = []
    for name, N, H, W, G, gc, s, full in shapes:
        for dtype, dn in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
            for dist in ("T", "M"):
                (inp, off, m, gout), Ho, Wo = inputs(N, H, W, G, gc, s, dist, dtype, full)
                args = (3, 3, s, s, 1, 1, 1, 1, G, gc, 1.0)
                fb, bb = alg_bytes(N, H, W, G * gc, G, 9, Ho, Wo, inp.element_size())
                lib.gp_set_tuning(8, 8, 2, 8)
                tf = {}
                for fm in (0, 1):
                    lib.gp_set_option(OPT_FWD_MODE, fm)
                    tf[fm] = timeit(lambda: F.dcnv3_forward(inp, off, m, *args, 256, 0))
                o0 = None
                if True:
                    lib.gp_set_option(OPT_FWD_MODE, 0)
                    o0 = F.dcnv3_forward(inp, off, m, *args, 256, 0)
                    lib.gp_set_option(OPT_FWD_MODE, 1)
                    o1 = F.dcnv3_forward(inp, off, m, *args, 256, 0)
                    ferr = rel(o1, o0)
                    del o0, o1
                lib.gp_set_option(OPT_BWD_MODE, 0)
                ref = F.dcnv3_backward(inp, off, m, *args, gout, 256, 0)
                t0 = timeit(lambda: F.dcnv3_backward(inp, off, m, *args, gout, 256, 0))
                row = dict(shape=name, dtype=dn, dist=dist, fwd_ms_mode0=round(tf[0], 4), fwd_ms_mode1=round(tf[1], 4), fwd_mode1_vs_mode0=ferr,
                           bwd_mode0_ms=round(t0, 4), bwd_alg_bytes=bb, fwd_alg_bytes=fb)
                print(json.dumps(row), flush=True)
                rows.append(row)
                lib.gp_set_option(OPT_BWD_MODE, 1)
                for th, tw, nt in ([] if a.fwd_only else variants):
                    lib.gp_set_option(OPT_GIN_TH, th)
                    lib.gp_set_option(OPT_GIN_TW, tw)
                    lib.gp_set_option(OPT_GIN_NT, nt)
                    got = F.dcnv3_backward(inp, off, m, *args, gout, 256, 0)
                    errs = [rel(g_, r_) for g_, r_ in zip(got, ref)]
                    del got
                    t1 = timeit(lambda: F.dcnv3_backward(inp, off, m, *args, gout, 256, 0))
                    row = dict(shape=name, dtype=dn, dist=dist, gin_tile=(th, tw), gin_threads=nt, bwd_mode1_ms=round(t1, 4),
                               speedup=round(t0 / t1, 3), err_vs_mode0=[float("%.2e" % e) for e in errs],
                               bwd_GBps=round(bb / t1 / 1e6, 1))
                    print(json.dumps(row), flush=True)
                    rows.append(row)
                lib.gp_set_option(OPT_BWD_MODE, 2)
                for th, tw in (() if a.fwd_only else ((8, 8), (4, 8), (8, 4), (4, 4))):
                    lib.gp_set_option(OPT_GIN_TH, th)
                    lib.gp_set_option(OPT_GIN_TW, tw)
                    got = F.dcnv3_backward(inp, off, m, *args, gout, 256, 0)
                    errs = [rel(g_, r_) for g_, r_ in zip(got, ref)]
                    del got
                    t2 = timeit(lambda: F.dcnv3_backward(inp, off, m, *args, gout, 256, 0))
                    row = dict(shape=name, dtype=dn, dist=dist, fused_tile=(th, tw), bwd_mode2_ms=round(t2, 4), speedup=round(t0 / t2, 3),
                               err_vs_mode0=[float("%.2e" % e) for e in errs], bwd_GBps=round(bb / t2 / 1e6, 1))
                    print(json.dumps(row), flush=True)
                    rows.append(row)
                lib.gp_set_option(OPT_BWD_MODE, 0)
                lib.gp_set_option(OPT_GIN_TH, 8)
                lib.gp_set_option(OPT_GIN_TW, 8)
                lib.gp_set_option(OPT_GIN_NT, 192)
                del inp, off, m, gout, ref
                torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
