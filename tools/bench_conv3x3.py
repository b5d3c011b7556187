"""Decoder 3x3 convolution (256 -> 256, bf16, channel-last): hand-written tcgen05 implicit GEMM (+ GroupNorm statistics in the
epilogue) against cuDNN (+ the separate gn_stats pass it needs), CUDA events, inputs larger than L2 at the large sizes."""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from givepose_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    out = []
    torch.backends.cudnn.benchmark = True
    for N, R in ((1024, 16), (1024, 32), (256, 64), (1024, 64)):
        x = torch.randn(N, R, R, 256, device="cuda").bfloat16()
        w = (torch.randn(256, 256, 3, 3, device="cuda") / 48).bfloat16()
        wp = ops.pack_conv3x3_weight(w)
        wcl = w.contiguous(memory_format=torch.channels_last)
        gamma, beta = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
        flops = 2.0 * N * R * R * 256 * 2304
        from givepose_b200._lib import lib
        lib.gp_conv3x3_set_pair(1)
        t_pair = timeit(lambda: ops.conv3x3_gn_bf16(x, wp))
        lib.gp_conv3x3_set_pair(0)
        t_tc = timeit(lambda: ops.conv3x3_gn_bf16(x, wp))
        t_tc_nostats = timeit(lambda: ops.conv3x3_gn_bf16(x, wp, stats=False))
        xn = x.permute(0, 3, 1, 2)
        t_cudnn = timeit(lambda: F.conv2d(xn, wcl, None, 1, 1))
        y = F.conv2d(xn, wcl, None, 1, 1).permute(0, 2, 3, 1).contiguous()
        t_gn_full = timeit(lambda: ops.groupnorm_act(y, gamma, beta, 32, 1e-5, "gelu"))
        _, st = ops.conv3x3_gn_bf16(x, wp)
        t_gn_apply = timeit(lambda: ops.groupnorm_apply(y, st, gamma, beta, 32, 1e-5, "gelu"))
        rec = {"N": N, "res": R, "tc_ms": round(t_tc, 4), "tc_nostats_ms": round(t_tc_nostats, 4), "cudnn_ms": round(t_cudnn, 4),
               "tc_TFLOPs": round(flops / t_tc / 1e9, 1), "pair_ms": round(t_pair, 4), "pair_TFLOPs": round(flops / t_pair / 1e9, 1), "cudnn_TFLOPs": round(flops / t_cudnn / 1e9, 1),
               "gn_stats_plus_apply_ms": round(t_gn_full, 4), "gn_apply_only_ms": round(t_gn_apply, 4),
               "convmodule_tc_ms": round(t_tc + t_gn_apply, 4), "convmodule_cudnn_ms": round(t_cudnn + t_gn_full, 4)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
