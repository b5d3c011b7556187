#!/bin/bash
# Round-2 session i: 16-warp conv epilogue, forward occupancy variants, e2e chunk count.
TAG=${1:-r2i}
mkdir -p gpurun_out
{
echo "== pytest conv3x3"; timeout 600 python -m pytest tests/test_conv3x3_gpu.py -m gpu -q -x 2>&1 | tail -5
echo "== bench conv3x3"; timeout 300 python tools/bench_conv3x3.py gpurun_out/${TAG}_conv3x3.json 2>&1 | tail -8
for mb in 4 5 6; do echo "== sweep fwd GP_ROWS_MINB=$mb"; GP_ROWS_MINB=$mb timeout 600 python tools/sweep_bwd.py --fwd-only --out gpurun_out/${TAG}_sweep_fwd_minb$mb.json 2>&1 | head -4 | cut -c1-200; done
for ch in 8 16 32; do echo "== e2e chunks $ch"; timeout 600 python bench.py --no-posenet --no-cpu-baseline --no-ceilings --e2e-chunks $ch 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['e2e'])[:400])"; done
} > gpurun_out/${TAG}_log.txt 2>&1
tail -60 gpurun_out/${TAG}_log.txt | cut -c1-600
