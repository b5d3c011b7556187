#!/bin/bash
# Round-2 session j: two-stream decoder: parity + PoseNet throughput A/B.
TAG=${1:-r2j}
mkdir -p gpurun_out
{
echo "== pytest posenet"; timeout 1200 python -m pytest tests/test_posenet_gpu.py -m gpu -q --maxfail=8 2>&1 | tail -12
for st in 1 2; do echo "== posenet profile GP_DECODER_STREAMS=$st"; GP_DECODER_STREAMS=$st timeout 600 python tools/profile_posenet.py 1024 2>&1 | head -9; done
for st in 1 2; do echo "== posenet 4096 GP_DECODER_STREAMS=$st"; GP_DECODER_STREAMS=$st timeout 900 python bench.py --no-cpu-baseline --no-ceilings --no-e2e --no-posenet-fp32 --train-rois 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read())['posenet']; print({k: d[k] for k in ('value','ms_per_batch','tflops')}, d['e2e']['value'], d['e2e_host_crops']['value'], d['latency_8_rois'])"; done
} > gpurun_out/${TAG}_log.txt 2>&1
tail -60 gpurun_out/${TAG}_log.txt | cut -c1-400
