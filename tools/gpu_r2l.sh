#!/bin/bash
# Round-2 session l: CTA-pair (cta_group::2) convolution: parity under hard timeouts, then bench.
TAG=${1:-r2l}
mkdir -p gpurun_out
{
echo "== pair smoke (1 shape)"; timeout 90 python -m pytest tests/test_conv3x3_gpu.py -m gpu -q -x -k "cta_pair and matches_torch and 1-16-16-256" 2>&1 | tail -15
rc=${PIPESTATUS[0]}; echo "rc=$rc"
if [ "$rc" = "0" ]; then
  echo "== pytest conv3x3 (both variants)"; timeout 300 python -m pytest tests/test_conv3x3_gpu.py -m gpu -q -x 2>&1 | tail -15
  echo "== bench conv3x3"; timeout 300 python tools/bench_conv3x3.py gpurun_out/${TAG}_conv3x3.json 2>&1 | tail -8
fi
nvidia-smi --query-gpu=name,memory.used --format=csv
} > gpurun_out/${TAG}_log.txt 2>&1
tail -60 gpurun_out/${TAG}_log.txt | cut -c1-700
