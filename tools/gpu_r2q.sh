#!/bin/bash
# Round-2 session q: CTA-pair dense layer: parity under hard timeouts, then bench.
TAG=${1:-r2q}
mkdir -p gpurun_out
{
echo "== pair smoke"; timeout 90 python -m pytest tests/test_tc_linear_gpu.py -m gpu -q -x -k "cta_pair and 256-256-256 and none" 2>&1 | tail -8
rc=${PIPESTATUS[0]}; echo "rc=$rc"
if [ "$rc" = "0" ]; then
  echo "== pytest tc_linear"; timeout 300 python -m pytest tests/test_tc_linear_gpu.py -m gpu -q -x 2>&1 | tail -8
  echo "== bench tc_linear"; timeout 300 python tools/bench_tc_linear.py 2>&1 | tail -12
fi
} > gpurun_out/${TAG}_log.txt 2>&1
tail -40 gpurun_out/${TAG}_log.txt | cut -c1-400
