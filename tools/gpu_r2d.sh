#!/bin/bash
# Round-2 session d: tcgen05 3x3 convolution parity + bench, mode-2 backward fix, PoseNet tests with the new decoder path.
TAG=${1:-r2d}
mkdir -p gpurun_out
{
echo "== pytest conv3x3"; timeout 600 python -m pytest tests/test_conv3x3_gpu.py -m gpu -q -x 2>&1 | tail -25
echo "== bench conv3x3"; timeout 300 python tools/bench_conv3x3.py gpurun_out/${TAG}_conv3x3.json 2>&1 | tail -8
echo "== pytest dcnv3 variants"; timeout 900 python -m pytest tests/test_dcnv3_gpu.py -m gpu -q --maxfail=5 -k "variants" 2>&1 | tail -8
echo "== pytest posenet"; timeout 1200 python -m pytest tests/test_posenet_gpu.py -m gpu -q --maxfail=8 2>&1 | tail -25
} > gpurun_out/${TAG}_log.txt 2>&1
tail -90 gpurun_out/${TAG}_log.txt | cut -c1-600
