#!/bin/bash
# Round-2 session u: fused input GroupNorm+GELU in the CTA-pair convolution: parity under hard timeouts, decoder/PoseNet A/B.
TAG=${1:-r2u}
mkdir -p gpurun_out
{
echo "== fused smoke"; timeout 90 python -m pytest tests/test_conv3x3_gpu.py -m gpu -q -x -k "fused_input and cta_pair and 3-16-16" 2>&1 | tail -12
rc=${PIPESTATUS[0]}; echo "rc=$rc"
if [ "$rc" = "0" ]; then
  echo "== pytest conv3x3"; timeout 300 python -m pytest tests/test_conv3x3_gpu.py -m gpu -q -x 2>&1 | tail -8
  echo "== pytest posenet"; timeout 900 python -m pytest tests/test_posenet_gpu.py -m gpu -q --maxfail=5 2>&1 | tail -8
  for f in 0 1; do echo "== posenet profile GP_DECODER_FUSE_NORM=$f"; GP_DECODER_FUSE_NORM=$f timeout 600 python tools/profile_posenet.py 1024 2>&1 | head -9; done
  for f in 0 1; do echo "== posenet 4096 GP_DECODER_FUSE_NORM=$f"; GP_DECODER_FUSE_NORM=$f timeout 900 python bench.py --no-cpu-baseline --no-ceilings --no-e2e --no-posenet-fp32 --train-rois 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read())['posenet']; print({k: d[k] for k in ('value','ms_per_batch','tflops')}, d['e2e']['value'])"; done
fi
} > gpurun_out/${TAG}_log.txt 2>&1
tail -60 gpurun_out/${TAG}_log.txt | cut -c1-300
