#!/bin/bash
# epilogue-warp count of the CTA-pair convolution (4 / 8 / 16), steady-state A/B; parity for each
mkdir -p gpurun_out
cp givepose_b200/lib/libgivepose_b200.so /tmp/orig.so
for v in epi16 epi8 epi4 epi16 epi4; do
  cp tools/_bin/variants/lib_$v.so givepose_b200/lib/libgivepose_b200.so
  timeout 200 python tools/ab_conv_pair.py $v 2>&1 | grep -v Warning
done | tee gpurun_out/r2y_ab.txt
for v in epi8 epi4; do cp tools/_bin/variants/lib_$v.so givepose_b200/lib/libgivepose_b200.so; echo "== parity $v"; timeout 200 python -m pytest tests/test_conv3x3_gpu.py -m gpu -q -x 2>&1 | tail -2; done | tee -a gpurun_out/r2y_ab.txt
cp /tmp/orig.so givepose_b200/lib/libgivepose_b200.so
