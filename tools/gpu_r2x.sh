#!/bin/bash
# ring-depth variants of the CTA-pair convolution (slabs x weight stages), steady-state A/B
mkdir -p gpurun_out
cp givepose_b200/lib/libgivepose_b200.so /tmp/orig.so
for v in s3_b6 s4_b4 s2_b8 s3_b6; do
  cp tools/_bin/variants/lib_$v.so givepose_b200/lib/libgivepose_b200.so
  timeout 200 python tools/ab_conv_pair.py $v 2>&1 | grep -v Warning
done | tee gpurun_out/r2x_ab.txt
cp /tmp/orig.so givepose_b200/lib/libgivepose_b200.so
