#!/bin/bash
# tuning only: compares register budgets (GP_MIN_BLOCKS builds, givepose_b200/lib/libgp_mb*.so) of the tiled kernels
mkdir -p gpurun_out
for mb in 4 5 6 8; do
  lib=givepose_b200/lib/libgp_mb$mb.so; [ $mb = 4 ] && lib=givepose_b200/lib/libgivepose_b200.so
  echo "== min_blocks=$mb"
  GIVEPOSE_B200_LIB=$PWD/$lib python tools/sweep_dcnv3.py --quick --out gpurun_out/sweep_mb$mb.json 2>&1 | grep K_N64 | grep '"T"'
done > gpurun_out/minblocks_log.txt 2>&1
tail -40 gpurun_out/minblocks_log.txt
