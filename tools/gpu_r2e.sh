#!/bin/bash
# Round-2 session e: TMA-row forward kernel: parity (both forward modes), A/B timing, ncu; bf16 PoseNet bar.
TAG=${1:-r2e}
mkdir -p gpurun_out
{
echo "== pytest dcnv3"; timeout 1200 python -m pytest tests/test_dcnv3_gpu.py tests/test_ref_ext_gpu.py -m gpu -q --maxfail=8 2>&1 | tail -25
echo "== sweep fwd"; timeout 600 python tools/sweep_bwd.py --fwd-only --out gpurun_out/${TAG}_sweep_fwd.json 2>&1 | tail -12
echo "== pytest posenet bf16"; timeout 900 python -m pytest tests/test_posenet_gpu.py -m gpu -q --maxfail=8 -k "bf16" 2>&1 | tail -25
} > gpurun_out/${TAG}_log.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dcnv3_fwd" -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_fwd \
    python tools/profile_target.py f32 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_prof_fwd.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_fwd_raw.csv 2>/dev/null
tail -70 gpurun_out/${TAG}_log.txt | cut -c1-700
