#!/bin/bash
# Round-2 third GPU session: fused backward (mode 2) parity + sweep + ncu; re-run of the tests that failed in r2b; bench lines.
TAG=${1:-r2c}
mkdir -p gpurun_out
{
echo "== pytest dcnv3 + failed"; timeout 1500 python -m pytest tests/test_dcnv3_gpu.py tests/test_dropin_reference_binding.py -m gpu -q --maxfail=15 2>&1 | tail -25
timeout 900 python -m pytest tests/test_posenet_gpu.py -m gpu -q --maxfail=10 -k "graphed or train_step or bf16_64 or backbone" 2>&1 | tail -25
echo "== sweep bwd"; timeout 1500 python tools/sweep_bwd.py --out gpurun_out/${TAG}_sweep_bwd.json 2>&1 | grep -v '"gin_tile"' | tail -60
echo "== bench f32"; timeout 1500 python bench.py 2>&1 | tail -1
} > gpurun_out/${TAG}_log.txt 2>&1
GP_BWD_MODE=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dcnv3_bwd" -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_fused \
    python tools/profile_target.py f32 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_prof_fused.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_fused_raw.csv 2>/dev/null
tail -70 gpurun_out/${TAG}_log.txt | cut -c1-700
