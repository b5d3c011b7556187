#!/bin/bash
# Round-2 first GPU session: parity of the DCNv3 kernels (all backward variants), A/B sweep, micro-benchmarks, ncu.
TAG=${1:-r2a}
mkdir -p gpurun_out
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== pytest dcnv3"; timeout 1200 python -m pytest tests/test_dcnv3_gpu.py tests/test_ref_ext_gpu.py -m gpu -q --maxfail=10 2>&1 | tail -25
echo "== gather rates"; timeout 200 tools/_bin/gather_rates 2>&1 | tail -20
echo "== sweep bwd"; timeout 1500 python tools/sweep_bwd.py --out gpurun_out/${TAG}_sweep_bwd.json 2>&1 | tail -120
} > gpurun_out/${TAG}_log.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/profile_target.py f32 3 > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dcnv3_(gin|bwd)" -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_f32 \
    python tools/profile_target.py f32 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_prof_f32.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_f32_raw.csv 2>/dev/null
tail -60 gpurun_out/${TAG}_log.txt | cut -c1-400
