#!/bin/bash
# One GPU-box session: tests, smoke, bench (+reference arm), tuning sweep, ncu launch list + full capture.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== host"; nproc; grep -m1 "model name" /proc/cpuinfo
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=25 2>&1 | tail -40
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1
echo "== bench f32"; timeout 600 python bench.py 2>&1 | tail -1
echo "== bench bf16"; timeout 600 python bench.py --dtype bf16 --no-cpu-baseline 2>&1 | tail -1
echo "== bench f32 dist M"; timeout 600 python bench.py --dist M --no-cpu-baseline --no-e2e 2>&1 | tail -1
} > gpurun_out/${TAG}_log.txt 2>&1
# ncu: launch list of the bench command, then one full capture of the two sampling kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcnv3_ -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_f32 \
    python tools/profile_target.py f32 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcnv3_ -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_bf16 \
    python tools/profile_target.py bf16 3 >> gpurun_out/${TAG}_ncu_full.log 2>&1
tail -5 gpurun_out/${TAG}_log.txt
