#!/bin/bash
# Quick GPU iteration: parity tests, tuning sweep, one ncu full capture of the two sampling kernels.
# Usage (under gpurun): bash tools/gpu_quick.sh <tag> [sweep|nosweep] [ncu|noncu]
TAG=${1:-q}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== bench f32"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1
echo "== bench bf16"; timeout 600 python bench.py --dtype bf16 --no-cpu-baseline 2>&1 | tail -1
if [ "${2:-sweep}" = "sweep" ]; then
echo "== sweep"; timeout 900 python tools/sweep_dcnv3.py --out gpurun_out/sweep_${TAG}.json 2>&1 | tail -150
fi
} > gpurun_out/${TAG}_log.txt 2>&1
if [ "${3:-ncu}" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcnv3_ -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_f32 \
    python tools/profile_target.py f32 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
tail -3 gpurun_out/${TAG}_log.txt
