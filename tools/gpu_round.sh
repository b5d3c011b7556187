#!/bin/bash
# One GPU-box session: tests, smoke, bench (+reference arm), ncu launch list + full captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
# gpurun_out/ must stay under 64 MiB to be copied back: only the f32 DCNv3 report is kept as .ncu-rep, the others are
# exported to CSV on the box and deleted.
TAG=${1:-r01}
mkdir -p gpurun_out
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== host"; nproc; grep -m1 "model name" /proc/cpuinfo
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=25 2>&1 | tail -40
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1
echo "== bench f32"; timeout 600 python bench.py 2>&1 | tail -1
echo "== bench bf16"; timeout 600 python bench.py --dtype bf16 --no-cpu-baseline --no-posenet 2>&1 | tail -1
echo "== bench f32 dist M"; timeout 600 python bench.py --dist M --no-cpu-baseline --no-e2e --no-posenet 2>&1 | tail -1
echo "== train step profile"; timeout 600 python tools/profile_train.py 48 2>&1 | grep -v "^-" | cut -c1-90,186-250 | head -40
} > gpurun_out/${TAG}_log.txt 2>&1
# ncu: launch list of the bench command, then one full capture of the two sampling kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-posenet > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcnv3_ -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_f32 \
    python tools/profile_target.py f32 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_prof_f32.ncu-rep --page source --csv > gpurun_out/${TAG}_prof_f32_source.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:dcnv3_ -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_bf16 \
    python tools/profile_target.py bf16 3 >> gpurun_out/${TAG}_ncu_full.log 2>&1
# PoseNet glue kernels (second forward only: the first one pays cuDNN's algorithm selection)
timeout 900 ncu --set full --clock-control none -k regex:"gn_|upsample2x|small_?k|stem_s2d|maxpool|dwconv|pose_decode|dcnv3_" -s 45 -c 45 -f -o gpurun_out/${TAG}_prof_posenet \
    python tools/profile_target_posenet.py 256 >> gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_prof_posenet.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_posenet_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_prof_posenet.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | tail -20
tail -5 gpurun_out/${TAG}_log.txt | cut -c1-400
