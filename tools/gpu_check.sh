#!/bin/bash
# One GPU-box session: parity tests, smoke, both bench arms, launch list, ncu full captures.
# Usage (under gpurun): bash tools/gpu_check.sh <tag>
TAG=${1:-chk}
mkdir -p gpurun_out
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== host"; nproc; grep -m1 "model name" /proc/cpuinfo
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 2>&1 | tail -30
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1
echo "== bench f32"; timeout 900 python bench.py 2>&1 | tail -1
echo "== bench bf16"; timeout 600 python bench.py --dtype bf16 --no-posenet 2>&1 | tail -1
echo "== bench f32 dist M"; timeout 600 python bench.py --dist M --no-cpu-baseline --no-e2e --no-posenet 2>&1 | tail -1
echo "== tc_linear"; timeout 300 python tools/bench_tc_linear.py 2>&1 | tail -12
echo "== red rates"; timeout 120 tools/_bin/red_rates 2>&1 | tail -14
echo "== sanitizer"; for tool in memcheck racecheck initcheck synccheck; do timeout 600 compute-sanitizer --tool $tool python tools/sanitize_target.py 2>&1 | grep -E "SUMMARY|tour done" | tr '\n' ' '; echo " [$tool]"; done
} > gpurun_out/${TAG}_log.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-posenet > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcnv3_ -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_f32 \
    python tools/profile_target.py f32 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:dcnv3_ -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_bf16 \
    python tools/profile_target.py bf16 3 >> gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_bf16_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_tcl_mem \
    python tools/profile_target_linear.py >> gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_bf16_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_tcl_fc1 \
    python tools/profile_target_linear.py >> gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"roi_crop|smallk_fused|gn_bwd|upsample2x_bwd|gn_apply|gn_stats|stem_s2d_gemm" -f -o gpurun_out/${TAG}_prof_new \
    python tools/profile_target_r06.py >> gpurun_out/${TAG}_ncu_full.log 2>&1
for f in prof_bf16 prof_tcl_mem prof_tcl_fc1 prof_new; do
  ncu -i gpurun_out/${TAG}_${f}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null
done
rm -f gpurun_out/${TAG}_prof_bf16.ncu-rep gpurun_out/${TAG}_prof_new.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | tail -24
cat gpurun_out/${TAG}_log.txt | cut -c1-700 | tail -70
