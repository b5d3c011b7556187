#!/bin/bash
# Round-2 session k: compute-sanitizer over every hand-written kernel incl. the round-2 ones; ncu launch list of the bench command.
TAG=${1:-r2k}
mkdir -p gpurun_out
{
echo "# compute-sanitizer over tools/sanitize_target.py (every hand-written kernel, small border-heavy shapes), B200, round 2"
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"; timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py 2>&1 | grep -E "tour done|SUMMARY|COMPUTE-SANITIZER|Error|error|hazard|Hazard|Invalid|Uninit" | head -40; echo "exit=${PIPESTATUS[0]}"
done
} > gpurun_out/${TAG}_sanitizer.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-posenet --no-cpu-baseline --no-ceilings > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
tail -40 gpurun_out/${TAG}_sanitizer.txt | cut -c1-300
