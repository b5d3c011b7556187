#!/bin/bash
# Round-2 second GPU session: full parity suite, both bench arms, micro-benchmarks incl. scatter ceilings, bf16 diagnosis.
TAG=${1:-r2b}
mkdir -p gpurun_out
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== host"; nproc; grep -m1 "model name" /proc/cpuinfo; numactl -H 2>/dev/null | head -5
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 2>&1 | tail -40
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== gather rates"; timeout 200 tools/_bin/gather_rates 2>&1 | tail -24
echo "== diag bf16"; timeout 600 python tools/diag_bf16.py ${TAG} 2>&1 | tail -60
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1
echo "== bench f32"; timeout 1200 python bench.py 2>&1 | tail -1
echo "== bench bf16"; timeout 600 python bench.py --dtype bf16 --no-posenet --no-cpu-baseline 2>&1 | tail -1
echo "== posenet profile"; timeout 600 python tools/profile_posenet.py 2>&1 | head -60
} > gpurun_out/${TAG}_log.txt 2>&1
tail -80 gpurun_out/${TAG}_log.txt | cut -c1-600
