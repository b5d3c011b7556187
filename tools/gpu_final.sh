#!/bin/bash
# Final validation of a tree: full GPU suite, smoke, reference arm, bench (fp32 line incl. PoseNet, bf16 DCNv3 line).
TAG=${1:-final}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -8
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1
echo "== bench f32"; timeout 1500 python bench.py 2>&1 | tail -1
echo "== bench bf16"; timeout 600 python bench.py --dtype bf16 --no-posenet --no-cpu-baseline 2>&1 | tail -1
} > gpurun_out/${TAG}_log.txt 2>&1
grep -A1 "== pytest\|== smoke" gpurun_out/${TAG}_log.txt | tail -8 | cut -c1-300
