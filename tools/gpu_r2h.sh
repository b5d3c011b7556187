#!/bin/bash
# Round-2 session h: the full GPU suite, smoke, both bench arms, PoseNet kernel breakdown.
TAG=${1:-r2h}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -30
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== posenet profile"; timeout 600 python tools/profile_posenet.py 1024 gpurun_out/${TAG}_posenet_kernels.json 2>&1 | head -70
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1
echo "== bench f32"; timeout 1500 python bench.py 2>&1 | tail -1
echo "== bench bf16"; timeout 600 python bench.py --dtype bf16 --no-posenet --no-cpu-baseline 2>&1 | tail -1
} > gpurun_out/${TAG}_log.txt 2>&1
tail -120 gpurun_out/${TAG}_log.txt | cut -c1-900
