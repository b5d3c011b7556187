#!/bin/bash
# Round-2 session n (2 GPUs): the scaling path with the round-2 kernels: gpu tests on 1 GPU, bench at N=1 and N=2 via torchrun.
TAG=${1:-r2n}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -8
echo "== bench 1 gpu"; timeout 1200 python bench.py --no-cpu-baseline 2>&1 | tail -1
echo "== bench 2 gpus"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1
} > gpurun_out/${TAG}_log.txt 2>&1
tail -30 gpurun_out/${TAG}_log.txt | cut -c1-1500
