#!/bin/bash
# Round-2 session p (8 GPUs): scaling run of the final tree.
TAG=${1:-r2p}
mkdir -p gpurun_out
{
for n in 8 4; do
echo "== bench $n gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --posenet-steps 3 2>&1 | tail -1
done
} > gpurun_out/${TAG}_log.txt 2>&1
tail -4 gpurun_out/${TAG}_log.txt | cut -c1-600
