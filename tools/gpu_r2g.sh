#!/bin/bash
# Round-2 session g: rows forward kernel (aligned TMA boxes) parity + A/B + ncu; conv with TMA-store epilogue parity + bench + ncu.
TAG=${1:-r2g}
mkdir -p gpurun_out
{
echo "== pytest conv3x3"; timeout 600 python -m pytest tests/test_conv3x3_gpu.py -m gpu -q -x 2>&1 | tail -15
echo "== bench conv3x3"; timeout 300 python tools/bench_conv3x3.py gpurun_out/${TAG}_conv3x3.json 2>&1 | tail -8
echo "== pytest dcnv3"; timeout 1200 python -m pytest tests/test_dcnv3_gpu.py tests/test_ref_ext_gpu.py -m gpu -q --maxfail=8 2>&1 | tail -25
echo "== sweep fwd"; timeout 600 python tools/sweep_bwd.py --fwd-only --out gpurun_out/${TAG}_sweep_fwd.json 2>&1 | tail -12
echo "== pytest posenet"; timeout 1200 python -m pytest tests/test_posenet_gpu.py -m gpu -q --maxfail=8 2>&1 | tail -25
} > gpurun_out/${TAG}_log.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dcnv3_fwd" -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_fwd \
    python tools/profile_target.py f32 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_prof_fwd.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_fwd_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_gn" -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_conv \
    python -c "
import torch, sys
sys.path.insert(0,'.')
from givepose_b200 import ops
x=torch.randn(1024,64,64,256,device='cuda').bfloat16(); w=(torch.randn(256,256,3,3,device='cuda')/48).bfloat16(); wp=ops.pack_conv3x3_weight(w)
for _ in range(3): ops.conv3x3_gn_bf16(x,wp)
torch.cuda.synchronize()
" > gpurun_out/${TAG}_ncu_conv.log 2>&1
ncu -i gpurun_out/${TAG}_prof_conv.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_conv_raw.csv 2>/dev/null
tail -90 gpurun_out/${TAG}_log.txt | cut -c1-500
