#!/bin/bash
# Round-2 session f: TMA box probes + sanitizer on the rows forward kernel; slab version of the tcgen05 convolution.
TAG=${1:-r2f}
mkdir -p gpurun_out
{
echo "== tma probes"
for a in "144 36 8 8 3" "144 32 8 8 3" "144 36 8 8 2" "144 36 8 1 3" "144 40 8 8 3" "144 64 8 8 3" "72 20 8 8 3" "72 16 8 8 3" "72 32 8 8 3"; do timeout 60 tools/_bin/tma_rows_probe $a; done
echo "== sanitizer rows fwd"
timeout 300 compute-sanitizer --tool memcheck python -c "
import torch, sys
sys.path.insert(0,'.')
import givepose_b200.functions as F
g=torch.Generator().manual_seed(0)
N,H,W,G,gc=1,16,16,4,32
inp=torch.randn(N,H,W,G*gc,generator=g).cuda(); off=torch.randn(N,H,W,G*18,generator=g).cuda(); m=torch.rand(N,H,W,G*9,generator=g).cuda()
out=F.dcnv3_forward(inp,off,m,3,3,1,1,1,1,1,1,G,gc,1.0,256,0); torch.cuda.synchronize(); print('ok',out.abs().sum().item())
" 2>&1 | grep -v "^=========     at\|^=========         in\|Host Frame\|^=========$" | head -30
echo "== pytest conv3x3 (GP_FWD_MODE=0)"; GP_FWD_MODE=0 timeout 600 python -m pytest tests/test_conv3x3_gpu.py -m gpu -q -x 2>&1 | tail -15
echo "== bench conv3x3"; GP_FWD_MODE=0 timeout 300 python tools/bench_conv3x3.py gpurun_out/${TAG}_conv3x3.json 2>&1 | tail -8
} > gpurun_out/${TAG}_log.txt 2>&1
GP_FWD_MODE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_gn" -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_conv \
    python -c "
import torch, sys
sys.path.insert(0,'.')
from givepose_b200 import ops
x=torch.randn(1024,64,64,256,device='cuda').bfloat16(); w=(torch.randn(256,256,3,3,device='cuda')/48).bfloat16(); wp=ops.pack_conv3x3_weight(w)
for _ in range(3): ops.conv3x3_gn_bf16(x,wp)
torch.cuda.synchronize()
" > gpurun_out/${TAG}_ncu_conv.log 2>&1
ncu -i gpurun_out/${TAG}_prof_conv.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_conv_raw.csv 2>/dev/null
tail -80 gpurun_out/${TAG}_log.txt | cut -c1-500
