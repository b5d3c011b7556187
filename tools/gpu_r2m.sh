#!/bin/bash
TAG=${1:-r2m}
mkdir -p gpurun_out
GP_CONV_PAIR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_gn" -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_conv_pair \
    python -c "
import torch, sys
sys.path.insert(0,'.')
from givepose_b200 import ops
x=torch.randn(1024,64,64,256,device='cuda').bfloat16(); w=(torch.randn(256,256,3,3,device='cuda')/48).bfloat16(); wp=ops.pack_conv3x3_weight(w)
for _ in range(3): ops.conv3x3_gn_bf16(x,wp)
torch.cuda.synchronize()
" > gpurun_out/${TAG}_ncu_conv.log 2>&1
ncu -i gpurun_out/${TAG}_prof_conv_pair.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_conv_pair_raw.csv 2>/dev/null
# steady-state A/B: alternate variants, many iterations each, report per-variant median of 5 rounds
timeout 600 python - <<'PY' > gpurun_out/${TAG}_ab.txt 2>&1
import torch, sys, statistics
sys.path.insert(0,'.')
import torch.nn.functional as F
from givepose_b200 import ops
from givepose_b200._lib import lib
x=torch.randn(1024,64,64,256,device='cuda').bfloat16(); w=(torch.randn(256,256,3,3,device='cuda')/48).bfloat16(); wp=ops.pack_conv3x3_weight(w)
wcl=w.contiguous(memory_format=torch.channels_last); xn=x.permute(0,3,1,2)
torch.backends.cudnn.benchmark=True
def t(fn,it=30):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/it
res={'one':[], 'pair':[], 'cudnn':[]}
for r in range(5):
    lib.gp_conv3x3_set_pair(0); res['one'].append(t(lambda: ops.conv3x3_gn_bf16(x,wp)))
    lib.gp_conv3x3_set_pair(1); res['pair'].append(t(lambda: ops.conv3x3_gn_bf16(x,wp)))
    res['cudnn'].append(t(lambda: F.conv2d(xn,wcl,None,1,1)))
fl=2.0*1024*64*64*256*2304
for k,v in res.items(): print(k, [round(a,3) for a in v], 'median', round(statistics.median(v),3), 'ms', round(fl/statistics.median(v)/1e9,1), 'TFLOP/s')
PY
cat gpurun_out/${TAG}_ab.txt
