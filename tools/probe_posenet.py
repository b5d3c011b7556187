import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
from oracle import posenet as OP
from givepose_b200.posenet import PoseNet, PoseNetConfig
GOLD = np.load('tests/golden/posenet.npz')
def rel(a,b):
    a,b=torch.as_tensor(a).double().cpu(),torch.as_tensor(b).double().cpu(); return ((a-b).abs().max()/b.abs().max()).item()
for mode in ('o1','reference'):
    ora = OP.PoseNet().eval(); OP.init_weights(ora, mode, 0)
    for prec in ('fp32','bf16'):
        net = PoseNet(PoseNetConfig(precision=prec)).eval(); net.load_state_dict(ora.state_dict()); net.cuda()
        with torch.no_grad(): out = net(OP.make_inputs(8,0),'cuda')
        print(mode, prec, {k: f"{rel(out[k], GOLD[f'{mode}/{k}']):.2e}" for k in ('rot','trans','size','nocs_coor','ivfc_coor')})
# timing B=64 fp32 / bf16
import time
ora = OP.PoseNet().eval(); OP.init_weights(ora, 'o1', 0)
for prec in ('fp32','bf16'):
    net = PoseNet(PoseNetConfig(precision=prec)).eval(); net.load_state_dict(ora.state_dict()); net.cuda()
    for B in (64, 256):
        data = {k: v.cuda() for k,v in OP.make_inputs(B,0).items()}
        with torch.no_grad():
            for _ in range(3): net(data,'cuda')
            torch.cuda.synchronize(); t0=time.perf_counter()
            for _ in range(5): net(data,'cuda')
            torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/5
        print(prec, 'B',B, f'{dt*1e3:.1f} ms  {B/dt:.0f} RoIs/s')
from torch.profiler import profile, ProfilerActivity
net = PoseNet(PoseNetConfig(precision='bf16')).eval(); net.load_state_dict(ora.state_dict()); net.cuda()
data = {k: v.cuda() for k,v in OP.make_inputs(256,0).items()}
with torch.no_grad():
    net(data,'cuda')
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net(data,'cuda'); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
