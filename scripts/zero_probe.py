import sys; sys.path.insert(0,'/root/repo')
import torch, statistics
from pctrans_b200 import _cabi, synth
dev='cuda:0'
value, sh, lsi, loc, attn = synth.make_inputs(16, synth.CVPPP_512, M=8, D=32, P=4, regime="R", seed=0)
g=torch.randn(16, value.shape[1], 256)
v,s,l,lo,at,go=(t.to(dev) for t in (value,sh,lsi,loc,attn,g))
grads=(torch.empty_like(v),torch.empty_like(lo),torch.empty_like(at))
for _ in range(5): _cabi.backward(v,s,l,lo,at,go,out=grads)
ts=[]
for _ in range(50):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); _cabi.backward(v,s,l,lo,at,go,out=grads); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print("bwd call (zero+kernel) median us", statistics.median(ts)*1e3)
z=torch.empty(22*1024*1024//4, device=dev)
ts=[]
for _ in range(50):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); z.zero_(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print("torch zero_ 22MB us", statistics.median(ts)*1e3)
