import sys, os, json, torch
sys.path.insert(0, os.getcwd())
import xcorr_b200 as xb
from bench import make_inputs
n,R,K=21_600_000,16,8_100_000
dev=torch.device("cuda:0")
dt=torch.float64 if os.environ.get("XCB_DTYPE","f32")=="f64" else torch.float32
pairs=xb.all_pairs(R)
plan=xb.CorrelationPlan(n,R,pairs,dtype=dt,index_min=n-1,index_max=n-1+K,device=dev)
x=make_inputs(torch,R,n,dev,dt,3000)
ccf=torch.empty((len(pairs),1,K),dtype=dt,device=dev)
for _ in range(3): plan.forward(x); plan.correlate(out=ccf)
torch.cuda.synchronize()
plan.profile(True)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): plan.forward(x); plan.correlate(out=ccf)
e1.record(); torch.cuda.synchronize()
idx,_=plan.argmax(ccf)
ok=bool((idx[:,0]==torch.tensor([137*(j-i) for i,j in pairs],device=dev)).all())
print(os.environ.get("XCB_LIB","default"), "ms/step", round(e0.elapsed_time(e1)/5,3), "ok", ok, [(k,round(ms/5,3)) for k,ms,c in plan.profile_read() if c])
