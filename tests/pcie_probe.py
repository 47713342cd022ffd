import torch, time
dev=torch.device("cuda:0")
for gb in (1.0, 3.9):
    n=int(gb*1e9/4)
    h=torch.empty(n,dtype=torch.float32).pin_memory(); d=torch.empty(n,dtype=torch.float32,device=dev)
    for name,fn in (("d2h",lambda: h.copy_(d,non_blocking=True)),("h2d",lambda: d.copy_(h,non_blocking=True))):
        fn(); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); [fn() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
        print(name, gb, "GB:", round(3*gb/(e0.elapsed_time(e1)*1e-3),1), "GB/s")
# both directions at once
h2=torch.empty(int(1.38e9/4),dtype=torch.float32).pin_memory(); d2=torch.empty_like(h2,device=dev)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print("duplex: d2h 3.9GB + h2d 1.38GB x3 in", round(dt*1e3,1), "ms ->", round(3*3.9/dt,1), "GB/s d2h")
