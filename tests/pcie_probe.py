#!/usr/bin/env python
"""Page-locked host <-> device copy bandwidth: the ceiling of the end-to-end figure (bench.py `e2e`).

    python tests/pcie_probe.py [out.json]                       one GPU
    torchrun --nproc-per-node N tests/pcie_probe.py [out.json]  N ranks copying at once (they share the host)

Per rank: D2H and H2D alone (1 GB and 3.9 GB buffers), and both directions at once with the sizes of one
config-3 window (3.9 GB of CCFs out, 1.38 GB of samples in).  With several ranks the minimum over ranks and the
aggregate are reported."""
import json
import os
import sys
import time

import torch

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


res = {}
for gb in (1.0, 3.9):
    n = int(gb * 1e9 / 4)
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[f"{name}_{gb:g}GB_gbs"] = 3 * gb / (e0.elapsed_time(e1) * 1e-3)
h2 = torch.empty(int(1.3824e9 / 4), dtype=torch.float32, pin_memory=True)
d2 = torch.empty_like(h2, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
barrier()
t0 = time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
res["duplex_window_ms"] = dt * 1e3
res["duplex_d2h_gbs"] = 3.9 / dt
res["duplex_h2d_gbs"] = 1.3824 / dt
keys = sorted(res)
t = torch.tensor([res[k] for k in keys], dtype=torch.float64, device=dev)
if dist is not None:
    mn, sm = t.clone(), t.clone()
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
else:
    mn, sm = t, t
if rank == 0:
    out = {"ranks": world, "gpu": torch.cuda.get_device_name(dev),
           "min_over_ranks": {k: round(float(v), 2) for k, v in zip(keys, mn)},
           "sum_over_ranks": {k: round(float(v), 2) for k, v in zip(keys, sm) if k.endswith("gbs")},
           "note": "page-locked buffers, CUDA events (duplex: wall clock around both streams), 3 repetitions after one warm-up"}
    line = json.dumps(out)
    print(line)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "a") as f:
            f.write(line + "\n")
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
