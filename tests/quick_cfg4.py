#!/usr/bin/env python
"""Per-pass timing of the config-4 step (4 receivers, 256 stacked 1 h windows) for candidate factorizations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xcorr_b200 as xb
from bench import make_inputs
dev = torch.device("cuda:0")
n, R, W = 900_000, 4, 256
pairs = xb.all_pairs(R)
x = torch.stack([make_inputs(torch, R, n, dev, torch.float32, 100 + w) for w in range(8)], dim=1).repeat(1, W // 8, 1).contiguous()
for dims in [None] + [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]:
    try:
        plan = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=torch.float32, stack=True, device=dev, dims=dims)
    except Exception as e:
        print(dims, "failed", e); continue
    stack = torch.zeros((len(pairs), plan.n_lags), dtype=torch.float32, device=dev)
    for _ in range(2):
        stack.zero_(); plan.forward(x); plan.correlate(out=stack)
    torch.cuda.synchronize()
    plan.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        stack.zero_(); plan.forward(x); plan.correlate(out=stack)
    e1.record(); torch.cuda.synchronize()
    print("dims", plan.dims, "groups", plan.n_groups, "ms/step", round(e0.elapsed_time(e1) / 3, 2),
          [(k, c // 3, round(ms / 3, 2)) for k, ms, c in plan.profile_read() if c])
    del plan, stack
    torch.cuda.empty_cache()
