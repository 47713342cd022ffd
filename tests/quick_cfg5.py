#!/usr/bin/env python
"""Per-pass timing of the config-5 step (365 days, 67 159 day pairs, fused arg-max) on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xcorr_b200.signal import Correlate2dPlan, timelapse_pairs
dev = torch.device("cuda:0")
F, T, D = 1024, 128, int(os.environ.get("XCB_DAYS", 365))
pairs = timelapse_pairs(D)
plan = Correlate2dPlan(F, T, D, pairs, dtype=torch.float32, device=dev,
                       max_work_bytes=int(float(os.environ.get('XCB_WORK_MB', 0)) * 2**20))
g = torch.Generator(device=dev); g.manual_seed(5)
days = torch.randn((D, F, T), generator=g, device=dev) ** 2
for _ in range(2):
    plan.forward(days); plan.correlate_argmax()
torch.cuda.synchronize()
plan.plan.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    plan.forward(days); plan.correlate_argmax()
e1.record(); torch.cuda.synchronize()
print("dims", plan.plan.dims, "groups", plan.plan.n_groups, "jobs in flight", plan.plan.jobs_in_flight, "ms/step",
      round(e0.elapsed_time(e1) / 3, 2), [(k, c // 3, round(ms / 3, 2)) for k, ms, c in plan.plan.profile_read() if c])
