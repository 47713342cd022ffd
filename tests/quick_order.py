#!/usr/bin/env python
"""Does the config-5 step depend on what ran before it in the process?  (116 k - 170 k day-pairs/s
depending on run order.)  Runs the cfg5 step with per-pass timing, optionally after other legs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xcorr_b200.signal import Correlate2dPlan, timelapse_pairs
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for name in [a for a in sys.argv[1:] if a]:
    r = bench.config_leg(name, torch, dev, 3, 2)
    print("before:", name, round(r["ms_per_step"], 3), "ms")
    torch.cuda.empty_cache()
free, total = torch.cuda.mem_get_info()
F, T, D = 1024, 128, 365
plan = Correlate2dPlan(F, T, D, timelapse_pairs(D), dtype=torch.float32, device=dev)
g = torch.Generator(device=dev); g.manual_seed(5)
days = torch.randn((D, F, T), generator=g, device=dev) ** 2
for _ in range(2):
    plan.forward(days); plan.correlate_argmax()
torch.cuda.synchronize()
period = float(os.environ.get("XCB_SAMPLER_MS", "0"))
sampler = None
if period > 0:
    sampler = bench.ClockSampler(0)
    sampler.period = period * 1e-3
    sampler.start()
plan.plan.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    plan.forward(days); plan.correlate_argmax()
e1.record(); torch.cuda.synchronize()
if sampler:
    print("sampler", period, "ms:", sampler.result())
print("cfg5 after", sys.argv[1:], "free GB", round(free / 1e9, 1), "workspace ptr %x" % plan.plan.workspace.data_ptr(), "ms/step",
      round(e0.elapsed_time(e1) / 3, 2), [(k, round(ms / 3, 2)) for k, ms, c in plan.plan.profile_read() if c])
