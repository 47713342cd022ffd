#!/usr/bin/env python
"""Timeline of a core.process() call of the e2e leg of bench.py (XCB_TIMELINE stamps per batch)."""
import os, sys, time, json
os.environ["XCB_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import xcorr_b200 as xb
import importlib
pm = importlib.import_module("xcorr_b200.core.process")
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
R, n, K = 16, 21_600_000, 8_100_000
x = bench.make_inputs(torch, R, n, dev, torch.float32, 3000)
x_host = [x.cpu().pin_memory()]
del x
def barrier(): torch.cuda.synchronize()
print("pcie", bench.pcie_probe(torch, dev, barrier))
steps = int(os.environ.get("XCB_STEPS", 6))
t0 = time.perf_counter()
ms, ds = bench.e2e_core_process(torch, xb, dev, x_host, R, n, K, steps, 3, min(steps, 6), barrier)
print("e2e ms/step", ms / steps, "total wall", time.perf_counter() - t0)
for tl in pm.last_stats["timeline"]:
    print(json.dumps(tl))
