#!/usr/bin/env python
"""Where does the FIRST batch of a core.process() call spend its host time?  (cProfile of launch())"""
import os, sys, time, json, cProfile, pstats
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
steps = int(os.environ.get("XCB_STEPS", 3))
pr = cProfile.Profile()
ms, ds = bench.e2e_core_process(torch, xb, dev, x_host, R, n, K, steps, 3, steps, barrier)
pr.enable()
ms, ds = bench.e2e_core_process(torch, xb, dev, x_host, R, n, K, steps, 3, steps, barrier)
pr.disable()
print("e2e ms/step", ms / steps)
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
