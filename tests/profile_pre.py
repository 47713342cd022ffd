#!/usr/bin/env python
"""ncu driver: config-2 preprocessing chain on 5 x 24h@250Hz traces."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs
from bench_configs import CFG2_OPS
from xcorr_b200.stream.process import PreprocessProgram
dev = torch.device("cuda:0")
dt = torch.float64 if os.environ.get("XCB_DTYPE", "f32") == "f64" else torch.float32
raw = make_inputs(torch, 5, 21_600_000, dev, dt, 2) + 3.0
prog = PreprocessProgram(CFG2_OPS, 21_600_000, 250.0)
for _ in range(2):
    y = prog.run(raw)
torch.cuda.synchronize()
print(y.shape, float(y.abs().max()))
