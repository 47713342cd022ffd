#!/usr/bin/env python
"""A/B timer: config-2 preprocessing chain (5 x 24h@250Hz) for the library in XCB_LIB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs, CFG2_OPS
from xcorr_b200.stream.process import PreprocessProgram
dev = torch.device("cuda:0")
raw = make_inputs(torch, 5, 21_600_000, dev, torch.float32, 2) + 3.0
ops = {"iir": CFG2_OPS,
       "fir": [op if op[0] != "decimate" else ("resample_poly", {"up": 1, "down": 5}) for op in CFG2_OPS]}
only = os.environ.get("XCB_PRE_ONLY")
for name, oplist in ops.items():
    if only and name != only:
        continue
    prog = PreprocessProgram(oplist, 21_600_000, 250.0)
    for _ in range(3):
        y = prog.run(raw)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        y = prog.run(raw)
    b.record()
    torch.cuda.synchronize()
    print(os.path.basename(os.environ.get("XCB_LIB", "default")), name, f"{a.elapsed_time(b) / 20:.4f} ms",
          tuple(y.shape), f"sum={float(y.double().sum()):.12e} absmax={float(y.abs().max()):.9e}", flush=True)
