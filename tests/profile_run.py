#!/usr/bin/env python
"""Short driver for ncu: one 24h@250Hz window, 16 receivers, all 120 pairs,
lags (0, 9h), run twice.  With the automatic plan (2 blocks, passes 200x200x240)
the launch order per iteration is: fwd pass 0 x4 (segments / blocks), fwd pass 1,
fwd pass 2 (row), sumsq_finish, rnorm, inv pass 2 (row, pair product, tile major),
inv pass 1, inv pass 0 (lag epilogue)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xcorr_b200 as xb
from bench import make_inputs

n = int(os.environ.get("XCB_N", 21_600_000))
R = int(os.environ.get("XCB_R", 16))
dtype = torch.float64 if os.environ.get("XCB_DTYPE", "f32") == "f64" else torch.float32
dims = tuple(int(v) for v in os.environ["XCB_DIMS"].split(",")) if os.environ.get("XCB_DIMS") else None
pairs = xb.all_pairs(R)[:int(os.environ.get("XCB_P", 120))]
dev = torch.device("cuda:0")
plan = xb.CorrelationPlan(n, R, pairs, dtype=dtype, index_min=n - 1, index_max=n - 1 + 3 * n // 8, dims=dims,
                          n_blocks=int(os.environ.get("XCB_NBLOCKS", 0)), block_len=int(os.environ.get("XCB_BLOCK_LEN", 0)))
x = make_inputs(torch, R, n, dev, dtype, seed=1)
for it in range(int(os.environ.get("XCB_ITERS", 2))):
    ccf = plan(x)
torch.cuda.synchronize()
idx, _ = plan.argmax(ccf)
print("dims", plan.dims, "argmax", idx[:3, 0].tolist())
