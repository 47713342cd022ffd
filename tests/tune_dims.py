#!/usr/bin/env python
"""Sweep pass factorisations for one transform length on the GPU and print the
per-kernel times (tuning aid; not a test)."""
import itertools
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xcorr_b200 as xb
from bench import make_inputs

MENU = [64, 80, 96, 100, 120, 128, 144, 150, 160, 180, 192, 200, 225, 240, 256, 300, 320, 400]
n = int(os.environ.get("XCB_N", 21_600_000))
R = int(os.environ.get("XCB_R", 16))
P = int(os.environ.get("XCB_P", 24))
dtype = torch.float64 if os.environ.get("XCB_DTYPE", "f32") == "f64" else torch.float32
slack = float(os.environ.get("XCB_SLACK", 1.0))
cands = set()
for combo in itertools.combinations_with_replacement(MENU, 3):
    m = combo[0] * combo[1] * combo[2]
    if n <= m <= n * slack:
        for perm in set(itertools.permutations(combo)):
            cands.add(perm)
if os.environ.get("XCB_DIMS"):
    cands = {tuple(int(v) for v in d.split(",")) for d in os.environ["XCB_DIMS"].split(";")}
if os.environ.get("XCB_MRANGE"):                 # all factorisations with product in [lo, hi], row axis <= 256
    lo, hi = (float(v) for v in os.environ["XCB_MRANGE"].split(","))
    cands = set()
    for combo in itertools.combinations_with_replacement(MENU, 3):
        m = combo[0] * combo[1] * combo[2]
        if lo <= m <= hi:
            for perm in set(itertools.permutations(combo)):
                if perm[2] <= 256 and perm[2] % 16 == 0 and perm[0] >= perm[1]:
                    cands.add(perm)
NB = int(os.environ.get("XCB_NBLOCKS", 1))       # block mode: dims give the short transform
LB = int(os.environ.get("XCB_BLOCK_LEN", 0))
KL = int(os.environ.get("XCB_LAGS", 3 * n // 8))
dev = torch.device("cuda:0")
x = make_inputs(torch, R, n, dev, dtype, seed=1)
pairs = xb.all_pairs(R)[:P]
rows = []
for dims in sorted(cands):
    try:
        lb = LB
        if NB > 1 and not lb:
            lb = -(-n // NB)
            row = 2 * dims[1] * dims[2]
            if n % row == 0 and -(-lb // row) * row + KL <= 2 * dims[0] * dims[1] * dims[2]:
                lb = -(-lb // row) * row
            elif -(-lb // 32) * 32 + KL <= 2 * dims[0] * dims[1] * dims[2]:
                lb = -(-lb // 32) * 32
        plan = xb.CorrelationPlan(n, R, pairs, dtype=dtype, index_min=n - 1, index_max=n - 1 + KL, dims=dims,
                                  n_blocks=NB, block_len=lb)
    except Exception as e:
        print(dims, "skip", e)
        continue
    for _ in range(2):
        ccf = plan(x)
    torch.cuda.synchronize()
    plan.profile(True)
    for _ in range(3):
        ccf = plan(x)
    torch.cuda.synchronize()
    k = plan.profile_read()
    fwd = sum(ms for name, ms, c in k if name.startswith("fwd")) / 3      # totals per iteration
    inv = sum(ms for name, ms, c in k if name.startswith("inv")) / 3
    idx, _ = plan.argmax(ccf)
    ok = idx[0, 0].item() == 137
    rows.append((fwd + inv * 120 / P, dims, fwd, inv, ok, [round(ms / 3, 3) for _, ms, c in k]))
    del plan, ccf
    torch.cuda.empty_cache()
rows.sort()
for est, dims, fwd, inv, ok, ks in rows:
    print(f"{dims}  est step {est:7.2f} ms  fwd {fwd:6.2f}  inv/{P} {inv:6.2f}  ok={ok}  {ks}")
