#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one
B200, device-resident inputs, CUDA-event timing.  `bench.py` is the contract
benchmark (config 3); this script fills in the rest of the measurement rows.

    python bench_configs.py [--configs cfg1,cfg2,cfg3_f64,cfg3_full,cfg4] > profiles/configs_rNN.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import xcorr_b200 as xb  # noqa: E402
from bench import make_inputs, peaks  # noqa: E402
from xcorr_b200.stream.process import PreprocessProgram  # noqa: E402

CFG2_OPS = [("detrend", {"type": "demean"}),
            ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0}),
            ("filter", {"type": "bandpass", "freqmin": 3.0, "freqmax": 12.0}),
            ("decimate", {"factor": 5}),
            ("detrend", {"type": "demean"}),
            ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0})]


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def b_alg(N, P, n, L, K, b, n_raw=0):
    """SURVEY 8d bytes per pair-window: [N(F + R) + P I] / P, with the problem's
    full-window transform length (not the plan's, which may use blocks)."""
    L = 2 * n                         # SURVEY 8 named shapes: L = 2n is smooth
    F = n * b + 3 * L * b
    R = 2 * n_raw * b + n * b if n_raw else 0
    return (N * (F + R) + P * (4 * L * b + K * b)) / P


def run(name, steps, warmup):
    dev = torch.device("cuda:0")
    peak, _ = peaks()
    out = {"config": name, "steps": steps, "warmup": warmup}
    if name in ("cfg1", "cfg1_f32"):
        dt = torch.float64 if name == "cfg1" else torch.float32
        n, R = 4_320_000, 2
        plan = xb.CorrelationPlan(n, R, [(0, 1)], dtype=dt, index_min=n - 1, index_max=n - 1 + 1_620_000, device=dev)
        x = make_inputs(torch, R, n, dev, dt, 1)
        ms = timed(lambda: plan(x), steps, warmup)
        b = 8 if dt == torch.float64 else 4
        out.update(desc="1 pair, 24h@50Hz, lags (0, 9h)", dtype=str(dt), ms_per_step=ms, unit="pair-days/s",
                   value=1 / (ms * 1e-3), passes=plan.dims, blocks=plan.n_blocks,
                   alg_bytes_per_unit=b_alg(2, 1, n, plan.fft_len, 1_620_000, b))
    elif name in ("cfg2", "cfg2_f64"):
        dt = torch.float64 if name.endswith("f64") else torch.float32
        b = 8 if dt == torch.float64 else 4
        n_raw, R = 21_600_000, 5
        pairs = xb.all_pairs(R)
        prog = PreprocessProgram(CFG2_OPS, n_raw, 250.0)
        n = prog.npts_out
        plan = xb.CorrelationPlan(n, R, pairs, dtype=dt, index_min=n - 1, index_max=n - 1 + 1_620_000, device=dev)
        raw = make_inputs(torch, R, n_raw, dev, dt, 2) + 3.0

        def step():
            plan(prog.run(raw))
        ms = timed(step, steps, warmup)
        ms_pre = timed(lambda: prog.run(raw), steps, warmup)
        out.update(desc="5 receivers / 10 pairs, 24h@250Hz raw -> detrend, taper, 3-12 Hz bandpass, decimate 5 -> cc at 50 Hz, lags (0, 9h)",
                   dtype=str(dt), ms_per_step=ms, ms_preprocess=ms_pre, unit="pair-days/s", value=len(pairs) / (ms * 1e-3),
                   passes=plan.dims, blocks=plan.n_blocks, alg_bytes_per_unit=b_alg(R, len(pairs), n, plan.fft_len, 1_620_000, b, n_raw))
    elif name in ("cfg3_f64", "cfg3_full", "cfg3"):
        dt = torch.float64 if name == "cfg3_f64" else torch.float32
        b = 8 if dt == torch.float64 else 4
        n, R = 21_600_000, 16
        pairs = xb.all_pairs(R)
        K = 2 * n - 1 if name == "cfg3_full" else 8_100_000
        i0 = 0 if name == "cfg3_full" else n - 1
        plan = xb.CorrelationPlan(n, R, pairs, dtype=dt, index_min=i0, index_max=i0 + K, device=dev)
        x = make_inputs(torch, R, n, dev, dt, 3)
        ccf = torch.empty((len(pairs), 1, K), dtype=dt, device=dev)

        def step():
            plan.forward(x)
            plan.correlate(out=ccf)
        ms = timed(step, steps, warmup)
        out.update(desc=f"16 receivers / 120 pairs, 24h@250Hz, K={K}", dtype=str(dt), ms_per_step=ms,
                   unit="pair-days/s", value=len(pairs) / (ms * 1e-3), passes=plan.dims, blocks=plan.n_blocks,
                   alg_bytes_per_unit=b_alg(R, len(pairs), n, plan.fft_len, K, b))
    elif name in ("cfg4", "cfg4_f64"):
        dt = torch.float64 if name.endswith("f64") else torch.float32
        b = 8 if dt == torch.float64 else 4
        n, R, W = 900_000, 4, 256
        pairs = xb.all_pairs(R)
        plan = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=dt, stack=True, device=dev)
        K = plan.n_lags
        x = torch.stack([make_inputs(torch, R, n, dev, dt, 100 + w) for w in range(8)], dim=1)   # [R, 8, n]
        x = x.repeat(1, W // 8, 1).contiguous()
        stack = torch.zeros((len(pairs), K), dtype=dt, device=dev)

        def step():
            plan.forward(x)
            plan.correlate(out=stack)
        ms = timed(step, steps, warmup)
        L = plan.fft_len
        # spectral stacking: I = 2 L b + (2 L b + K b) / W per pair-window
        alg = (R * (n * b + 3 * L * b) + len(pairs) * (2 * L * b + (2 * L * b + K * b) / W)) / len(pairs)
        out.update(desc=f"4 receivers / 6 pairs, {W} windows of 1h@250Hz per step, stacked on device (spectral stacking)",
                   dtype=str(dt), ms_per_step=ms, unit="pair-hours/s", value=len(pairs) * W / (ms * 1e-3),
                   passes=plan.dims, blocks=plan.n_blocks, alg_bytes_per_unit=alg)
    elif name in ("cfg5", "cfg5_f64"):
        from xcorr_b200.signal import Correlate2dPlan
        dt = torch.float64 if name.endswith("f64") else torch.float32
        b = 8 if dt == torch.float64 else 4
        F, T, D = 1024, 128, 48
        pairs = [(i, j) for i in range(D) for j in range(i, D)]          # lower triangle incl. diagonal
        plan = Correlate2dPlan(F, T, D, pairs, dtype=dt, device=dev)
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        days = torch.randn((D, F, T), generator=g, device=dev, dtype=dt) ** 2

        def step():
            cc2 = plan(days)
            plan.argmax(cc2)
        ms = timed(step, steps, warmup)
        A = 2 * plan.plan.spectrum_elems * b                              # bytes of one padded spectrum
        alg = (D * (F * T * b + 3 * A) + len(pairs) * 4 * A) / len(pairs)  # F = FTb + 3A, I = 4A
        out.update(desc=f"{D} day spectrograms 1024x128 -> {len(pairs)} day pairs, cc2 + arg-max (timelapse kernel)",
                   dtype=str(dt), ms_per_step=ms, unit="day-pairs/s", value=len(pairs) / (ms * 1e-3),
                   passes=plan.plan.dims, blocks=plan.plan.n_blocks, alg_bytes_per_unit=alg)
    else:
        raise SystemExit(f"unknown config {name}")
    out["step_roofline_frac"] = out["alg_bytes_per_unit"] * out["value"] / 1e9 / peak
    out["peak_gbs"] = peak
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg1,cfg1_f32,cfg2,cfg2_f64,cfg3,cfg3_f64,cfg3_full,cfg4,cfg4_f64,cfg5,cfg5_f64")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    for name in args.configs.split(","):
        res = run(name, args.steps, args.warmup)
        print(json.dumps(res), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
