#!/usr/bin/env python
"""Benchmark of the cross-correlation hot path (BASELINE.json metric:
pair-days/s for 24 h @ 250 Hz CCFs, and % of the HBM roofline).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload = "cfg3", BASELINE.json configs[2], the configuration
the metric is quoted on; it fits one GPU): all 120 pairs of a 16-receiver array,
one 24 h window at 250 Hz per step (n = 21.6 M samples, L = 43.2 M), float32,
lags cropped to (0, 9 h) = 8.1 M samples.  A step = 16 forward transforms + 120
pair correlations = 120 pair-days.  Multi-GPU: every rank processes its own
windows (weak scaling, no data-path collective; NCCL only for the barrier and
the max-over-ranks reduction of the timing).

Printed JSON line:
  value         device-resident throughput (inputs in HBM when the timed region starts)
  e2e           the same metric through the reference-facing driver `xcorr_b200.core.process`
                (core/process.py:82-165 replacement): page-locked HOST windows handed over by
                the waveform callback, H2D / compute / D2H into the dataset's `cc` array inside
                the timed region, against the PCIe ceiling measured in the same run
  roofline      dominant kernel vs the measured HBM peak (traffic: ncu capture under profiles/)
  step_roofline whole step vs SURVEY 8d's minimal-traffic model
  configs       short legs of the other named configurations (cfg1, cfg3/f64, cfg2 (+ its FIR-decimator variant), cfg4, cfg5)
  cpu_baseline  the reference's own xcorr/util/cc.py (oracle/_ref) on the host cores, FULL-SIZE windows

`--impl reference` times the reference's util/cc.py alone (full-size 24 h @ 250 Hz windows, one
single-threaded task per core like core/lazy.py) under a wall budget.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: receivers, n, kept lags (from zero lag), preprocess
    "cfg3": dict(receivers=16, n=21_600_000, lags=8_100_000,
                 desc="16 receivers / 120 pairs, 24h@250Hz, lags (0, 9h), one window per step"),
    "cfg3-full": dict(receivers=16, n=21_600_000, lags=0,
                      desc="16 receivers / 120 pairs, 24h@250Hz, all 2n-1 lags"),
    "cfg1": dict(receivers=2, n=4_320_000, lags=1_620_000,
                 desc="1 pair, 24h@50Hz, lags (0, 9h)"),
    "tiny": dict(receivers=4, n=200_000, lags=50_000, desc="4 receivers, n=200k (plumbing check)"),
}
REF_BUDGET_S = float(os.environ.get("XCB_REF_BUDGET_S", "200"))     # wall budget of the reference arm
REF_GB_PER_TASK = 8.5            # peak RSS of one full-size util/cc.py call (measured: 7.5 GB at n = 21.6 M, float32)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------
# CPU reference arm: the reference's own xcorr/util/cc.py (oracle/_ref, placed there by
# oracle/build_ref.py) on the host cores, one single-threaded task per pair-window like
# core/lazy.py does.  Falls back to the oracle port when oracle/_ref was never built.
# ---------------------------------------------------------------------------
def _cpu_task(args):
    n, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import build_ref, cc_oracle as orc
    ref = build_ref.load()
    cc = ref.cc if ref is not None else orc.cc
    rec = orc.synthetic_receivers(2, n, seed=seed, delay=137).astype(np.float32)
    t0 = time.perf_counter()
    # core/process.py:155-162: cc(x, y, normalize, pad=True, unbiased=False, dtype=dataset.cc.dtype)
    r = cc(rec[0], rec[1], normalize=True, pad=True, unbiased=False, dtype=np.float32)
    dt = time.perf_counter() - t0
    return dt, int(np.argmax(r)) - (n - 1)


def _usable_cores(n):
    cores = os.cpu_count() or 1
    try:
        cores = min(cores, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    avail = None
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    avail = int(line.split()[1]) * 1024
        for cg in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
            if os.path.exists(cg):
                v = open(cg).read().strip()
                if v.isdigit():
                    avail = min(avail, int(v)) if avail else int(v)
    except Exception:
        pass
    if avail:
        per = REF_GB_PER_TASK * 1e9 * n / 21_600_000
        cores = max(1, min(cores, int(0.8 * avail / per)))
    return cores


def cpu_reference(n, steps, budget_s=REF_BUDGET_S, cores=None):
    """Each step: `cores` independent FULL-SIZE pair-windows of n samples, one per core, through the
    reference's util/cc.py.  At least one step runs; further steps (up to `steps`) only while the wall
    budget allows, and the steps actually run are reported."""
    import multiprocessing as mp
    from oracle import build_ref
    kind = "reference" if build_ref.load() is not None else "port"
    cores = cores or _usable_cores(n)
    ctx = mp.get_context("fork")
    lags, done, per_step = [], 0, []
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_task, [(1000, c) for c in range(cores)])          # start the workers (imports), tiny job
        t_all = time.perf_counter()
        while done < max(1, steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_task, [(n, 20_000 + done * cores + c) for c in range(cores)])
            per_step.append(time.perf_counter() - t0)
            lags += [r[1] for r in res]
            done += 1
            if time.perf_counter() - t_all + 1.1 * per_step[-1] > budget_s:
                break
        wall = time.perf_counter() - t_all
    assert all(l == 137 for l in lags), "CPU reference lost the synthetic arrival"
    src = "xcorr/util/cc.py executed from oracle/_ref (byte-identical copy)" if kind == "reference" else \
        "oracle port of xcorr/util/cc.py (oracle/_ref not built)"
    return dict(value=done * cores / wall, wall=wall, cores=cores, n=n, steps_run=done, kind=kind,
                per_step_s=per_step,
                sample=f"{done} step(s) x {cores} single-threaded FULL-SIZE pair-windows (n={n}, float32, 2n-1 lags) "
                       f"on {cores} cores, {wall:.1f} s wall; {src}, numpy pocketfft")


def measured_traffic(workload, dtype, plan, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, read from the table
    tests/ncu_traffic.py wrote out of the committed `ncu --set full` capture (profiles/ncu_traffic.json);
    None when that capture did not run this plan."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        tab = json.load(f)
    key = f"{workload}/{dtype}/{'x'.join(str(d) for d in plan.dims)}/b{plan.n_blocks}"
    ent = tab.get(key, {}).get(kernel)
    return (ent, tab.get("_source")) if ent is not None else (None, None)


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.handle = None
        self.period = 0.005
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.handle = None

    def run(self):
        if self.handle is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def result(self):
        self.stop_flag.set()
        if self.handle is None or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def make_inputs(torch, receivers, n, device, dtype, seed):
    """SURVEY 8d synthetic receivers generated on the device: common source
    delayed by 137*r samples plus 0.5 sigma independent noise."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    src = torch.randn(n + 137 * receivers + 8, generator=g, device=device, dtype=dtype)
    x = torch.empty((receivers, n), device=device, dtype=dtype)
    for r in range(receivers):
        x[r] = src[137 * r:137 * r + n]
        x[r] += 0.5 * torch.randn(n, generator=g, device=device, dtype=dtype)
    return x


def pcie_probe(torch, dev, barrier, gb=1.0):
    """Page-locked H2D / D2H bandwidth of this rank while ALL ranks copy at once (the ceiling the
    end-to-end figure is stated against)."""
    n = int(gb * 1e9 / 4)
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d = torch.empty(n, dtype=torch.float32, device=dev)
    out = {}
    for name, fn in (("d2h_gbs", lambda: h.copy_(d, non_blocking=True)), ("h2d_gbs", lambda: d.copy_(h, non_blocking=True))):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = 3 * gb / (e0.elapsed_time(e1) * 1e-3)
    return out


def e2e_core_process(torch, xb, dev, x_host_windows, R, n, K, steps, warmup, windows_per_call, barrier):
    """`steps` windows through xcorr_b200.core.process (the drop-in of core/process.py:82-165): the
    waveform callback hands over page-locked host arrays, the driver copies them in, correlates all pairs
    and copies the CCFs into the dataset's (page-locked) `cc` variable.  Returns (ms, dataset)."""
    import pandas as pd
    from xcorr_b200 import core
    recv = [f"XX.R{r:02d}..HHZ" for r in range(R)]
    pairs = [f"{recv[i]}-{recv[j]}" for i in range(R) for j in range(i + 1, R)]
    fs = 250.0
    wl = n / fs
    Wc = windows_per_call
    t0 = pd.Timestamp("2020-01-01")
    ds = core.init(pairs, t0, t0 + pd.Timedelta(seconds=wl * (Wc - 1)), {"HHZ": []},
                   dict(institution="bench", author="bench", source="synthetic"), fs, window_length=wl,
                   window_overlap=0.0, clip_lag=(0.0, K / fs), closed=None, dtype=np.float32)
    assert ds.variables["cc"].data.shape == (len(pairs), Wc, K), ds.variables["cc"].data.shape
    times = list(pd.DatetimeIndex(ds.coords["time"].data))
    nw = len(x_host_windows)

    def provider(receiver, t):
        w = times.index(t)
        return x_host_windows[w % nw][int(receiver[4:6])].numpy(), fs, 0.0

    def run(k):
        done = 0
        while done < k:
            ds.variables["status"].data[...] = 0
            ds.variables["status"].data[:, min(Wc, k - done):] = 1          # windows beyond this call: left alone
            core.process(ds, provider, windows_per_batch=1, device=dev, verb=0)
            done += min(Wc, k - done)

    run(max(2, min(warmup, Wc)))
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(steps)
    b.record()
    barrier()
    if os.environ.get("XCB_TIMELINE"):
        import importlib
        for tl in importlib.import_module("xcorr_b200.core.process").last_stats.get("timeline", []):
            print("timeline", json.dumps(tl), file=sys.stderr)
    return a.elapsed_time(b), ds


CFG2_OPS = [("detrend", {"type": "demean"}),
            ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0}),
            ("filter", {"type": "bandpass", "freqmin": 3.0, "freqmax": 12.0}),
            ("decimate", {"factor": 5}),
            ("detrend", {"type": "demean"}),
            ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0})]


def b_alg(N, P, n, K, b, n_raw=0):
    """SURVEY 8d bytes per pair-window: [N(F + R) + P I] / P with the problem's full-window
    transform length L = 2n (not the plan's, which may use blocks)."""
    L = 2 * n
    F = n * b + 3 * L * b
    R = 2 * n_raw * b + n * b if n_raw else 0
    return (N * (F + R) + P * (4 * L * b + K * b)) / P


def config_leg(name, torch, dev, steps, warmup, rank=0, world=1, dist=None):
    """One short device-resident leg of another named configuration.  Returns a dict with value, unit,
    step_roofline_frac and the clocks sampled during its timed region.  Every rank runs the leg on its own
    windows (weak scaling); cfg4 adds the NCCL all-reduce of the stacks, cfg5 shards the day transforms and
    all-gathers the half spectra."""
    import xcorr_b200 as xb
    from xcorr_b200.stream.process import PreprocessProgram
    peak, _ = peaks()
    out = {"workload": name}
    extra = {}
    sub_timers = {}
    if name in ("cfg3_f64", "cfg3_full"):
        dt = torch.float64 if name == "cfg3_f64" else torch.float32
        b = 8 if dt == torch.float64 else 4
        n, R = 21_600_000, 16
        pairs = xb.all_pairs(R)
        K = 2 * n - 1 if name == "cfg3_full" else 8_100_000
        i0 = 0 if name == "cfg3_full" else n - 1
        plan = xb.CorrelationPlan(n, R, pairs, dtype=dt, index_min=i0, index_max=i0 + K, device=dev)
        x = make_inputs(torch, R, n, dev, dt, 3 + rank)
        ccf = torch.empty((len(pairs), 1, K), dtype=dt, device=dev)

        def step():
            plan.forward(x)
            plan.correlate(out=ccf)
        units, unit = len(pairs), "pair-days/s"
        alg = b_alg(R, len(pairs), n, K, b)
        extra = dict(passes=list(plan.dims), blocks=plan.n_blocks)
        desc = f"16 receivers / 120 pairs, 24h@250Hz, K={K}"
    elif name in ("cfg1", "cfg1_f32"):
        dt = torch.float64 if name == "cfg1" else torch.float32
        b = 8 if dt == torch.float64 else 4
        n, R, K = 4_320_000, 2, 1_620_000
        pairs = [(0, 1)]
        plan = xb.CorrelationPlan(n, R, pairs, dtype=dt, index_min=n - 1, index_max=n - 1 + K, device=dev)
        x = make_inputs(torch, R, n, dev, dt, 1 + rank)
        step, _ = plan.graphed(x)                   # ten ~15 us kernels: one graph launch instead of ten launches
        units, unit = 1, "pair-days/s"
        alg = b_alg(R, 1, n, K, b)
        extra = dict(passes=list(plan.dims), blocks=plan.n_blocks, launch="CUDA graph replay of forward + correlate")
        desc = "1 pair, 24h@50Hz, lags (0, 9h) (the reference's CPU-runnable case)"
    elif name in ("cfg2", "cfg2_f64", "cfg2_fir"):
        dt = torch.float64 if name.endswith("f64") else torch.float32
        b = 8 if dt == torch.float64 else 4
        n_raw, R = 21_600_000, 5
        pairs = xb.all_pairs(R)
        # cfg2_fir: the same chain with the north-star decimator (polyphase FIR, scipy.signal.resample_poly
        # semantics) in place of obspy's Chebyshev-II `decimate`
        ops = CFG2_OPS if name != "cfg2_fir" else [
            op if op[0] != "decimate" else ("resample_poly", {"up": 1, "down": 5}) for op in CFG2_OPS]
        prog = PreprocessProgram(ops, n_raw, 250.0)
        n = prog.npts_out
        plan = xb.CorrelationPlan(n, R, pairs, dtype=dt, index_min=n - 1, index_max=n - 1 + 1_620_000, device=dev)
        raw = make_inputs(torch, R, n_raw, dev, dt, 2 + rank) + 3.0

        def step():
            plan(prog.run(raw))
        sub_timers["ms_preprocess"] = lambda: prog.run(raw)
        units, unit = len(pairs), "pair-days/s"
        alg = b_alg(R, len(pairs), n, 1_620_000, b, n_raw)
        # the chain is bound by FP64 arithmetic, not by bytes: 10 second-order sections x 5 multiply-adds per raw
        # sample in float64 (exact recurrence, scipy.signal.sosfilt order) -- reported beside the byte roofline
        n_sec, fir_fma = (10, 0) if name != "cfg2_fir" else (4, R * n * 101)
        extra = dict(passes=list(plan.dims), blocks=plan.n_blocks,
                     preprocess_fp64_fma_per_step=int(R * n_raw * n_sec * 5 + fir_fma),
                     note="step_roofline_frac counts bytes only; the preprocessing chain (ms_preprocess) is FP64-throughput "
                          "bound: fp64_fma_per_step / ms_preprocess is its achieved multiply-add rate")
        desc = "5 receivers / 10 pairs, 24h@250Hz raw -> detrend, taper, 3-12 Hz bandpass, decimate 5 -> cc at 50 Hz, lags (0, 9h)"
        if name == "cfg2_fir":
            desc = desc.replace("decimate 5", "polyphase-FIR decimation by 5 (resample_poly, 101 taps)")
    elif name in ("cfg4", "cfg4_f64"):
        dt = torch.float64 if name.endswith("f64") else torch.float32
        b = 8 if dt == torch.float64 else 4
        n, R, W = 900_000, 4, 256
        pairs = xb.all_pairs(R)
        plan = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=dt, stack=True, device=dev)
        K = plan.n_lags
        x = torch.stack([make_inputs(torch, R, n, dev, dt, 100 + w + 8 * rank) for w in range(8)], dim=1)   # [R, 8, n]
        x = x.repeat(1, W // 8, 1).contiguous()
        stack = torch.zeros((len(pairs), K), dtype=dt, device=dev)

        def step():
            stack.zero_()
            plan.forward(x)
            plan.correlate(out=stack)
            if world > 1:                           # all_reduce(sum) of the P x K stack
                dist.all_reduce(stack, op=dist.ReduceOp.SUM)
        units, unit = len(pairs) * W, "pair-hours/s"
        L = 2 * n
        # spectral stacking: I = 2 L b + (2 L b + K b) / W per pair-window
        alg = (R * (n * b + 3 * L * b) + len(pairs) * (2 * L * b + (2 * L * b + K * b) / W)) / len(pairs)
        extra = dict(passes=list(plan.dims), windows_per_step=W,
                     collective="all_reduce of the stacks (NCCL)" if world > 1 else None)
        desc = f"4 receivers / 6 pairs, {W} windows of 1h@250Hz per step, stacked on device (spectral stacking)"
    elif name in ("cfg5", "cfg5_f64"):
        from xcorr_b200.signal import Correlate2dPlan, timelapse_pairs
        dt = torch.float64 if name.endswith("f64") else torch.float32
        b = 8 if dt == torch.float64 else 4
        F, T, D = 1024, 128, 365
        # the masked triangle of scripts/timelapse.py:348-360, tiled over the ranks; every rank transforms
        # its share of the days, the half spectra are all-gathered once per step
        pairs = timelapse_pairs(D)
        per_rank = -(-len(pairs) // world)                    # contiguous, equal shares of the (i-major) triangle: a rank's
        mine = pairs[rank * per_rank:(rank + 1) * per_rank]   # day pairs share days, which the staged pair product needs
        plan = Correlate2dPlan(F, T, D, mine, dtype=dt, device=dev, day_shard=(rank, world) if world > 1 else None)
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        days = torch.randn((D, F, T), generator=g, device=dev, dtype=dt) ** 2
        my_days = days[plan.day_slice]

        def step():
            plan.forward(my_days)
            plan.correlate_argmax()
        units, unit = len(pairs) / world, "day-pairs/s"      # per rank on average; x world below
        A = 2 * plan.plan.fft_len // 2 * b                                  # bytes of one padded half spectrum
        alg = (D * (F * T * b + 3 * A) + len(pairs) * 4 * A) / len(pairs)   # F = FTb + 3A, I = 4A
        extra = dict(passes=list(plan.plan.dims), days=D, day_pairs=len(pairs),
                     collective="all_gather of the day half-spectra (NCCL)" if world > 1 else None,
                     scaling="strong")
        if world > 1:
            sub_timers["ms_all_gather"] = plan.all_gather      # the collective alone (day half-spectra, NCCL)
            extra["all_gather_bytes_received_per_rank"] = int(plan.plan.spec.numel() * plan.plan.spec.element_size())
        desc = f"{D} day spectrograms 1024x128 -> {len(pairs)} day pairs (masked triangle), value + (dF, dT) per pair (timelapse kernel)"
    else:
        raise SystemExit(f"unknown config {name}")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 2)):
        step()
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    clocks = sampler.result()
    for key, fn in sub_timers.items():               # parts of the step, timed on their own afterwards
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        a.record()
        for _ in range(steps):
            fn()
        b2.record()
        torch.cuda.synchronize()
        extra[key] = a.elapsed_time(b2) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    value = world * units / (ms * 1e-3)
    out.update(description=desc, dtype="f64" if dt == torch.float64 else "f32", value=value, unit=unit, ms_per_step=ms,
               steps=steps, alg_bytes_per_unit=alg, step_roofline_frac=alg * value / world / 1e9 / peak,
               clocks=clocks, **extra)
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import xcorr_b200 as xb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = WORKLOADS[args.workload]
    R, n = wl["receivers"], wl["n"]
    K = wl["lags"] or (2 * n - 1)
    i0, i1 = ((n - 1, n - 1 + K) if wl["lags"] else (0, 2 * n - 1))
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    b = 4 if args.dtype == "f32" else 8
    pairs = xb.all_pairs(R)
    P = len(pairs)
    plan = xb.CorrelationPlan(n, R, pairs, n_windows=1, dtype=tdt, index_min=i0, index_max=i1, device=dev)
    L, M = plan.fft_len, plan.spectrum_elems
    nd = len(plan.dims)
    chunks = (plan.n_jobs + plan.jobs_in_flight - 1) // plan.jobs_in_flight
    fwd0_launches = (plan.spectrum_elems // (plan.fft_len // 2)) if plan.n_blocks > 1 else 1
    # forward: pass 0 per segment/block + the other passes + norms; inverse: 1/||x|| + nd passes per chunk
    launches_per_step = fwd0_launches + (nd - 1) + 1 + 1 + nd * chunks

    x = make_inputs(torch, R, n, dev, tdt, seed=3000 + rank)
    ccf = torch.empty((P, 1, K), dtype=tdt, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    for _ in range(args.warmup):
        plan.forward(x)
        plan.correlate(out=ccf)
    barrier()
    # parity spot-check of what is being timed: arrival of every pair (tests/test_named_configs_gpu.py compares
    # 8 cells of this very plan in full with the oracle)
    idx, _ = plan.argmax(ccf)
    exp = torch.tensor([137 * (j - i) + (n - 1 - i0) for i, j in pairs], device=dev)
    assert bool((idx[:, 0] == exp).all()), "benchmark output lost the synthetic arrival"

    sampler = ClockSampler(local)
    sampler.start()
    plan.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        plan.forward(x)
        plan.correlate(out=ccf)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.result()
    kinds = plan.profile_read()
    plan.profile(False)
    plan_info = dict(dims=list(plan.dims), n_blocks=plan.n_blocks, block_len=plan.block_len, n_groups=plan.n_groups,
                     n_jobs=plan.n_jobs, spectrum_elems=plan.spectrum_elems, fft_len=plan.fft_len)

    class _P:               # what the bookkeeping below needs once the plan's buffers are released
        dims, n_blocks, block_len = tuple(plan.dims), plan.n_blocks, plan.block_len
        spectrum_elems, fft_len, n_jobs = plan.spectrum_elems, plan.fft_len, plan.n_jobs
    pinfo = _P
    x_host = [x.cpu().pin_memory()]
    del plan, ccf, idx, exp
    torch.cuda.empty_cache()

    # ---------------- end to end through core.process, against the measured PCIe ceiling ----------------
    pcie = pcie_probe(torch, dev, barrier)
    # windows per process() call (a dataset of that many 24 h windows): 3.9 GB of page-locked `cc` each, so
    # 10 windows on one rank, fewer when several ranks share the host; every call pays one pipeline ramp
    # (first H2D + first compute, ~42 ms, before its first D2H can start)
    wpc = max(1, min(args.steps, max(2, 10 // world)))
    e2e_ms, ds = e2e_core_process(torch, xb, dev, x_host, R, n, K, args.steps, args.warmup, wpc, barrier)
    got = ds.variables["cc"].data
    assert int(np.argmax(got[0, 0])) == (n - 1 - i0) + 137 and ds.variables["status"].data.min() == 1
    del ds, got, x_host
    from xcorr_b200.core import process as _proc_mod   # noqa: F401  (function; the module holds the caches)
    import importlib
    importlib.import_module("xcorr_b200.core.process").clear_caches()   # plans / staging of the e2e leg: 25 GB of HBM
    torch.cuda.empty_cache()

    # ---------------- the other named configurations (short legs) ----------------
    legs = []
    if not args.no_legs and args.workload == "cfg3" and args.dtype == "f32":
        for name in ("cfg1", "cfg3_f64", "cfg2", "cfg2_fir", "cfg4", "cfg5"):
            try:
                legs.append(config_leg(name, torch, dev, 3, 2, rank, world, dist))
            except Exception as error:                      # a leg must not take the headline down
                legs.append({"workload": name, "error": f"{type(error).__name__}: {error}"[:300]})
            torch.cuda.empty_cache()

    # ---------------- reduce over ranks ----------------
    tmax = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    tmin = torch.tensor([pcie["h2d_gbs"], pcie["d2h_gbs"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    dev_ms, e2e_ms = float(tmax[0]), float(tmax[1])
    pcie = {"h2d_gbs": float(tmin[0]), "d2h_gbs": float(tmin[1])}
    units = world * args.steps * P
    value = units / (dev_ms * 1e-3)
    e2e_value = units / (e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        # per-launch algorithmic bytes of each pass kind: inputs read once + outputs written
        # once, for the transforms the plan actually runs (F = plan.fft_len; block mode: 2B spectra per signal)
        F, B = L, pinfo.n_blocks
        per = pinfo.spectrum_elems // (F // 2)          # spectra per signal (2B, or 2B-1 when the last segment aliases)
        if B > 1:
            lb = pinfo.block_len
            nseg = per - B
            valid = sum(min(F, n - k * lb) for k in range(nseg)) + sum(min(lb, n - k * lb) for k in range(B))
        else:
            valid = n
        per_step = {}
        for a in range(nd):
            rd = R * valid * b if a == 0 else R * per * F * b
            per_step[f"fwd_pass{a}_N{pinfo.dims[a]}"] = rd + R * per * F * b
            if a == nd - 1:
                # pair product: every distinct spectrum of the window is read from HBM once per launch
                # (the pairs that share it are served by L2 / the staged shared-memory rows), one packed
                # spectrum written per pair
                bytes_inv = R * per * F * b + pinfo.n_jobs * (F * b if nd > 1 else K * b)
            elif a == 0:
                bytes_inv = (F * b + K * b) * pinfo.n_jobs
            else:
                bytes_inv = 2 * F * b * pinfo.n_jobs
            per_step[f"inv_pass{a}_N{pinfo.dims[a]}"] = bytes_inv
        table = []
        for name, ms, cnt in kinds:
            if cnt == 0:
                continue
            total_bytes = per_step[name] * args.steps
            table.append(dict(kernel=name, launches=cnt, total_ms=ms, avg_ms=ms / cnt,
                              alg_bytes_per_launch=total_bytes / cnt,
                              achieved_gbs=total_bytes / (ms * 1e-3) / 1e9))
        dom = max(table, key=lambda t: t["total_ms"])
        traffic, traffic_src = measured_traffic(args.workload, args.dtype, pinfo, dom["kernel"])
        # step model of SURVEY 8d: a property of the PROBLEM (full-window transform length L_alg, two HBM
        # passes per transform), independent of how the plan factorises or blocks the work
        L_alg = 2 * n                     # SURVEY 8 named shapes: L = 2n is smooth
        b_step = (R * (n * b + 3 * L_alg * b) + P * (4 * L_alg * b + K * b)) / P
        ws_mb = (R * n * b + R * pinfo.spectrum_elems * 2 * b + P * K * b) / 1e6
        h2d, d2h = int(R * n * b), int(P * K * b)
        pcie_floor_ms = max(h2d / (pcie["h2d_gbs"] * 1e9), d2h / (pcie["d2h_gbs"] * 1e9)) * 1e3
        out = {
            "metric": "pair-days/s (24h@250Hz CCF)" if n == 21_600_000 else "pair-windows/s",
            "value": value, "unit": "pair-days/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": args.workload, "description": wl["desc"], "receivers": R, "pairs": P,
                       "n": n, "fft_len": L, "passes": list(pinfo.dims), "blocks": pinfo.n_blocks,
                       "block_len": pinfo.block_len, "lags": K, "plan": plan_info,
                       "l2": f"inputs larger than L2: working set per step {ws_mb:.0f} MB (inputs + spectra + CCFs) "
                             "vs 126 MB; no flush needed",
                       "timing": "CUDA events on the launching stream; per-kernel events are recorded inside the "
                                 "timed region (they cost a few us per launch and are included in `value`)",
                       "sharding": "windows over ranks, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": "pair-days/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pcie": pcie, "pcie_floor_ms_per_step": pcie_floor_ms,
                    "frac_of_pcie": pcie_floor_ms / (e2e_ms / args.steps),
                    "how": "xcorr_b200.core.process (drop-in of core/process.py:82-165) on a CF-1.9 variable set of "
                           f"120 pairs x {wpc} windows per call: page-locked host windows from the waveform callback "
                           "-> H2D -> forward + correlate -> D2H straight into the page-locked dataset.cc, 3 streams, "
                           "double-buffered; pcie = page-locked copy bandwidth measured in this run with all ranks "
                           "copying at once; frac_of_pcie = time of the slower direction alone / e2e step time"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak,
                         "unit": "GB/s", "frac": dom["achieved_gbs"] / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "avg_launch_ms": dom["avg_ms"],
                         "alg_bytes_per_launch": dom["alg_bytes_per_launch"],
                         "share_of_step": dom["total_ms"] / sum(t["total_ms"] for t in table)},
            "step_roofline": {"model": "SURVEY 8d two-pass minimal traffic, full-window transforms (L = %d)" % L_alg,
                              "alg_bytes_per_pair_day": b_step,
                              "achieved": b_step * value / world / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": b_step * value / world / 1e9 / peak},
            "kernels": table,
            "configs": legs,
        }
        if world == 1 and not args.no_cpu:
            cb = cpu_reference(n, steps=1)
            out["cpu_baseline"] = {"value": cb["value"], "unit": "pair-days/s", "cores": cb["cores"],
                                   "kind": cb["kind"], "sample": cb["sample"]}
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cb = cpu_reference(wl["n"], steps=args.steps)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {
        "impl": "reference",
        "metric": "pair-days/s (24h@250Hz CCF)" if wl["n"] == 21_600_000 else "pair-windows/s",
        "value": cb["value"], "unit": "pair-days/s", "n_gpus": world, "steps": cb["steps_run"],
        "steps_requested": args.steps, "warmup": 0,
        "ms_per_step": cb["wall"] / cb["steps_run"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"], "n": wl["n"],
                   "note": "full-size 24h@250Hz pair-windows (n = 21.6 M, all 2n-1 lags computed as the reference does, "
                           f"then cropped); steps are capped by a {REF_BUDGET_S:.0f} s wall budget (one step = one "
                           "pair-window per core, ~1-2 min); no warm-up step (the worker pool is started with a "
                           "tiny job)"},
        "cpu_baseline": {"value": cb["value"], "unit": "pair-days/s", "cores": cb["cores"], "kind": cb["kind"],
                         "sample": cb["sample"]},
        "e2e": {"value": cb["value"], "unit": "pair-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def _claim_stdout():
    """stdout must carry exactly one JSON line: anything a library prints on fd 1
    (e.g. NCCL's version banner) is diverted to stderr; emit() writes the line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


_REAL_STDOUT = None


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-legs", action="store_true", help="skip the legs of the other named configurations")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
