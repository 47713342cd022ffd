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

Printed JSON line: `value` = device-resident throughput (inputs in HBM), `e2e`
= through the public API with pinned HOST buffers (H2D of the raw windows and
D2H of the CCFs inside the timed region, copies overlapped with compute on
separate streams), `roofline` = dominant kernel vs the measured HBM peak,
`cpu_baseline` = the reference algorithm (oracle port of util/cc.py) on the
host cores over a bounded sample.
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
# algorithmic bytes per pair-window = [N(n b + 3 L b) + P(4 L b + K b)] / P
CPU_SAMPLE_FRACTION = 16          # CPU arms time 1/16-day windows (n = 1.35 M), see cpu_reference()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------
# CPU reference arm: the reference's algorithm (oracle port of util/cc.py) on
# host cores, one single-threaded task per pair-window like core/lazy.py does.
# ---------------------------------------------------------------------------
def _cpu_task(args):
    n, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import cc_oracle as orc
    rec = orc.synthetic_receivers(2, n, seed=seed, delay=137).astype(np.float32)
    t0 = time.perf_counter()
    r = orc.cc(rec[0], rec[1], normalize=True, pad=True, unbiased=False, dtype=np.float32)
    dt = time.perf_counter() - t0
    return dt, int(np.argmax(r)) - (n - 1)


def cpu_reference(n_day, steps, warmup, cores=None, fraction=CPU_SAMPLE_FRACTION):
    """Each step: `cores` independent pair-windows of n_day/fraction samples,
    each counted as 1/fraction pair-day.  (A full 24h@250Hz window costs the
    reference ~75 s on one core -- its length-(2n-1) FFT falls to Bluestein --
    so the timed sample is bounded; FFT cost grows slightly faster than n, so
    the scaled figure flatters the CPU.)"""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    n = n_day // fraction
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(warmup):
            pool.map(_cpu_task, [(n, 10_000 + w * cores + c) for c in range(cores)])
        t0 = time.perf_counter()
        lags = []
        for s in range(steps):
            res = pool.map(_cpu_task, [(n, 20_000 + s * cores + c) for c in range(cores)])
            lags += [r[1] for r in res]
        wall = time.perf_counter() - t0
    assert all(l == 137 for l in lags), "CPU reference lost the synthetic arrival"
    units = steps * cores / fraction
    return dict(value=units / wall, wall=wall, cores=cores, n=n,
                sample=f"{steps} step(s) x {cores} single-threaded pair-windows of 1/{fraction} day "
                       f"(n={n}) on {cores} cores, each counted as 1/{fraction} pair-day; "
                       "oracle port of xcorr/util/cc.py (numpy pocketfft, float32 in)")


def measured_traffic(args, plan, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed
    `ncu --set full` capture (profiles/ncu_r01_summary.md); only for the configuration that capture ran."""
    captured = {("cfg3", "f32", (256, 150, 256), 2): {
        "inv_pass2_N256": 3.618e9 + 9.384e9,      # xcb_pair_product_grouped: every spectrum once + one spectrum per pair
        "inv_pass1_N150": 9.437e9 + 9.378e9,
        "inv_pass0_N256": 9.437e9 + 3.864e9,
        "fwd_pass1_N150": 3.775e9 + 3.715e9,
        "fwd_pass2_N256": 3.776e9 + 3.720e9,
    }}
    key = (args.workload, args.dtype, tuple(plan.dims), plan.n_blocks)
    return captured.get(key, {}).get(kernel)


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.handle = None
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
            self.stop_flag.wait(0.005)

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
    # parity spot-check of what is being timed: arrival of pair (0, 1) at +137 samples
    idx, _ = plan.argmax(ccf)
    assert int(idx[0, 0]) == (n - 1 - i0) + 137, "benchmark output lost the synthetic arrival"

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

    # ---------------- end to end with host buffers ----------------
    x_host = x.cpu().pin_memory()
    out_host = [torch.empty((P, 1, K), dtype=tdt).pin_memory() for _ in range(2)]
    x_dev = [torch.empty_like(x) for _ in range(2)]
    ccf_dev = [ccf, torch.empty_like(ccf)]
    s_in, s_run, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_run = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_steps(k):
        for i in range(k):
            s = i % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_run[s])             # x_dev[s] free again
                x_dev[s].copy_(x_host, non_blocking=True)
                ev_in[s].record(s_in)
            with torch.cuda.stream(s_run):
                s_run.wait_event(ev_in[s])
                s_run.wait_event(ev_out[s])            # ccf_dev[s] drained
                plan.forward(x_dev[s])
                plan.correlate(out=ccf_dev[s])
                ev_run[s].record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_run[s])
                out_host[s].copy_(ccf_dev[s], non_blocking=True)
                ev_out[s].record(s_out)

    e2e_steps(max(2, args.warmup))
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for s_ in (s_in, s_run, s_out):
        s_.wait_event(t0)
    e2e_steps(args.steps)
    cur = torch.cuda.current_stream(dev)
    for s_ in (s_in, s_run, s_out):
        ev = torch.cuda.Event()
        ev.record(s_)
        cur.wait_event(ev)
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    assert int(np.argmax(out_host[(args.steps - 1) % 2][0, 0].numpy())) == (n - 1 - i0) + 137

    # ---------------- reduce over ranks ----------------
    tmax = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(tmax[0]), float(tmax[1])
    units = world * args.steps * P
    value = units / (dev_ms * 1e-3)
    e2e_value = units / (e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        # per-launch algorithmic bytes of each pass kind: inputs read once + outputs written
        # once, for the transforms the plan actually runs (F = plan.fft_len; block mode: 2B spectra per signal)
        F, B = L, plan.n_blocks
        per = plan.spectrum_elems // (F // 2)          # spectra per signal (2B, or 2B-1 when the last segment aliases)
        if B > 1:
            lb = plan.block_len
            nseg = per - B
            valid = sum(min(F, n - k * lb) for k in range(nseg)) + sum(min(lb, n - k * lb) for k in range(B))
        else:
            valid = n
        per_step = {}
        for a in range(nd):
            rd = R * valid * b if a == 0 else R * per * F * b
            per_step[f"fwd_pass{a}_N{plan.dims[a]}"] = rd + R * per * F * b
            if a == nd - 1:
                # pair product: every distinct spectrum of the window is read from HBM once per launch
                # (the pairs that share it are served by L2 / the staged shared-memory rows), one packed
                # spectrum written per pair
                bytes_inv = R * per * F * b + plan.n_jobs * (F * b if nd > 1 else K * b)
            elif a == 0:
                bytes_inv = (F * b + K * b) * plan.n_jobs
            else:
                bytes_inv = 2 * F * b * plan.n_jobs
            per_step[f"inv_pass{a}_N{plan.dims[a]}"] = bytes_inv
        table = []
        for name, ms, cnt in kinds:
            if cnt == 0:
                continue
            total_bytes = per_step[name] * args.steps
            table.append(dict(kernel=name, launches=cnt, total_ms=ms, avg_ms=ms / cnt,
                              alg_bytes_per_launch=total_bytes / cnt,
                              achieved_gbs=total_bytes / (ms * 1e-3) / 1e9))
        dom = max(table, key=lambda t: t["total_ms"])
        # step model of SURVEY 8d: a property of the PROBLEM (full-window transform length L_alg, two HBM
        # passes per transform), independent of how the plan factorises or blocks the work
        L_alg = 2 * n                     # SURVEY 8 named shapes: L = 2n is smooth
        b_alg = (R * (n * b + 3 * L_alg * b) + P * (4 * L_alg * b + K * b)) / P
        out = {
            "metric": "pair-days/s (24h@250Hz CCF)" if n == 21_600_000 else "pair-windows/s",
            "value": value, "unit": "pair-days/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": args.workload, "description": wl["desc"], "receivers": R, "pairs": P,
                       "n": n, "fft_len": L, "passes": list(plan.dims), "blocks": plan.n_blocks,
                       "block_len": plan.block_len, "lags": K,
                       "l2": "working set per step (inputs 1.4 GB, spectra 2.8 GB, CCFs 3.9 GB) exceeds the 126 MB L2; no flush needed",
                       "sharding": "windows over ranks, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": "pair-days/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(R * n * b), "d2h_bytes_per_step": int(P * K * b),
                    "how": "pinned host windows -> H2D -> forward+correlate -> D2H of cropped CCFs, 3 streams, double-buffered"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak,
                         "unit": "GB/s", "frac": dom["achieved_gbs"] / peak,
                         "traffic": measured_traffic(args, plan, dom["kernel"]),
                         "peak_source": peak_src, "avg_launch_ms": dom["avg_ms"],
                         "alg_bytes_per_launch": dom["alg_bytes_per_launch"],
                         "share_of_step": dom["total_ms"] / sum(t["total_ms"] for t in table)},
            "step_roofline": {"model": "SURVEY 8d two-pass minimal traffic, full-window transforms (L = %d)" % L_alg,
                              "alg_bytes_per_pair_day": b_alg,
                              "achieved": b_alg * value / world / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": b_alg * value / world / 1e9 / peak},
            "kernels": table,
        }
        if world == 1 and not args.no_cpu:
            cb = cpu_reference(n, steps=1, warmup=0)
            out["cpu_baseline"] = {"value": cb["value"], "unit": "pair-days/s", "cores": cb["cores"],
                                   "kind": "port", "sample": cb["sample"]}
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cb = cpu_reference(wl["n"], steps=args.steps, warmup=min(args.warmup, 1))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {
        "impl": "reference",
        "metric": "pair-days/s (24h@250Hz CCF)" if wl["n"] == 21_600_000 else "pair-windows/s",
        "value": cb["value"], "unit": "pair-days/s", "n_gpus": world, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": cb["wall"] / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"]},
        "cpu_baseline": {"value": cb["value"], "unit": "pair-days/s", "cores": cb["cores"], "kind": "port",
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
