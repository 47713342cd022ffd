"""The batched process driver (core/process.py:82-165 replacement) against a
cell-by-cell oracle loop."""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

OPS = [("detrend", {"type": "demean"}),
       ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0}),
       ("filter", {"type": "bandpass", "freqmin": 3.0, "freqmax": 12.0}),
       ("decimate", {"factor": 5}),
       ("detrend", {"type": "demean"}),
       ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0})]


@pytest.mark.parametrize("dtype,unbiased", [(np.float32, False), (np.float64, True)])
def test_process_matches_cell_by_cell_oracle(dtype, unbiased):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xcorr_b200 import core
    from oracle import cc_oracle as orc
    from oracle import preprocess_oracle as pre

    recv = ["IM.H10N1..EDH", "IM.H10N2..EDH", "IM.H03S1..EDH"]
    pairs = [f"{recv[0]}-{recv[1]}", f"{recv[0]}-{recv[2]}", f"{recv[1]}-{recv[2]}"]
    fs_raw, fs, wl = 250.0, 50.0, 400.0
    ds = core.init(pairs, pd.Timestamp("2015-01-15"), pd.Timestamp("2015-01-15 00:30:00"), {"EDH": OPS},
                   dict(institution="i", author="a", source="s"), fs, window_length=wl, window_overlap=0.5,
                   clip_lag=(-20.0, 60.0), unbiased_cc=unbiased, dtype=dtype)
    times = pd.DatetimeIndex(ds.coords["time"].data)
    W, n_raw = len(times), int(wl * fs_raw)
    assert W == 9
    rng = np.random.default_rng(7)
    data = {}
    for w in range(W):
        src = orc.synthetic_receivers(3, n_raw, seed=300 + w, delay=65)
        t = np.arange(n_raw)
        for k, r in enumerate(recv):
            data[(r, times[w])] = src[k] + np.sin(2 * np.pi * 7.5 * t / fs_raw) * 0.1 + 2.0
    missing = {(recv[2], times[3])}
    short = {(recv[1], times[5])}
    ds.variables["status"].data[0, 1] = 1          # already processed: must be left alone
    ds.variables["cc"].data[0, 1] = 7.0

    def provider(receiver, time):
        if (receiver, time) in missing:
            return None
        d = data[(receiver, time)]
        if (receiver, time) in short:
            d = d[: n_raw // 2]                    # fails the strict duration check
        return d, fs_raw, 0.004 * recv.index(receiver)

    core.process(ds, provider, windows_per_batch=4, verb=0)
    status, cc = ds.variables["status"].data, ds.variables["cc"].data
    i0, i1 = ds.coords["lag"].attrs["index_min"], ds.coords["lag"].attrs["index_max"]
    npts = ds.coords["lag"].attrs["npts"]
    assert cc.shape == (3, W, i1 - i0) and cc.dtype == dtype
    tol = 1e-5 if dtype == np.float32 else 1e-11
    for q, p in enumerate(pairs):
        a, b = p.split("-")
        for w in range(W):
            if (q, w) == (0, 1):
                assert status[q, w] == 1 and np.all(cc[q, w] == 7.0)
                continue
            bad = any((r, times[w]) in missing or (r, times[w]) in short for r in (a, b))
            if bad:
                assert status[q, w] == -1 and not np.any(cc[q, w])
                continue
            assert status[q, w] == 1
            xa, _ = pre.process(data[(a, times[w])], fs_raw, OPS)
            xb_, _ = pre.process(data[(b, times[w])], fs_raw, OPS)
            ref = orc.process_cell(xa.astype(dtype), xb_.astype(dtype), npts, i0, i1, dtype=np.float64,
                                   bias_correct=unbiased)
            peak = np.max(np.abs(ref))
            assert np.max(np.abs(cc[q, w] - ref)) <= tol * peak, (q, w)
            assert np.argmax(cc[q, w]) == np.argmax(ref)
            assert abs(ds.variables["pair_offset"].data[q, w] - 0.004 * (recv.index(a) - recv.index(b))) < 1e-12
    assert ds.variables["cc"].attrs["unbiased"] == (1 if unbiased else 0)
    # retry_missing reprocesses -1 cells only
    missing.clear()
    core.process(ds, provider, retry_missing=True, verb=0)
    assert status[1, 3] == 1 and status[0, 5] == -1 and np.all(cc[0, 1] == 7.0)


def test_buffered_windows_trim_injection_and_config_errors():
    """Client semantics (stream/client.py:499-519, 996-1050): traces arrive with a 5 % buffer on both sides,
    ``('trim', {})`` gets the window injected (stream/process.py:153-196), the processed trace is trimmed to the
    window and must pass the strict duration check.  A configuration error raises instead of marking cells -1."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xcorr_b200 import core
    from oracle import cc_oracle as orc
    from oracle import preprocess_oracle as pre

    ops = [("detrend", {"type": "demean"}),
           ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0}),
           ("filter", {"type": "highpass", "freq": 3.0}),
           ("decimate", {"factor": 5}),
           ("trim", {}),
           ("detrend", {"type": "demean"})]
    recv = ["IM.H10N1..EDH", "IM.H10N2..EDH"]
    pairs = [f"{recv[0]}-{recv[1]}"]
    fs_raw, fs, wl, buf = 250.0, 50.0, 200.0, 0.05
    ds = core.init(pairs, pd.Timestamp("2015-01-15"), pd.Timestamp("2015-01-15 00:10:00"), {"EDH": ops},
                   dict(institution="i", author="a", source="s"), fs, window_length=wl, window_overlap=0.0,
                   clip_lag=(-10.0, 30.0), dtype=np.float64)
    times = pd.DatetimeIndex(ds.coords["time"].data)
    n_buf = int(round(buf * wl * fs_raw))
    n_raw = int(wl * fs_raw) + 2 * n_buf + 1            # [t0 - 5 %, t1 + 5 %], both ends included
    data = {}
    for w, t in enumerate(times):
        src = orc.synthetic_receivers(2, n_raw, seed=500 + w, delay=40)
        for k, r in enumerate(recv):
            data[(r, t)] = src[k] + 1.5
    # receiver 1 starts 3 raw samples late: its window is cut 3/250 s later than receiver 0's -> rounding to the grid
    late = {recv[0]: 0, recv[1]: 3}
    too_short = {(recv[1], times[2])}

    def provider(receiver, time):
        d = data[(receiver, time)]
        s = late[receiver]
        if (receiver, time) in too_short:
            return d[: n_buf + int(0.9 * wl * fs_raw)], fs_raw, -buf * wl      # ends inside the window
        return d[s:], fs_raw, -buf * wl + s / fs_raw

    core.process(ds, provider, windows_per_batch=2, verb=0)
    status, cc = ds.variables["status"].data, ds.variables["cc"].data
    i0, i1 = ds.coords["lag"].attrs["index_min"], ds.coords["lag"].attrs["index_max"]
    npts = ds.coords["lag"].attrs["npts"]
    for w, t in enumerate(times):
        if (recv[1], t) in too_short:
            assert status[0, w] == -1 and not np.any(cc[0, w])
            continue
        assert status[0, w] == 1
        xs, starts = [], []
        for r in recv:
            s = late[r]
            y, rate, t_first = pre.process(data[(r, t)][s:], fs_raw, ops, trace_start=-buf * wl + s / fs_raw,
                                           starttime=0.0, endtime=wl, return_start=True)
            y, t_first = pre.trim_time(y, rate, t_first, 0.0, wl)             # client.py:1017
            assert rate == fs and len(y) >= npts
            xs.append(y[:npts])
            starts.append(t_first)
        ref = orc.process_cell(xs[0], xs[1], npts, i0, i1, dtype=np.float64)
        assert np.max(np.abs(cc[0, w] - ref)) <= 1e-11 * np.max(np.abs(ref)), w
        assert abs(ds.variables["pair_offset"].data[0, w] - (starts[0] - starts[1])) < 1e-12
        assert abs(ds.variables["time_offset"].data[0, w] - starts[0]) < 1e-12
    from xcorr_b200.io import hash_dataset
    assert ds.attrs["sha256_hash_metadata"] == hash_dataset(ds)              # core/process.py:183
    # configuration errors are raised, not recorded as missing data
    for bad in ([("remove_response", {"output": "VEL"})], [("decimate", {"factor": 32})], [("detrend", {"typ": 1})]):
        ds2 = core.init(pairs, pd.Timestamp("2015-01-15"), pd.Timestamp("2015-01-15 00:10:00"), {"EDH": bad},
                        dict(institution="i", author="a", source="s"), fs, window_length=wl, window_overlap=0.0,
                        dtype=np.float64)
        with pytest.raises((NotImplementedError, ArithmeticError, TypeError)):
            core.process(ds2, provider, verb=0)
        assert not np.any(ds2.variables["status"].data)


def test_overlapping_windows_share_one_transfer_and_day_batched_files(tmp_path):
    """core/init.py:28 default overlap (0.875): eight windows a day share 87.5 % of their samples.  Windows handed
    over as views of ONE longer trace are copied to the device once (union), results equal the window-by-window
    path.  Then the day-batched task loop (core/lazy.py:26-82 for all pairs of a day): per-pair files, resume."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xcorr_b200 import core, io
    import importlib
    proc_mod = importlib.import_module("xcorr_b200.core.process")
    from oracle import cc_oracle as orc

    recv = ["IM.H10N1..EDH", "IM.H10N2..EDH", "IM.H03S1..EDH"]
    pairs = [f"{recv[0]}-{recv[1]}", f"{recv[0]}-{recv[2]}", f"{recv[1]}-{recv[2]}"]
    fs, wl, ov = 50.0, 800.0, 0.875
    ops = {"EDH": [("detrend", {"type": "demean"}), ("taper", {"type": "cosine", "max_percentage": 0.05})]}
    t0 = pd.Timestamp("2015-01-15")
    init_args = dict(preprocess=ops, attrs=dict(institution="i", author="a", source="s"), sampling_rate=fs,
                     window_length=wl, window_overlap=ov, clip_lag=(-5.0, 20.0), dtype=np.float32)
    ds = core.init(pairs, t0, t0 + pd.Timedelta(seconds=800), **init_args)
    times = pd.DatetimeIndex(ds.coords["time"].data)
    W, n, hop = len(times), int(wl * fs), int(wl * (1 - ov) * fs)
    assert W == 8 and hop == 5000
    long = {r: (orc.synthetic_receivers(3, n + W * hop, seed=77, delay=31)[k] + 0.5).astype(np.float32)
            for k, r in enumerate(recv)}

    def views(receiver, time):
        w = times.get_loc(time)
        return long[receiver][w * hop:w * hop + n], fs, 0.0          # a view of the continuous trace

    def copies(receiver, time):
        d, rate, off = views(receiver, time)
        return d.copy(), rate, off

    core.process(ds, views, windows_per_batch=8, verb=0)
    h2d_shared = proc_mod.last_stats["h2d_bytes"]
    ref = core.init(pairs, t0, t0 + pd.Timedelta(seconds=800), **init_args)
    core.process(ref, copies, windows_per_batch=8, verb=0)
    assert proc_mod.last_stats["h2d_bytes"] == 3 * W * n * 4
    assert h2d_shared == 3 * (n + (W - 1) * hop) * 4 < 0.25 * proc_mod.last_stats["h2d_bytes"]
    assert np.array_equal(ds.variables["cc"].data, ref.variables["cc"].data) and ds.variables["status"].data.min() == 1
    i0, i1 = ds.coords["lag"].attrs["index_min"], ds.coords["lag"].attrs["index_max"]
    from oracle import preprocess_oracle as pre
    x0, _ = pre.process(long[recv[0]][2 * hop:2 * hop + n], fs, ops["EDH"])
    x1, _ = pre.process(long[recv[1]][2 * hop:2 * hop + n], fs, ops["EDH"])
    cell = orc.process_cell(x0.astype(np.float32), x1.astype(np.float32), n, i0, i1, dtype=np.float64)
    assert np.max(np.abs(ds.variables["cc"].data[0, 2] - cell)) <= 1e-5 * np.max(np.abs(cell))

    # ---- the day-batched task loop ----
    root = str(tmp_path)
    day_args = dict(init_args, window_overlap=0.0)
    gone = {"n": 0}

    def provider(receiver, time):
        if receiver == recv[2] and gone["n"] == 0:
            return None                                               # first run: one receiver is missing
        w = (time - t0) // pd.Timedelta(seconds=wl)
        return long[receiver][int(w) * hop:int(w) * hop + n], fs, 0.0

    res = core.process_day(pairs, t0, day_args, provider, root, freq="1600s", verb=0)
    assert res == {pairs[0]: True, pairs[1]: False, pairs[2]: False}
    f0 = io.ncfile(pairs[0], t0, root)
    assert os.path.exists(f0) and not os.path.exists(io.ncfile(pairs[1], t0, root))
    first = io.read(f0)
    gone["n"] = 1
    res = core.process_day(pairs, t0, day_args, provider, root, freq="1600s", verb=0)
    assert res == {p: True for p in pairs}                           # pair 0 was complete: skipped, the others written
    again = io.read(f0)
    assert np.array_equal(first.variables["cc"].data, again.variables["cc"].data)
    one = io.read(io.ncfile(pairs[2], t0, root))
    assert one.variables["cc"].data.shape[0] == 1 and one.variables["status"].data.min() == 1
    assert list(one.coords["pair"].data) == [pairs[2]]
    assert core.day_tasks(pairs, pd.date_range(t0, periods=2, freq="D"),
                          done=lambda p, t: p == pairs[0]) == {t0: pairs[1:], t0 + pd.Timedelta("1D"): pairs[1:]}
