"""The batched process driver (core/process.py:82-165 replacement) against a
cell-by-cell oracle loop."""
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
