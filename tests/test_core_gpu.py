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
