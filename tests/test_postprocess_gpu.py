"""GPU parity of the CCF post-processing operators (core/postprocess.py:133-141 chain:
unbias, demean, taper, timeshift, zero-phase Butterworth, taper; and signal/snr.py)
against oracle.signal_oracle, which calls the very numpy / scipy routines the
reference calls."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {np.float32: 1e-5, np.float64: 1e-12}


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import xcorr_b200.signal as sg
    from oracle import signal_oracle as so
    return torch, sg, so


def ccfs(shape, seed):
    rng = np.random.default_rng(seed)
    n = shape[-1]
    t = np.arange(n)
    burst = np.exp(-0.5 * ((t - 0.6 * n) / (0.02 * n)) ** 2) * np.sin(2 * np.pi * t / 9.0)
    return 0.1 * rng.standard_normal(shape) + burst + 0.05


def _max_rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_demean_taper_unbias(env, dtype):
    torch, sg, so = env
    x = ccfs((2, 3, 5000), 1).astype(dtype)
    xd = torch.from_numpy(x).cuda()
    assert _max_rel(sg.demean(xd).cpu().numpy(), so.demean(x.astype(np.float64))) <= TOL[dtype]
    for kw in ({"max_length": 5.0}, {"max_percentage": 0.1}, {"wtype": "blackman", "max_length": 1.5}, {}):
        got = sg.taper(xd, 50.0, **kw).cpu().numpy()
        ref = so.taper(x, 50.0, **{"wtype": "hann", **kw})
        assert got.dtype == x.dtype and _max_rel(got, ref) <= (0 if dtype == np.float64 else 1e-7)
    w = sg.unbias_weights(40_000, 39_999, 39_999 + 5000)
    assert np.array_equal(sg.unbias(xd, w).cpu().numpy(), x * w.astype(dtype))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [4096, 5000, 5001, 16_383])
def test_timeshift(env, n, dtype):
    torch, sg, so = env
    x = ccfs((3, 2, n), n).astype(dtype)
    delay = np.array([[0.013, -0.4], [1.234567, 0.0], [-2.5, 0.02]])
    got = sg.timeshift(torch.from_numpy(x).cuda(), delay, 50.0).cpu().numpy()
    ref = so.timeshift(x.astype(np.float64), delay, 0.0, 1 / 50.0)
    assert got.shape == ref.shape and got.shape[-1] == n - (n & 1)
    assert _max_rel(got, ref) <= TOL[dtype]
    # an integer number of samples is an exact circular roll
    k = 7
    got = sg.timeshift(torch.from_numpy(x[..., : n - (n & 1)].astype(np.float64)).cuda(), k / 50.0, 50.0).cpu().numpy()
    assert _max_rel(got, np.roll(x[..., : n - (n & 1)].astype(np.float64), k, axis=-1)) <= 1e-12


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kw", [dict(frequency=3.0, btype="highpass", order=2),
                                dict(frequency=(1.0, 8.0), btype="bandpass", order=4),
                                dict(frequency=5.0, btype="lowpass", order=3)])
def test_zero_phase_butterworth(env, kw, dtype):
    torch, sg, so = env
    x = ccfs((4, 20_001), 7).astype(dtype)
    got = sg.filter(torch.from_numpy(x).cuda(), 50.0, **kw).cpu().numpy()
    ref = so.filter(x.astype(np.float64), 50.0, **kw)
    assert got.dtype == x.dtype
    assert _max_rel(got, ref) <= (1e-11 if dtype == np.float64 else 1e-5)


def test_snr(env):
    torch, sg, so = env
    x = ccfs((5, 30_000), 3)
    smask = np.zeros(30_000, bool); smask[15_000:21_000] = True
    nmask = np.zeros(30_000, bool); nmask[2_000:9_000] = True
    ratio, s, n, idx = sg.snr(torch.from_numpy(x).cuda(), (15_000, 21_000), (2_000, 9_000), extend=True)
    r_ref, s_ref, n_ref, i_ref = so.snr(x, smask, nmask)
    assert np.array_equal(idx, i_ref) and np.array_equal(s, s_ref)
    assert np.allclose(n, n_ref, rtol=1e-13) and np.allclose(ratio, r_ref, rtol=1e-13)
    db = sg.snr(torch.from_numpy(x.astype(np.float32)).cuda(), (15_000, 21_000), (2_000, 9_000), decibels=True)
    assert np.allclose(db, 20 * np.log10(so.snr(x.astype(np.float32).astype(np.float64), smask, nmask)[0]), rtol=1e-6)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_postprocess_chain_and_dataset_driver(env, dtype):
    """core.postprocess on a small variable set == the oracle chain cell by cell."""
    torch, sg, so = env
    import pandas as pd
    import xcorr_b200.core as core
    from xcorr_b200.util.cc import weight
    fs, npts = 50.0, 40_000
    pairs = ["NET.AAA..HHZ-NET.BBB..HHZ", "NET.AAA..HHZ-NET.CCC..HHZ"]
    ds = core.init(pairs[0], pd.Timestamp("2020-01-01"), pd.Timestamp("2020-01-01 18:00"),
                   preprocess={"HHZ": [("detrend", {"type": "demean"})]}, sampling_rate=fs,
                   window_length=npts / fs, window_overlap=0.0, clip_lag=(0.0, 200.0), dtype=dtype,
                   attrs=dict(title="t", institution="i", author="a", source="s"))
    nt, nl = len(ds.coords["time"].data), len(ds.coords["lag"].data)
    rng = np.random.default_rng(5)
    cc = ccfs((1, nt, nl), 9).astype(dtype)
    ds.variables["cc"].data[...] = cc
    ds.variables["status"].data[...] = 1
    ds.variables["status"].data[0, 1] = -1
    ds.variables["pair_offset"].data[...] = 0.0123
    ds.variables["time_offset"].data[...] = rng.uniform(-0.01, 0.01, size=(1, nt))
    out = core.postprocess(ds, lag_lim=(10.0, 150.0))
    lag = ds.coords["lag"].data
    lm = (lag >= 10.0) & (lag <= 150.0)
    keep = np.array([True] + [False] + [True] * (nt - 2))
    assert np.array_equal(out.coords["time"].data, ds.coords["time"].data[keep])
    i0 = ds.coords["lag"].attrs["index_min"]
    l0 = int(np.flatnonzero(lm)[0]); l1 = int(np.flatnonzero(lm)[-1]) + 1
    w = weight(npts, pad=True)[i0 + l0:i0 + l1]
    delay = -(ds.variables["pair_offset"].data + ds.variables["time_offset"].data)[:, keep]
    ref = so.postprocess(cc[:, keep][:, :, l0:l1], w, delay, lag[l0], fs)
    got = out.variables["cc"].data
    assert got.shape == ref.shape and got.dtype == np.dtype(dtype)
    assert _max_rel(got, ref.astype(np.float64)) <= (1e-10 if dtype == np.float64 else 2e-5)
    assert out.variables["cc"].attrs["unbiased"] == 1 and "postprocess" in out.variables["cc"].attrs


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kw", [dict(duration=2.5, padding_factor=4), dict(duration=1.0, padding_factor=2, scaling="spectrum"),
                                dict(duration=0.64, padding_factor=1)])
def test_spectrogram(env, kw, dtype):
    """signal/spectrogram.py (timelapse input): scipy.signal.spectrogram with one segment per lag sample."""
    torch, sg, so = env
    x = ccfs((2, 3, 3001), 21).astype(dtype)
    got, f = sg.spectrogram(torch.from_numpy(x).cuda(), 50.0, **kw)
    ref, f_ref = so.spectrogram(x.astype(np.float64), 50.0, **kw)
    got = got.cpu().numpy()
    assert got.shape == ref.shape and np.array_equal(f, f_ref)
    assert np.max(np.abs(got - ref)) <= (1e-11 if dtype == np.float64 else 2e-5) * np.max(ref)
    edge = int(np.rint(int(kw["duration"] * 50.0) / 2))
    assert np.all(got[..., :edge] == 0) and np.all(got[..., -1] == 0)


def test_linear_detrend_and_rms(env):
    """signal/detrend.py (scipy.stats.linregress line) and signal/rms.py."""
    import scipy.stats
    torch, sg, so = env
    x = ccfs((2, 4, 3000), 33) + 1e-3 * np.arange(3000)
    lag = np.arange(3000) / 50.0 - 7.0
    got = sg.detrend(torch.from_numpy(x).cuda(), type="linear").cpu().numpy()
    ref = np.empty_like(x)
    for idx in np.ndindex(x.shape[:-1]):
        m, b, *_ = scipy.stats.linregress(lag, x[idx])
        ref[idx] = x[idx] - (m * lag + b)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert np.max(np.abs(sg.detrend(torch.from_numpy(x).cuda()).cpu().numpy() - so.demean(x))) <= 1e-12
    assert np.allclose(sg.rms(torch.from_numpy(x).cuda()), np.sqrt(np.mean(x * x, axis=-1)), rtol=1e-13)
