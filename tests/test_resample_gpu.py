"""GPU parity of the sampling-rate operations of the op-lists (obspy Trace.resample,
Trace.interpolate(method='lanczos')) against the numpy restatement of the obspy
algorithms (oracle.preprocess_oracle, parity unpinned -- see its header)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xcorr_b200.stream import process as sp
    from oracle import preprocess_oracle as pre
    return torch, sp, pre


def traces(n_sig, n, fs, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    return rng.standard_normal((n_sig, n)) + np.sin(2 * np.pi * 1.5 * t)


@pytest.mark.parametrize("npts,fs,new_fs", [(4000, 40.0, 50.0), (4001, 40.0, 50.0), (9000, 250.0, 50.0),
                                            (7777, 100.0, 30.0), (144_000, 40.0, 50.0)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_resample_matches_oracle(env, npts, fs, new_fs, dtype):
    torch, sp, pre = env
    x = traces(2, npts, fs, seed=npts)
    got, fs_out = sp.process(torch.from_numpy(x.astype(dtype)).cuda(), [("resample", {"sampling_rate": new_fs})],
                             sampling_rate=fs, raise_error=True)
    assert fs_out == new_fs
    got = got.cpu().numpy()
    for s in range(2):
        ref, _ = pre.resample(x[s].astype(dtype).astype(np.float64), fs, new_fs)
        assert got[s].shape == ref.shape
        tol = 1e-12 if dtype == np.float64 else 1e-5
        assert np.max(np.abs(got[s] - ref)) <= tol * np.max(np.abs(ref))


def test_resample_with_anti_alias_filter_and_chain(env):
    torch, sp, pre = env
    fs, npts = 250.0, 30_000
    x = traces(2, npts, fs, seed=2) + 3.0
    ops = [("detrend", {"type": "demean"}), ("resample", {"sampling_rate": 50.0, "no_filter": False}),
           ("taper", {"type": "cosine", "max_percentage": 0.05})]
    got, fs_out = sp.process(torch.from_numpy(x).cuda(), ops, sampling_rate=fs, raise_error=True)
    got = got.cpu().numpy()
    for s in range(2):
        ref, fo = pre.process(x[s], fs, ops)
        assert fo == fs_out == 50.0
        assert np.max(np.abs(got[s] - ref)) <= 1e-11 * np.max(np.abs(ref))


@pytest.mark.parametrize("npts,fs,new_fs,a", [(4000, 40.0, 50.0, 20), (5000, 100.0, 37.5, 8), (1000, 20.0, 80.0, 12)])
def test_lanczos_interpolate_matches_oracle(env, npts, fs, new_fs, a):
    torch, sp, pre = env
    x = traces(3, npts, fs, seed=a)
    ops = [("interpolate", {"sampling_rate": new_fs, "method": "lanczos", "a": a})]
    got, fs_out = sp.process(torch.from_numpy(x).cuda(), ops, sampling_rate=fs, raise_error=True)
    got = got.cpu().numpy()
    for s in range(3):
        ref, _ = pre.interpolate_lanczos(x[s], fs, new_fs, a)
        assert got[s].shape == ref.shape
        assert np.max(np.abs(got[s] - ref)) <= 1e-12 * np.max(np.abs(ref))
    # a band-limited signal is reproduced at the new times
    t_new = np.arange(got.shape[1]) / new_fs
    clean = np.sin(2 * np.pi * 1.5 * np.arange(npts) / fs)[None, :]
    got_c, _ = sp.process(torch.from_numpy(clean).cuda(), ops, sampling_rate=fs, raise_error=True)
    inner = slice(4 * a, got_c.shape[1] - 4 * a)
    assert np.max(np.abs(got_c.cpu().numpy()[0, inner] - np.sin(2 * np.pi * 1.5 * t_new)[inner])) < 5e-3


def test_unsupported_variants_raise(env):
    torch, sp, pre = env
    x = torch.zeros((1, 100), dtype=torch.float64, device="cuda")
    with pytest.raises(RuntimeError):
        sp.process(x, [("interpolate", {"sampling_rate": 10.0, "method": "cubic"})], sampling_rate=20.0, raise_error=True)
    with pytest.raises(RuntimeError):
        sp.process(x, [("resample", {"sampling_rate": 10.0, "window": "blackman"})], sampling_rate=20.0, raise_error=True)


@pytest.mark.parametrize("npts,fs,window", [(30_001, 50.0, 1.0), (30_000, 50.0, 10.0), (7_777, 40.0, 2.5)])
def test_running_rms_matches_oracle(env, npts, fs, window):
    torch, sp, pre = env
    x = traces(3, npts, fs, seed=int(window * 10))
    for dtype in (np.float32, np.float64):
        got, fs_out = sp.process(torch.from_numpy(x.astype(dtype)).cuda(), [("running_rms", {"window": window})],
                                 sampling_rate=fs, raise_error=True, out_dtype=torch.float64)
        assert fs_out == 1.0 / window
        for s in range(3):
            ref, _ = pre.running_rms(x[s].astype(dtype).astype(np.float64), fs, window)
            assert got[s].shape == ref.shape and np.allclose(got[s].cpu().numpy(), ref, rtol=1e-13)


def test_normalize_matches_oracle(env):
    torch, sp, pre = env
    x = traces(3, 5000, 50.0, seed=4) * np.array([[1.0], [-7.5], [1e-3]])
    ops = [("detrend", {"type": "demean"}), ("normalize", {}), ("taper", {"type": "hann", "max_percentage": 0.05})]
    got, _ = sp.process(torch.from_numpy(x).cuda(), ops, sampling_rate=50.0, raise_error=True)
    for s in range(3):
        ref, _ = pre.process(x[s], 50.0, ops)
        assert np.max(np.abs(got[s].cpu().numpy() - ref)) <= 1e-13
    got, _ = sp.process(torch.from_numpy(x).cuda(), [("normalize", {"norm": -4.0})], sampling_rate=50.0, raise_error=True)
    assert np.allclose(got.cpu().numpy(), x / 4.0, rtol=1e-15)
