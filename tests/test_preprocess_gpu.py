"""GPU parity of the preprocessing operations (C ABI: xcb_stats,
xcb_pointwise_apply, xcb_sosfilt) against the scipy restatement of the obspy
algorithms (oracle.preprocess_oracle, parity unpinned -- see its header)."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG2_OPS = [("detrend", {"type": "demean"}),
            ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0}),
            ("filter", {"type": "bandpass", "freqmin": 3.0, "freqmax": 12.0}),
            ("decimate", {"factor": 5}),
            ("detrend", {"type": "demean"}),
            ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0})]


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import xcorr_b200
    from xcorr_b200.stream import process as sp
    from oracle import preprocess_oracle as pre
    from oracle import cc_oracle as orc
    return torch, xcorr_b200, sp, pre, orc


def raw_traces(n_sig, n, fs, seed):
    """SURVEY 8d raw inputs: noise + in-band 7.5 Hz line + linear trend + DC offset."""
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = rng.standard_normal((n_sig, n)) + np.sin(2 * np.pi * 7.5 * t / fs) + 1e-6 * t + 3.0
    return x


def run(env, x, fs, ops, dtype=np.float64):
    torch, _, sp, pre, _ = env
    xt = torch.from_numpy(x.astype(dtype)).cuda()
    got, fs_out = sp.process(xt, ops, sampling_rate=fs, raise_error=True)
    got = got.cpu().numpy()
    ref = np.stack([pre.process(row.astype(dtype), fs, ops)[0] for row in x])
    fs_ref = pre.process(x[0], fs, ops)[1]
    assert fs_out == fs_ref and got.shape == ref.shape
    return got, ref


@pytest.mark.parametrize("ops", [
    [("detrend", {"type": "demean"})],
    [("detrend", {"type": "constant"})],
    [("detrend", {"type": "linear"})],
    [("detrend", {"type": "simple"})],
    [("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0})],
    [("taper", {"type": "hann", "max_percentage": 0.1})],
    [("taper", {"type": "cosine", "max_percentage": 0.05, "side": "left"})],
    [("taper", {"type": "hann", "max_percentage": 0.5})],
    [("detrend", {"type": "linear"}), ("taper", {"type": "cosine", "max_percentage": 0.05})],
    [("taper", {"type": "cosine", "max_percentage": 0.05}), ("detrend", {"type": "demean"})],
])
def test_pointwise_ops(env, ops):
    x = raw_traces(3, 100_001, 250.0, 1)
    got, ref = run(env, x, 250.0, ops)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("ops", [
    [("filter", {"type": "bandpass", "freqmin": 3.0, "freqmax": 12.0})],
    [("filter", {"type": "highpass", "freq": 0.05})],                       # pole radius 0.99952: long memory
    [("filter", {"type": "lowpass", "freq": 20.0, "corners": 2})],
    [("filter", {"type": "bandpass", "freqmin": 3.0, "freqmax": 12.0, "zerophase": True})],
    [("filter", {"type": "lowpass_cheby_2", "freq": 25.0})],
    [("decimate", {"factor": 5})],
    [("decimate", {"factor": 4, "no_filter": True})],
    [("filter", {"type": "highpass", "freq": 0.05}), ("detrend", {"type": "demean"}),
     ("decimate", {"factor": 5}), ("trim", {"start": 100, "length": 30_000})],
])
def test_filters_and_decimation(env, ops):
    x = raw_traces(2, 250 * 800 + 3, 250.0, 2)
    got, ref = run(env, x, 250.0, ops)
    err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
    print("relative error", ops[0], err)
    # a 0.05 Hz corner at 250 Hz (pole radius 0.99952) amplifies rounding by ~1/(1-|p|)^2 = 4e6:
    # two float64 evaluation orders of the same recurrence differ at the 1e-10 level there
    long_memory = any(p.get("freq") == 0.05 for _, p in ops)
    assert err <= (1e-9 if long_memory else 1e-12), err


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-6)])
def test_config2_chain_then_cc(env, dtype, tol):
    """BASELINE config 2 chain (250 Hz -> 50 Hz) followed by cc(): the
    preprocessed traces and the resulting CCF both match the CPU restatement."""
    torch, xb, sp, pre, orc = env
    fs, n_raw = 250.0, 250 * 3600
    src = orc.synthetic_receivers(3, n_raw, seed=2000, delay=137 * 5)
    t = np.arange(n_raw)
    x = src + np.sin(2 * np.pi * 7.5 * t / fs) + 1e-6 * t + 3.0
    got, ref = run(env, x, fs, CFG2_OPS, dtype)
    assert got.shape[1] == n_raw // 5
    print('trace error', dtype.__name__, np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
    assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref))
    n = got.shape[1]
    plan = xb.CorrelationPlan(n, 3, xb.all_pairs(3), dtype=torch.float32 if dtype == np.float32 else torch.float64)
    ccf = plan(torch.from_numpy(got).cuda()).cpu().numpy()[:, 0]
    for q, (i, j) in enumerate(xb.all_pairs(3)):
        r = orc.cc(ref[i], ref[j], unbiased=False)
        assert np.argmax(ccf[q]) == np.argmax(r)
        e = np.max(np.abs(ccf[q] - r)) / np.max(np.abs(r))
        print('ccf error', dtype.__name__, e)
        assert e <= (1e-12 if dtype == np.float64 else 1e-5)


def test_oplist_error_behaviour(env):
    torch, _, sp, _, _ = env
    x = torch.zeros((1, 1000), dtype=torch.float64, device="cuda")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert sp.process(x, [("remove_response", {})], sampling_rate=50.0) is None
        out = sp.process(x, [("not_an_op", {}), ("detrend", {"type": "demean"})], sampling_rate=50.0)
        assert out is not None and out[0].shape == (1, 1000)
        assert len(w) >= 2
    with pytest.raises(ValueError):
        sp.process(x, [("not_an_op", {})], sampling_rate=50.0, raise_error=True)
    with pytest.raises(RuntimeError):
        sp.process(x, [("remove_response", {})], sampling_rate=50.0, raise_error=True)
    with pytest.raises(RuntimeError):
        sp.process(x, [("decimate", {"factor": 20})], sampling_rate=50.0, raise_error=True)
    assert set(sp.device_operations()) == {"decimate", "detrend", "filter", "interpolate", "normalize", "resample", "resample_poly", "running_rms", "taper", "trim"}
    assert "remove_response" in sp.list_operations() and sp.is_operation("merge")


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-6)])
def test_polyphase_fir_rate_change_matches_scipy(env, dtype, tol):
    """('resample_poly', {up, down}): the north-star polyphase-FIR decimator against scipy.signal.resample_poly
    itself (zero-phase Kaiser-windowed FIR, upfirdn), alone and inside the config-2 style chain."""
    torch, _, sp, pre, _ = env
    from scipy import signal
    fs = 250.0
    x = raw_traces(3, 100_003, fs, seed=77)
    for up, down in ((1, 5), (1, 2), (2, 5), (3, 7), (5, 2), (1, 16), (4, 6), (1, 1)):
        ops = [("resample_poly", {"up": up, "down": down})]
        got, ref = run(env, x, fs, ops, dtype)
        assert np.array_equal(ref, np.stack([signal.resample_poly(r.astype(dtype).astype(np.float64), up, down) for r in x]))
        assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref)), (up, down)
    # other windows, short traces (shorter than the filter), and the decimating chain of config 2 with the FIR
    got, ref = run(env, x[:, :37], fs, [("resample_poly", {"up": 1, "down": 5, "window": "hamming"})], dtype)
    assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref))
    chain = [("detrend", {"type": "demean"}),
             ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0}),
             ("filter", {"type": "bandpass", "freqmin": 3.0, "freqmax": 12.0}),
             ("resample_poly", {"up": 1, "down": 5}),
             ("detrend", {"type": "demean"}),
             ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0})]
    got, ref = run(env, x, fs, chain, dtype)
    assert got.shape[1] == 20_001
    assert np.max(np.abs(got - ref)) <= (1e-11 if dtype == np.float64 else 1e-5) * np.max(np.abs(ref))
    with pytest.raises(RuntimeError):
        sp.process(torch.zeros((1, 100), device="cuda"), [("resample_poly", {"up": 1, "down": 5, "padtype": "line"})],
                   sampling_rate=fs, raise_error=True)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_windowed_carry_equals_hierarchical_carry(env, dtype):
    """xcb_sos.carry_window (start states of fast-decaying cascades in one parallel launch, every run of chunks
    warmed up over K chunks) against the unconditionally exact fold / sweep / expand form of the same program,
    forward (decimating, zero-input fast path) and zero-phase (reverse pass, full-rate stores); a slowly decaying
    high-pass must keep the hierarchical form."""
    torch, _, sp, _, _ = env
    fs, n = 250.0, 400_003                      # > 2 x 32 chunks, ragged last chunk, odd length
    x = raw_traces(3, n, fs, 77)
    x[1, 100_000:100_050] += 1e6                # a glitch eight orders above the noise
    xt = torch.from_numpy(x.astype(dtype)).cuda()
    for ops in (CFG2_OPS,
                [("filter", {"type": "bandpass", "freqmin": 3.0, "freqmax": 12.0, "zerophase": True})]):
        prog = sp.PreprocessProgram(ops, n, fs)
        prog._host_design()
        windows = [st[1].f.carry_window for st in prog.steps if st[0] == 'stage' and st[1].f is not None]
        assert windows and all(1 <= w <= 16 for w in windows)
        got = prog.run(xt, out_dtype=torch.float64).cpu().numpy()
        for st in prog.steps:
            if st[0] == 'stage' and st[1].f is not None:
                st[1].f.carry_window = 0
        ref = prog.run(xt, out_dtype=torch.float64).cpu().numpy()
        assert np.isfinite(ref).all()
        scale = np.abs(ref).max(axis=1, keepdims=True)
        assert (np.abs(got - ref) / scale).max() <= 1e-14
        half = got.shape[1] // 2                # long after the glitch: absolute agreement at the noise's own rounding
        assert np.abs(got[1, half:] - ref[1, half:]).max() <= 1e-13
    slow = sp.PreprocessProgram([("filter", {"type": "highpass", "freq": 0.05})], n, fs)
    slow._host_design()
    assert [st[1].f.carry_window for st in slow.steps if st[0] == 'stage' and st[1].f is not None] == [0]
