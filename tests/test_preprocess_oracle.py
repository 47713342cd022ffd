"""Known-answer tests pinning the (otherwise unpinned) obspy restatement."""
import numpy as np
import pytest
from scipy import signal

from oracle import preprocess_oracle as pre


def test_detrend_known_answers():
    i = np.arange(1000.0)
    x = 3.0 + 0.25 * i
    assert np.max(np.abs(pre.detrend(x, "linear"))) < 1e-9
    assert np.max(np.abs(pre.detrend(x, "simple"))) < 1e-9
    d = pre.detrend(x + np.sin(i), "demean")
    assert abs(d.mean()) < 1e-12
    assert np.allclose(pre.detrend(x, "constant"), x - x.mean())


def test_cosine_taper_and_window_shape():
    w = pre.cosine_taper(11, 1.0)
    assert w[0] == 0.0 and w[-1] == 0.0 and abs(w[5] - 1.0) < 1e-15 or w.max() <= 1.0
    # reference parameters: 5 %, at most 30 s, at 50 Hz over a day
    n, fs = 4_320_000, 50.0
    win = pre.taper_window(n, fs, max_percentage=0.05, type="cosine", max_length=30.0)
    wlen = 1500
    assert len(win) == n and win[0] == 0.0 and win[-1] == 0.0
    assert np.all(win[wlen:n - wlen] == 1.0) and np.all(np.diff(win[:wlen]) >= 0)
    assert np.allclose(win[:wlen], win[::-1][:wlen])
    # half-cosine ramp of 2*wlen+1 points
    assert np.allclose(win[:wlen], 0.5 * (1 - np.cos(np.pi * np.arange(wlen) / (wlen - 1) * (wlen - 1) / wlen)), atol=1e-3)
    h = pre.taper_window(1000, 1.0, max_percentage=0.1, type="hann")
    assert np.allclose(h[:100], signal.windows.hann(201)[:100])


def test_filter_designs_match_survey_numbers():
    sos = pre.butter_sos(250.0, "bandpass", freqmin=3.0, freqmax=12.0, corners=4)
    assert sos.shape == (4, 6)
    _, p, _ = signal.sos2zpk(sos)
    assert abs(np.max(np.abs(p)) - 0.9819) < 2e-4
    c = pre.cheby2_lowpass_sos(250.0, 25.0)
    assert c.shape == (6, 6)
    _, p, _ = signal.sos2zpk(c)
    assert abs(np.max(np.abs(p)) - 0.96036) < 2e-4
    hp = pre.butter_sos(250.0, "highpass", freq=0.05, corners=4)
    _, p, _ = signal.sos2zpk(hp)
    assert abs(np.max(np.abs(p)) - 0.99952) < 1e-4


def test_filter_response_and_decimate():
    fs = 250.0
    t = np.arange(200_000) / fs
    inband, outband = np.sin(2 * np.pi * 7.5 * t), np.sin(2 * np.pi * 40.0 * t)
    y = pre.filter_(inband + outband, fs, "bandpass", freqmin=3.0, freqmax=12.0)
    tail = slice(50_000, None)
    assert abs(np.std(y[tail]) / np.std(inband[tail]) - 1.0) < 2e-2
    y2 = pre.filter_(outband, fs, "bandpass", freqmin=3.0, freqmax=12.0)
    assert np.std(y2[tail]) < 3e-3                          # -50 dB at 40 Hz
    zp = pre.filter_(inband, fs, "bandpass", freqmin=3.0, freqmax=12.0, zerophase=True)
    lag = np.argmax(signal.correlate(zp[60_000:120_000], inband[60_000:120_000], mode="same")) - 30_000
    assert lag == 0                                         # zero phase
    d, fs2 = pre.decimate(inband, fs, 5)
    assert fs2 == 50.0 and len(d) == 40_000
    assert abs(np.std(d[10_000:]) / np.std(inband) - 1.0) < 2e-2
    d_nf, _ = pre.decimate(inband, fs, 5, no_filter=True)
    assert np.array_equal(d_nf, inband[::5])


def test_process_chain_config2():
    fs = 250.0
    n = 250 * 600
    rng = np.random.default_rng(1)
    x = rng.standard_normal(n) + 3.0 + 1e-6 * np.arange(n) + np.sin(2 * np.pi * 7.5 * np.arange(n) / fs)
    ops = [("detrend", {"type": "demean"}),
           ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0}),
           ("filter", {"type": "bandpass", "freqmin": 3.0, "freqmax": 12.0}),
           ("decimate", {"factor": 5}),
           ("detrend", {"type": "demean"}),
           ("taper", {"type": "cosine", "max_percentage": 0.05, "max_length": 30.0})]
    y, fs2 = pre.process(x, fs, ops)
    assert fs2 == 50.0 and len(y) == n // 5 and abs(y.mean()) < 1e-3 * np.std(y) + 1e-12
    assert y[0] == 0.0 and abs(y[-1]) < 1e-12


# ---- sampling-rate changes and the CCF post-processing restatements (known answers) ----
def test_resample_known_answers():
    from oracle import preprocess_oracle as pre
    fs, new_fs, n = 40.0, 50.0, 4000
    t = np.arange(n) / fs
    f0 = 1.0
    y, fo = pre.resample(np.sin(2 * np.pi * f0 * t), fs, new_fs)
    assert fo == new_fs and len(y) == int(n / (fs / new_fs)) == 5000
    # a bin-centred sinusoid keeps its frequency; its amplitude is the Hann spectral window at f0
    gain = 0.5 + 0.5 * np.cos(np.pi * f0 / (fs / 2))
    ref = gain * np.sin(2 * np.pi * f0 * np.arange(5000) / new_fs)
    assert np.max(np.abs(y - ref)) < 1e-9
    # no window: exact band-limited interpolation; same rate: identity
    y, _ = pre.resample(np.sin(2 * np.pi * f0 * t), fs, new_fs, window=None)
    assert np.max(np.abs(y - np.sin(2 * np.pi * f0 * np.arange(5000) / new_fs))) < 1e-9
    x = np.random.default_rng(0).standard_normal(1001)
    y, _ = pre.resample(x, 20.0, 20.0, window=None)
    assert np.max(np.abs(y - x)) < 1e-12
    # downsampling by 5 keeps int(npts / 5) samples
    assert len(pre.resample(np.zeros(21_601), 250.0, 50.0)[0]) == 4320


def test_lanczos_known_answers():
    from oracle import preprocess_oracle as pre
    x = np.random.default_rng(1).standard_normal(500)
    # same rate, no shift: the kernel is 1 at 0 and 0 at the other integers -> identity
    y, _ = pre.interpolate_lanczos(x, 10.0, 10.0, a=5)
    assert len(y) == 500 and np.max(np.abs(y - x)) < 1e-12
    # output length: floor(duration / new_dt) + 1
    assert len(pre.interpolate_lanczos(x, 10.0, 25.0, a=5)[0]) == int(np.floor(49.9 * 25.0)) + 1
    # a slow sinusoid is reproduced between the samples
    t = np.arange(2000) / 40.0
    y, _ = pre.interpolate_lanczos(np.sin(2 * np.pi * 0.5 * t), 40.0, 50.0, a=20)
    tn = np.arange(len(y)) / 50.0
    assert np.max(np.abs(y[100:-100] - np.sin(2 * np.pi * 0.5 * tn)[100:-100])) < 5e-4   # Lanczos kernel ripple


def test_signal_oracle_known_answers():
    from oracle import signal_oracle as so
    n, fs = 2000, 50.0
    x = np.random.default_rng(2).standard_normal((2, n))
    # integer-sample delay == circular roll; zero delay == identity
    assert np.max(np.abs(so.timeshift(x, 5 / fs, 0.0, 1 / fs) - np.roll(x, 5, axis=-1))) < 1e-12
    assert np.max(np.abs(so.timeshift(x, 0.0, 0.0, 1 / fs) - x)) < 1e-13
    assert so.timeshift(x[:, :1999], 0.0, 0.0, 1 / fs).shape[-1] == 1998       # odd length: one sample trimmed
    # window: ones in the middle, periodic Hann halves of nmax samples at both ends
    w = so.window(n, fs, max_length=5.0)
    assert w[0] == 0.0 and np.all(w[250:-250] == 1.0) and abs(w[125] - 0.5) < 1e-12
    # zero-phase high-pass removes the mean and keeps a 10 Hz line's phase
    t = np.arange(n) / fs
    y = so.filter(np.sin(2 * np.pi * 10 * t) + 4.0, fs, 3.0)
    assert abs(np.mean(y[200:-200])) < 1e-3
    assert np.max(np.abs(y[200:-200] - np.sin(2 * np.pi * 10 * t)[200:-200])) < 5e-3
    r, s, nn, idx = so.snr(np.r_[np.ones(10), [-7.0], np.ones(9)], np.r_[np.zeros(8, bool), np.ones(12, bool)],
                           np.r_[np.ones(8, bool), np.zeros(12, bool)])
    assert s == 7.0 and nn == 1.0 and idx == 10 and r == 7.0


def test_time_based_trim_matches_the_obspy_restatement():
    """Host arithmetic of ('trim', {starttime, endtime}) in the op compiler against the oracle's
    Trace._ltrim / _rtrim restatement, incl. the buffered-window case of stream/client.py:499-519."""
    from xcorr_b200.stream.process import PreprocessProgram, trim_samples, inject_dynamic_parameters
    rng = np.random.default_rng(11)
    for _ in range(300):
        n = int(rng.integers(5, 400))
        fs = float(rng.choice([1.0, 10.0, 50.0, 250.0]))
        t0 = float(rng.uniform(-20, 20))
        a = t0 + float(rng.uniform(-5, 0.9 * n / fs))
        b = a + float(rng.uniform(0.01, 1.5 * n / fs))
        x = np.arange(n, dtype=np.float64)
        y, t1 = pre.trim_time(x, fs, t0, a, b)
        if len(y) < 1:
            with pytest.raises(ValueError):
                trim_samples(n, fs, t0, a, b)
            continue
        s, k = trim_samples(n, fs, t0, a, b)
        assert (s, k) == (int(y[0]), len(y)) and abs(t1 - (t0 + s / fs)) < 1e-12
    # the reference's window: 24 h + 5 % buffer on both sides at 250 Hz, decimated by 5, then ('trim', {})
    wl, fs = 86400.0, 250.0
    n_raw = int(1.1 * wl * fs) + 1
    ops = [("decimate", {"factor": 5}), ("trim", {})]
    ops = [(o, inject_dynamic_parameters(o, p, starttime=0.0, endtime=wl)) for o, p in ops]
    prog = PreprocessProgram(ops, n_raw, fs, trace_start=-0.05 * wl)
    assert (prog.npts_out, prog.fs_out, prog.trace_start_out) == (int(wl * 50) + 1, 50.0, 0.0)
    assert inject_dynamic_parameters("trim", {"start": 3}, starttime=0.0, endtime=1.0) == {"start": 3}


def test_scan_host_design_carry_window_and_chunk_model():
    """Host side of the chunked scan (stream/process.py): the state-transition matrix over a chunk reproduces
    scipy's sosfilt across chunk boundaries, `carry_window` is a valid certificate (||P^K||_inf <= 2^-80, and the
    truncated recurrence reproduces the true chunk start states), slowly decaying filters get 0, and decimating
    chunk lengths are whole periods and multiples of 8 samples."""
    from xcorr_b200.stream import process as sp
    sos = np.vstack([pre.butter_sos(250.0, "bandpass", freqmin=3.0, freqmax=12.0, corners=4),
                     pre.cheby2_lowpass_sos(250.0, 25.0)])
    chunk = 480
    P, _ = sp._carry_matrices(sos, chunk)
    K = sp._carry_window(P)
    assert 1 <= K <= 16
    assert np.abs(np.linalg.matrix_power(P, K)).sum(axis=1).max() <= 2.0 ** -80
    # true chunk start states from one long sosfilt run vs the windowed recurrence s <- P s + z over K chunks
    rng = np.random.default_rng(5)
    nchunks = 3 * K
    x = rng.standard_normal(nchunks * chunk)
    S = sos.shape[0]
    zi = np.zeros((S, 2))
    starts, zfinal = [], []
    for c in range(nchunks):
        starts.append(zi.reshape(-1).copy())
        _, zi = signal.sosfilt(sos, x[c * chunk:(c + 1) * chunk], zi=zi)
        _, z0 = signal.sosfilt(sos, x[c * chunk:(c + 1) * chunk], zi=np.zeros((S, 2)))
        zfinal.append(z0.reshape(-1))
    for c in range(K, nchunks):
        s = np.zeros(2 * S)
        for k in range(c - K, c):
            s = P @ s + zfinal[k]
        assert np.max(np.abs(s - starts[c])) <= 1e-12 * max(1.0, np.max(np.abs(starts[c])))
    slow = pre.butter_sos(250.0, "highpass", freq=0.05, corners=4)
    assert sp._carry_window(sp._carry_matrices(slow, chunk)[0]) == 0
    prog = sp.PreprocessProgram([("decimate", {"factor": 5})], 100_000, 250.0)
    prog._host_design()
    f = [st[1].f for st in prog.steps if st[0] == "stage" and st[1].f is not None][0]
    assert f.chunk % 5 == 0 and f.chunk % 8 == 0 and f.carry_window >= 1
