"""Value-level parity at the NAMED sizes of BASELINE.json (configs 2, 3, 4, 5), through the
C ABI, against the oracle -- in float32 (<= 1e-5 of the peak) and float64 (<= 1e-12),
arg-max lag exact.  Cell semantics: core/process.py:155-162 (cc of the first npts
samples, crop [index_min:index_max]).

* config 3 is run with EXACTLY the plan and the inputs bench.py times (16 receivers, all
  120 pairs, automatic plan: 2 blocks, grouped / staged pair product); 8 cells are compared in
  full with the smooth-length oracle (the reference's own 2n-1 Bluestein transform takes ~75 s
  per cell; `cc_smooth` agrees with it to 1e-15, tests/test_oracle.py) and the arrival is checked
  for all 120 pairs.
* config 2: 5 receivers of 24 h @ 250 Hz raw samples through bench.py's op-list (detrend, taper, 3-12 Hz
  band-pass, decimate 5, detrend, taper) against the scipy restatement of the obspy chain (oracle unpinned, see
  oracle/preprocess_oracle.py), then all 10 pairs with lags (0, 9 h) against the smooth-length cc oracle.
* config 4: 4 receivers, 64 stacked 1 h windows at 250 Hz against the SUM of per-window oracle CCFs.
* config 5: 1024 x 128 spectrograms against oracle.correlate2d, full cc2 and (value, dF, dT).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {"f32": 1e-5, "f64": 1e-12}


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import xcorr_b200 as xb
    from oracle import cc_oracle as orc
    return torch, xb, orc


def _tdt(torch, dtype):
    return torch.float32 if dtype == "f32" else torch.float64


# ---------------------------------------------------------------------------
# config 3: the benchmark's own plan
# ---------------------------------------------------------------------------
CFG3_N, CFG3_R, CFG3_K = 21_600_000, 16, 8_100_000
CFG3_CELLS = [(0, 1), (0, 7), (0, 15), (1, 7), (1, 15), (7, 8), (7, 15), (8, 15)]   # receivers {0, 1, 7, 8, 15}


@pytest.fixture(scope="module")
def cfg3(env):
    """bench.py's inputs (float32 values) and the float64 oracle of 8 cells, computed once for both dtypes
    (the float64 run gets the same float32-representable samples)."""
    torch, xb, orc = env
    from bench import make_inputs
    dev = torch.device("cuda:0")
    x = make_inputs(torch, CFG3_R, CFG3_N, dev, torch.float32, seed=3000)
    n, L = CFG3_N, 2 * CFG3_N
    recv = sorted({r for c in CFG3_CELLS for r in c})
    spec, norm = {}, {}
    for r in recv:
        h = x[r].cpu().numpy().astype(np.float64)
        spec[r] = np.fft.rfft(h, L)
        norm[r] = np.linalg.norm(h)
    refs = {}
    for (i, j) in CFG3_CELLS:
        # util/cc.py:75-78 on a smooth length (oracle.cc_smooth, spectra shared between the cells)
        r = np.fft.irfft(spec[i] * np.conjugate(spec[j]) / (norm[i] * norm[j]), L)
        full_peak = max(np.max(np.abs(r[:n])), np.max(np.abs(r[L - (n - 1):])))
        refs[(i, j)] = (r[:CFG3_K].copy(), full_peak)         # lags 0 .. K-1 = indices [n-1, n-1+K)
    del spec
    return x, refs


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_cfg3_benchmark_plan_cells_in_full(env, cfg3, dtype):
    torch, xb, orc = env
    x32, refs = cfg3
    n, K = CFG3_N, CFG3_K
    pairs = xb.all_pairs(CFG3_R)
    x = x32 if dtype == "f32" else x32.double()
    plan = xb.CorrelationPlan(n, CFG3_R, pairs, n_windows=1, dtype=_tdt(torch, dtype),
                              index_min=n - 1, index_max=n - 1 + K, device=x.device)
    # the plan bench.py times: block mode, 256-point row axis, grouped / staged pair product
    assert plan.n_blocks == 2 and plan.dims[-1] == 256 and plan.n_groups > 0, (plan.dims, plan.n_blocks, plan.n_groups)
    ccf = plan(x)
    idx, val = plan.argmax(ccf)
    torch.cuda.synchronize()
    assert idx[:, 0].cpu().tolist() == [137 * (j - i) for i, j in pairs]
    worst = 0.0
    for (i, j) in CFG3_CELLS:
        ref, peak = refs[(i, j)]
        got = ccf[pairs.index((i, j)), 0].cpu().numpy().astype(np.float64)
        assert int(np.argmax(got)) == int(np.argmax(ref)) == 137 * (j - i)
        err = float(np.max(np.abs(got - ref))) / peak
        worst = max(worst, err)
        assert err <= TOL[dtype], ((i, j), err)
    print(f"cfg3 {dtype}: dims {plan.dims} blocks {plan.n_blocks} groups {plan.n_groups} worst {worst:.2e} of peak")


def test_cfg3_three_windows_chunked_equals_single_windows(env):
    """3 windows x 120 pairs in one plan (several launch chunks of the staged pair product): arrival exact
    for every cell and a window processed alone gives bit-identical CCFs."""
    torch, xb, orc = env
    from bench import make_inputs
    dev = torch.device("cuda:0")
    n, R, W, K = CFG3_N, CFG3_R, 3, CFG3_K
    pairs = xb.all_pairs(R)
    x = torch.stack([make_inputs(torch, R, n, dev, torch.float32, 10 + w) for w in range(W)], dim=1).contiguous()
    plan = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=torch.float32, index_min=n - 1, index_max=n - 1 + K,
                              max_work_bytes=12 << 30)
    assert plan.jobs_in_flight < plan.n_jobs                     # more than one launch chunk
    ccf = plan(x)
    idx, val = plan.argmax(ccf)
    exp = torch.tensor([137 * (j - i) for i, j in pairs], device=dev).view(-1, 1).expand(-1, W)
    assert bool((idx == exp).all())
    assert 0.7 < float(val.min()) and float(val.max()) < 0.9     # 1 / (1 + 0.5^2) = 0.8
    plan1 = xb.CorrelationPlan(n, R, pairs, n_windows=1, dtype=torch.float32, index_min=n - 1, index_max=n - 1 + K)
    c1 = plan1(x[:, 1:2].contiguous())
    assert torch.equal(c1[:, 0], ccf[:, 1])


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_sixteen_receivers_midsize_every_cell(env, dtype):
    """Same structure as config 3 (16 receivers, 120 pairs, cropped lags -> blocks, 16 x R2 row axis,
    grouped kernel) at a size where EVERY cell is compared with the reference restatement."""
    torch, xb, orc = env
    n, R, K = 120_000, 16, 16_000
    pairs = xb.all_pairs(R)
    rec = orc.synthetic_receivers(R, n, seed=3001, delay=137).astype(np.float32)
    x = torch.from_numpy(rec).cuda().to(_tdt(torch, dtype))
    # the benchmark's geometry in small: 2 blocks, middle pass of 150, 256-point row axis (2M = 76800 >= 60000 + K)
    plan = xb.CorrelationPlan(n, R, pairs, dtype=_tdt(torch, dtype), index_min=n - 1, index_max=n - 1 + K,
                              dims=(150, 256), n_blocks=2, block_len=60_000)
    assert plan.n_blocks >= 2 and plan.n_groups > 0 and 3 * plan.n_groups <= len(pairs), (plan.dims, plan.n_groups)
    got = plan(x)[:, 0].cpu().numpy()
    r64 = rec.astype(np.float64)
    for q, (i, j) in enumerate(pairs):
        ref = orc.cc(r64[i], r64[j], unbiased=False)
        peak = np.max(np.abs(ref))
        assert np.max(np.abs(got[q] - ref[n - 1:n - 1 + K])) <= TOL[dtype] * peak, (i, j)
        assert int(np.argmax(got[q])) == 137 * (j - i)


# ---------------------------------------------------------------------------
# config 4: short windows, stacked on the device
# ---------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg4(env):
    torch, xb, orc = env
    n, R, W = 900_000, 4, 64
    rec = np.stack([orc.synthetic_receivers(R, n, seed=4000 + w, delay=137).astype(np.float32) for w in range(W)],
                   axis=1)                                                    # [R, W, n]
    pairs = xb.all_pairs(R)
    L = 1_800_000
    stack = {p: np.zeros(2 * n - 1) for p in pairs}
    for w in range(W):
        h = rec[:, w].astype(np.float64)
        spec = [np.fft.rfft(h[r], L) for r in range(R)]
        nrm = [np.linalg.norm(h[r]) for r in range(R)]
        for (i, j) in pairs:
            # one reference cell per (pair, window), summed afterwards (SURVEY 8d config 4 oracle)
            r = np.fft.irfft(spec[i] * np.conjugate(spec[j]) / (nrm[i] * nrm[j]), L)
            stack[(i, j)] += np.concatenate((r[L - (n - 1):], r[:n]))
    return rec, pairs, stack


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_cfg4_hour_windows_stacked(env, cfg4, dtype):
    torch, xb, orc = env
    rec, pairs, stack = cfg4
    R, W, n = rec.shape
    x = torch.from_numpy(rec).cuda().to(_tdt(torch, dtype))
    plan = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=_tdt(torch, dtype), stack=True)
    out = torch.zeros((len(pairs), plan.n_lags), dtype=plan.dtype, device=x.device)
    plan.forward(x)
    plan.correlate(out=out)
    got = out.cpu().numpy().astype(np.float64)
    for q, p in enumerate(pairs):
        ref = stack[p]
        peak = np.max(np.abs(ref))
        assert int(np.argmax(got[q])) == int(np.argmax(ref)) == (n - 1) + 137 * (p[1] - p[0])
        # W terms are summed: the float32 tolerance is per term
        assert np.max(np.abs(got[q] - ref)) <= TOL[dtype] * peak * (8 if dtype == "f32" else 1), p
    # a second call accumulates (ccf += ...), as the 5-year stack is built batch by batch
    plan.correlate(out=out)
    assert np.max(np.abs(out.cpu().numpy() - 2 * got)) <= 1e-6 * np.max(np.abs(got))


# ---------------------------------------------------------------------------
# config 5: timelapse kernel at 1024 x 128
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_cfg5_spectrogram_pairs_full_size(env, dtype):
    torch, xb, orc = env
    from scipy.ndimage import uniform_filter
    from xcorr_b200 import signal
    rng = np.random.default_rng(5005)
    F, T, D = 1024, 128, 5
    base = uniform_filter(rng.standard_normal((F + 40, T + 40)) ** 2, size=3) + 0.01
    shifts = [(0, 0), (1, 2), (3, -1), (-2, 4), (-4, -3)]
    days = np.stack([base[20 + a:20 + a + F, 20 + b:20 + b + T] +
                     0.02 * uniform_filter(rng.standard_normal((F, T)) ** 2, size=3) for a, b in shifts])
    days = days.astype(np.float32).astype(np.float64)
    pairs = [(i, j) for i in range(D) for j in range(i, D)]
    plan = signal.Correlate2dPlan(F, T, D, pairs, dtype=_tdt(torch, dtype))
    x = torch.from_numpy(days).cuda().to(_tdt(torch, dtype))
    val, dF, dT = plan.correlate_argmax(x)                       # fused: (value, dF, dT) per day pair only
    cc2 = plan(x)                                                # full cc2 kept for parity
    val2, dF2, dT2 = plan.argmax(cc2)
    assert torch.equal(dF, dF2) and torch.equal(dT, dT2) and torch.equal(val, val2)
    for q, (i, j) in enumerate(pairs):
        ref = orc.correlate2d(days[i], days[j])
        a, b = np.unravel_index(np.argmax(ref), ref.shape)       # scripts/timelapse.py:261-268
        assert (int(dF[q]), int(dT[q])) == (a - (F - 1), b - (T - 1)) == \
            (shifts[j][0] - shifts[i][0], shifts[j][1] - shifts[i][1]), (i, j)
        assert abs(float(val[q]) - ref.max()) <= TOL[dtype] * ref.max()
        if q % 4 == 0:
            got = cc2[q].cpu().numpy()
            assert np.max(np.abs(got - ref)) <= TOL[dtype] * np.max(np.abs(ref)), (i, j)


# ---------------------------------------------------------------------------
# config 2: raw 250 Hz traces -> preprocessing chain -> cc at 50 Hz, full size
# ---------------------------------------------------------------------------
CFG2_NRAW, CFG2_R, CFG2_K = 21_600_000, 5, 1_620_000


@pytest.fixture(scope="module")
def cfg2(env):
    """SURVEY 8d raw inputs (common source + noise, 7.5 Hz line, trend, offset) as float32 values, the float64
    oracle chain of every trace and the oracle CCF of all 10 pairs (spectra shared between the cells)."""
    torch, xb, orc = env
    from bench import CFG2_OPS
    from oracle import preprocess_oracle as pre
    fs = 250.0
    src = orc.synthetic_receivers(CFG2_R, CFG2_NRAW, seed=2000, delay=137 * 5)
    t = np.arange(CFG2_NRAW)
    raw = (src + np.sin(2 * np.pi * 7.5 * t / fs) + 1e-6 * t + 3.0).astype(np.float32)
    del src, t
    ref = np.stack([pre.process(row.astype(np.float64), fs, CFG2_OPS)[0] for row in raw])
    n = ref.shape[1]
    L = 2 * n
    spec = [np.fft.rfft(r, L) for r in ref]
    norm = [np.linalg.norm(r) for r in ref]
    cells = {}
    for (i, j) in xb.all_pairs(CFG2_R):
        r = np.fft.irfft(spec[i] * np.conjugate(spec[j]) / (norm[i] * norm[j]), L)
        peak = max(np.max(np.abs(r[:n])), np.max(np.abs(r[L - (n - 1):])))
        cells[(i, j)] = (r[:CFG2_K].copy(), peak)
    return raw, ref, cells, CFG2_OPS


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_cfg2_full_size_chain_and_cells(env, cfg2, dtype):
    torch, xb, orc = env
    from xcorr_b200.stream.process import PreprocessProgram
    raw, ref, cells, ops = cfg2
    dev = torch.device("cuda:0")
    x = torch.from_numpy(raw).to(dev)
    if dtype == "f64":
        x = x.double()
    prog = PreprocessProgram(ops, CFG2_NRAW, 250.0)
    y = prog.run(x)
    assert tuple(y.shape) == ref.shape == (CFG2_R, CFG2_NRAW // 5) and prog.fs_out == 50.0
    got = y.cpu().numpy().astype(np.float64)
    err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
    assert err <= (1e-12 if dtype == "f64" else 2e-6), err
    n = y.shape[1]
    pairs = xb.all_pairs(CFG2_R)
    plan = xb.CorrelationPlan(n, CFG2_R, pairs, dtype=_tdt(torch, dtype), index_min=n - 1, index_max=n - 1 + CFG2_K,
                              device=dev)
    ccf = plan(y)
    worst = 0.0
    for q, (i, j) in enumerate(pairs):
        r, peak = cells[(i, j)]
        g = ccf[q, 0].cpu().numpy().astype(np.float64)
        assert int(np.argmax(g)) == int(np.argmax(r))
        e = float(np.max(np.abs(g - r))) / peak
        worst = max(worst, e)
        assert e <= TOL[dtype], ((i, j), e)
    print(f"cfg2 {dtype}: trace error {err:.2e}, worst cell {worst:.2e} of peak, dims {plan.dims}")
