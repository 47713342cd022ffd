"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle and the golden vectors generated from the reference itself.

Tolerances (BASELINE.json north_star): argmax lag bit-exact; float64 within
1e-12 of the peak; float32 within 1e-5 of the peak.
"""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


@pytest.fixture(scope="module")
def xb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import xcorr_b200
    return xcorr_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import cc_oracle
    return cc_oracle


def test_cc_matches_reference_golden_vectors(xb, golden):
    checked = 0
    for name, kw in zip(golden["_cases"].tolist(), golden["_kwargs"].tolist()):
        kw = json.loads(kw)
        if "dtype" in kw:
            kw["dtype"] = np.dtype(kw["dtype"])
        x, y, r = golden[name + "_x"], golden[name + "_y"], golden[name + "_r"]
        got = xb.cc(x, y, unbiased=False, **kw)
        assert got.shape == r.shape and got.dtype == r.dtype, name
        if name == "zero":
            assert np.all(np.isnan(got)) and np.all(np.isnan(r))
            continue
        peak = np.max(np.abs(r))
        assert np.max(np.abs(got - r)) <= TOL[r.dtype] * peak, name
        if r.dtype == np.float64 and len(r) > 8:
            # (the delta case peaks at -1: compare the location of the extremum)
            assert np.argmax(np.abs(got)) == np.argmax(np.abs(r)), name
        checked += 1
    assert checked >= 29


def test_shift_helpers_and_unbiased(xb, orc, golden):
    x, y = golden["shift7_x"], golden["shift7_y"]
    assert xb.compute_shift(x, y, unbiased=False) == -7
    s, m = xb.compute_shift_and_max(x, y, 0.02, unbiased=False)
    assert s == golden["shift7_extract_max"][0]
    assert abs(m - golden["shift7_extract_max"][1]) < 1e-12
    ub = xb.cc(x, y, unbiased=True)
    ref = orc.cc(x, y, unbiased=True)
    assert np.max(np.abs(ub - ref)) <= 1e-12 * np.max(np.abs(ref))
    with pytest.raises(AssertionError):
        xb.cc(np.zeros(3), np.zeros(4))
    with pytest.raises(AssertionError):
        xb.cc(np.zeros(3), np.zeros(3, dtype=np.float32))


MENU = [64, 80, 96, 100, 120, 128, 144, 150, 160, 180, 192, 200, 225, 240, 256, 300, 320, 400]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_every_kernel_size_in_every_role(xb, orc, dtype):
    import torch
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    for N in MENU:
        # single pass: row kernel with real-input load + pair product + lag epilogue
        n = N - 3
        rec = orc.synthetic_receivers(3, n, seed=N, delay=5)
        plan = xb.CorrelationPlan(n, 3, [(0, 1), (2, 0)], dtype=tdt, dims=(N,))
        got = plan(torch.from_numpy(rec.astype(dtype)).cuda()).cpu().numpy()[:, 0]
        for q, (i, j) in enumerate([(0, 1), (2, 0)]):
            ref = orc.cc(rec[i], rec[j], unbiased=False)
            assert np.max(np.abs(got[q] - ref)) <= TOL[np.dtype(dtype)] * np.max(np.abs(ref)), (N, "row")
        # three passes: N as first (column, real in / epilogue) and middle column pass
        dims = (N, N, 64)
        n = N * N * 64 - 11
        rec = orc.synthetic_receivers(2, n, seed=N + 1, delay=137)
        plan = xb.CorrelationPlan(n, 2, [(0, 1)], dtype=tdt, dims=dims, index_min=n - 1 - 500, index_max=n - 1 + 1500)
        got = plan(torch.from_numpy(rec.astype(dtype)).cuda()).cpu().numpy()[0, 0]
        ref = orc.cc_smooth(rec[0], rec[1])
        peak = np.max(np.abs(ref))
        ref = ref[n - 1 - 500:n - 1 + 1500]
        assert np.max(np.abs(got - ref)) <= TOL[np.dtype(dtype)] * peak, (N, "col")
        assert np.argmax(got) == np.argmax(ref) == 500 + 137
        del plan
        torch.cuda.empty_cache()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_batched_plan_all_pairs_windows_crop(xb, orc, dtype):
    import torch
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    n, R, W = 20_000, 4, 3
    pairs = xb.all_pairs(R, auto=True)
    rec = np.stack([orc.synthetic_receivers(R, n, seed=100 + w, delay=61) for w in range(W)], axis=1)  # [R, W, n]
    i0, i1 = orc.lag_clip_indices(n, 0.02, (-50.0, 120.0))
    x = torch.from_numpy(rec.astype(dtype)).cuda()
    plan = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=tdt, index_min=i0, index_max=i1)
    ccf = plan(x)
    idx, val = plan.argmax(ccf)
    got = ccf.cpu().numpy()
    assert got.shape == (len(pairs), W, i1 - i0)
    for q, (i, j) in enumerate(pairs):
        for w in range(W):
            ref = orc.cc(rec[i, w], rec[j, w], unbiased=False)
            peak = np.max(np.abs(ref))
            ref = ref[i0:i1]
            assert np.max(np.abs(got[q, w] - ref)) <= TOL[np.dtype(dtype)] * peak
            assert int(idx[q, w]) == int(np.argmax(ref)) == (n - 1 - i0) + 61 * (j - i)
            assert float(val[q, w]) == got[q, w, int(idx[q, w])]
    # unbiased weights fused into the epilogue == signal/unbias.py:85
    plan_u = xb.CorrelationPlan(n, R, pairs[:3], n_windows=W, dtype=tdt, index_min=i0, index_max=i1, unbiased=True)
    got_u = plan_u(x).cpu().numpy()
    for q, (i, j) in enumerate(pairs[:3]):
        ref = orc.process_cell(rec[i, 0], rec[j, 0], n, i0, i1, dtype=np.float64, bias_correct=True)
        assert np.max(np.abs(got_u[q, 0] - ref)) <= 20 * TOL[np.dtype(dtype)] * np.max(np.abs(ref))
    # stacking over windows
    plan_s = xb.CorrelationPlan(n, R, pairs[:4], n_windows=W, dtype=tdt, index_min=i0, index_max=i1, stack=True)
    got_s = plan_s(x).cpu().numpy()
    for q in range(4):
        ref = got[q].astype(np.float64).sum(axis=0)
        assert np.max(np.abs(got_s[q] - ref)) <= 4 * TOL[np.dtype(dtype)] * np.max(np.abs(ref))
    # normalize=False
    plan_n = xb.CorrelationPlan(n, R, [(0, 1)], n_windows=W, dtype=tdt, normalize=False)
    got_n = plan_n(x).cpu().numpy()
    ref = orc.cc(rec[0, 1], rec[1, 1], normalize=False, unbiased=False)
    assert np.max(np.abs(got_n[0, 1] - ref)) <= TOL[np.dtype(dtype)] * np.max(np.abs(ref))


def test_spectra_are_bit_reproducible(xb, orc):
    import torch
    rec = orc.synthetic_receivers(3, 100_000, seed=9)
    x = torch.from_numpy(rec.astype(np.float32)).cuda()
    plan = xb.CorrelationPlan(100_000, 3, xb.all_pairs(3), dtype=torch.float32)
    a = plan(x).clone()
    ss = plan.sumsq.clone()
    b = plan(x)
    assert torch.equal(a, b) and torch.equal(ss, plan.sumsq)
    assert np.allclose(ss.cpu().numpy(), np.sum(rec.astype(np.float32).astype(np.float64) ** 2, axis=1), rtol=1e-7)


def test_argmax_semantics(xb):
    import torch
    plan = xb.CorrelationPlan(100, 2, [(0, 1)])
    rng = np.random.default_rng(0)
    a = rng.standard_normal((7, 100_003)).astype(np.float32)
    a[1, 500] = a[1, 90_000] = 9.0          # tie -> first index
    a[2, :] = -1.0                           # all equal
    a[3, 77_777] = np.nan                    # NaN wins, like np.argmax
    a[4, 0] = 50.0
    a[5, -1] = 50.0
    idx, val = plan.argmax(torch.from_numpy(a).cuda())
    assert idx.cpu().tolist() == np.argmax(a, axis=1).tolist()
    v = val.cpu().numpy()
    assert np.isnan(v[3]) and v[1] == 9.0 and v[4] == 50.0
    a64 = rng.standard_normal((3, 1000))
    idx, _ = plan.argmax(torch.from_numpy(a64).cuda())
    assert idx.cpu().tolist() == np.argmax(a64, axis=1).tolist()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_day_at_50hz_against_oracle(xb, orc, dtype):
    """BASELINE config 1 shape: 24 h @ 50 Hz, n = 4.32 M, production crop (0, 9 h)."""
    import torch
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    n = 4_320_000
    rec = orc.synthetic_receivers(2, n, seed=1000, delay=137)
    i0, i1 = n - 1, n - 1 + 1_620_000
    plan = xb.CorrelationPlan(n, 2, [(0, 1)], dtype=tdt, index_min=i0, index_max=i1)
    assert n <= int(np.prod(plan.dims)) <= 1.03 * n
    got = plan(torch.from_numpy(rec.astype(dtype)).cuda()).cpu().numpy()[0, 0]
    ref = orc.cc_smooth(rec[0], rec[1])
    peak = np.max(np.abs(ref))
    ref = ref[i0:i1]
    assert np.argmax(got) == np.argmax(ref) == 137
    assert np.max(np.abs(got - ref)) <= TOL[np.dtype(dtype)] * peak
    if dtype == np.float64:
        # full cc() through the host-buffer C entry, all 2n-1 lags
        full = xb.cc(rec[0], rec[1], unbiased=False)
        assert np.max(np.abs(full[i0:i1] - ref)) <= 1e-12 * peak
        assert xb.extract_shift(full) == 137


def test_day_at_250hz_full_size(xb, orc):
    """BASELINE configs 2-3 shape: n = 21.6 M, L = 43.2 M, K = 8.1 M lags, float32;
    oracle = smooth-length rfft restatement (the reference's own Bluestein path
    takes ~75 s per pair) plus size-independent properties."""
    import torch
    n = 21_600_000
    rec = orc.synthetic_receivers(3, n, seed=3000, delay=137).astype(np.float32)
    i0, i1 = n - 1, n - 1 + 8_100_000
    pairs = [(0, 1), (0, 2), (1, 2), (1, 0)]
    plan = xb.CorrelationPlan(n, 3, pairs, dtype=torch.float32, index_min=i0, index_max=i1)
    assert n <= int(np.prod(plan.dims)) <= 1.03 * n
    x = torch.from_numpy(rec).cuda()
    ccf = plan(x)
    idx, val = plan.argmax(ccf)
    assert idx[:, 0].cpu().tolist()[:3] == [137, 274, 137]
    got = ccf[0, 0].cpu().numpy()
    ref = orc.cc_smooth(rec[0].astype(np.float64), rec[1].astype(np.float64))
    peak = np.max(np.abs(ref))
    assert np.max(np.abs(got - ref[i0:i1])) <= 1e-5 * peak
    # symmetry: cc(y, x)[k] == cc(x, y)[-k]  -> zero lag agrees between (0,1) and (1,0)
    assert abs(float(ccf[0, 0, 0]) - float(ccf[3, 0, 0])) <= 1e-6
    # linearity in the first argument: cc(a*x0 + x2', x1) -- checked through un-normalised plan
    plan_n = xb.CorrelationPlan(n, 3, [(0, 1), (2, 1)], dtype=torch.float32, index_min=i0, index_max=i0 + 4096, normalize=False)
    base = plan_n(x).double()
    x2 = x.clone()
    x2[0] = 2.0 * x[0] + x[2]
    mix = plan_n(x2).double()
    lin = 2.0 * base[0] + base[1]
    assert float((mix[0] - lin).abs().max()) <= 1e-5 * float(lin.abs().max())


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_block_mode_equals_classic_transform(xb, orc, dtype):
    """Cropped non-negative lags: the planner cuts the second signal into blocks
    and uses shorter transforms; results must equal the
    single full-length transform and the oracle."""
    import torch
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    n, R, W = 120_000, 6, 2
    pairs = xb.all_pairs(R)
    rec = np.stack([orc.synthetic_receivers(R, n, seed=400 + w, delay=61) for w in range(W)], axis=1)
    x = torch.from_numpy(rec.astype(dtype)).cuda()
    for lag0, K in ((0, 30_000), (1_000, 25_000)):
        i0 = n - 1 + lag0
        auto = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=tdt, index_min=i0, index_max=i0 + K)
        classic = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=tdt, index_min=i0, index_max=i0 + K, n_blocks=1)
        assert auto.n_blocks >= 2 and classic.n_blocks == 1
        assert 2 * int(np.prod(auto.dims)) < 2 * n            # transforms shorter than the window pair
        a, c = auto(x), classic(x)
        idx_a, _ = auto.argmax(a)
        idx_c, _ = classic.argmax(c)
        assert torch.equal(idx_a, idx_c)
        a, c = a.cpu().numpy(), c.cpu().numpy()
        tol = TOL[np.dtype(dtype)]
        for q, (i, j) in enumerate(pairs):
            for w in range(W):
                ref = orc.cc(rec[i, w], rec[j, w], unbiased=False)
                peak = np.max(np.abs(ref))
                assert np.max(np.abs(a[q, w] - ref[i0:i0 + K])) <= tol * peak
                assert np.max(np.abs(c[q, w] - ref[i0:i0 + K])) <= tol * peak
        assert np.allclose(auto.sumsq.cpu().numpy(), classic.sumsq.cpu().numpy(), rtol=1e-6)
    # forced geometry + unbiased weights
    forced = xb.CorrelationPlan(n, R, pairs[:3], n_windows=W, dtype=tdt, index_min=n - 1, index_max=n - 1 + 20_000,
                                dims=(64, 64, 16 * 4) if False else (200, 200), n_blocks=2, block_len=60_000, unbiased=True)
    got = forced(x).cpu().numpy()
    for q, (i, j) in enumerate(pairs[:3]):
        ref = orc.cc(rec[i, 1], rec[j, 1], unbiased=True)
        assert np.max(np.abs(got[q, 1] - ref[n - 1:n - 1 + 20_000])) <= 20 * TOL[np.dtype(dtype)] * np.max(np.abs(ref[n - 1:n - 1 + 20_000]))
    # symmetric and negative-only lag windows: segments start before the signal (zeros there)
    for i0, K in ((n - 1 - 20_000, 40_001), (n - 1 - 50_000, 30_000), (0, 25_000)):
        auto = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=tdt, index_min=i0, index_max=i0 + K)
        assert auto.n_blocks >= 2
        got = auto(x).cpu().numpy()
        for q, (i, j) in list(enumerate(pairs))[::4]:
            ref = orc.cc(rec[i, 0], rec[j, 0], unbiased=False)
            assert np.max(np.abs(got[q, 0] - ref[i0:i0 + K])) <= TOL[np.dtype(dtype)] * np.max(np.abs(ref)), (i0, q)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_staged_pair_product_chunks_stacking_and_groups(xb, orc, dtype):
    """The staged / grouped pair product (16 x R2 row axes) in the roles the benchmark does not cover:
    several launch chunks (small scratch), stacked windows with per-window norms, repeated and
    auto pairs, a single-pass plan, and classic (full-lag) plans."""
    import torch
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    tol = TOL[np.dtype(dtype)]
    n, R, W = 30_000, 6, 3
    pairs = xb.all_pairs(R, auto=True) + [(3, 1), (3, 1)]
    rec = np.stack([orc.synthetic_receivers(R, n, seed=700 + w, delay=23) for w in range(W)], axis=1)   # [R, W, n]
    x = torch.from_numpy(rec.astype(dtype)).cuda()
    refs = {(i, j, w): orc.cc(rec[i, w].astype(dtype).astype(np.float64), rec[j, w].astype(dtype).astype(np.float64),
                              unbiased=False) for (i, j) in set(pairs) for w in range(W)}
    per_job = 128 * 256 * (8 if dtype == np.float32 else 16)
    for kw in (dict(dims=(128, 256), max_work_bytes=5 * per_job),         # 69 jobs in chunks of 5
               dict(dims=(128, 240)), dict(dims=(160, 192))):
        plan = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=tdt, n_blocks=1, **kw)
        assert plan.n_blocks == 1
        got = plan(x).cpu().numpy()
        for q, (i, j) in enumerate(pairs):
            for w in range(W):
                ref = refs[(i, j, w)]
                assert np.max(np.abs(got[q, w] - ref)) <= tol * np.max(np.abs(ref)), (kw, q, w)
    # stacking: sum over windows of the per-window normalised CCFs, one inverse per pair
    plan = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=tdt, stack=True, dims=(128, 256))
    got = plan(x).cpu().numpy()
    for q, (i, j) in enumerate(pairs):
        ref = sum(refs[(i, j, w)] for w in range(W))
        assert np.max(np.abs(got[q] - ref)) <= tol * np.max(np.abs(ref)), q
    # cropped window around zero lag through blocks (segments start before the signal) with groups
    i0 = n - 1 - 2_000
    plan = xb.CorrelationPlan(n, R, pairs, n_windows=W, dtype=tdt, index_min=i0, index_max=i0 + 6_001,
                              dims=(80, 160), n_blocks=2, block_len=15_008)
    got = plan(x).cpu().numpy()
    for q, (i, j) in enumerate(pairs):
        for w in range(W):
            ref = refs[(i, j, w)]
            assert np.max(np.abs(got[q, w] - ref[i0:i0 + 6_001])) <= tol * np.max(np.abs(ref)), (q, w)
    # single-pass plan (one row): n = 120 -> 256 points
    small = orc.synthetic_receivers(5, 120, seed=9, delay=3)
    plan = xb.CorrelationPlan(120, 5, xb.all_pairs(5), dtype=tdt, dims=(256,))
    got = plan(torch.from_numpy(small.astype(dtype)).cuda().unsqueeze(1)).cpu().numpy()
    for q, (i, j) in enumerate(xb.all_pairs(5)):
        ref = orc.cc(small[i].astype(dtype).astype(np.float64), small[j].astype(dtype).astype(np.float64), unbiased=False)
        assert np.max(np.abs(got[q, 0] - ref)) <= tol * np.max(np.abs(ref))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_quarter_warp_row_axis_with_many_groups(xb, orc, dtype):
    """8 x 8 row axis (the 64^3 transform of config 5): warp-synchronous row pass with a quarter-warp per row and
    the staged pair product with two mirrored row pairs per warp.  40 receivers / 780 pairs give more job groups
    than the kernel's shared-memory record table holds (records mailed with the rows); the number of row pairs is
    odd (the last item holds one pair only); the first and the middle row mirror themselves."""
    import torch
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    tol = TOL[np.dtype(dtype)]
    n, R = 4_000, 40
    pairs = xb.all_pairs(R)
    rec = orc.synthetic_receivers(R, n, seed=4242, delay=5)
    x = torch.from_numpy(rec.astype(dtype)).cuda().unsqueeze(1)
    for kw in (dict(dims=(64, 64)), dict(dims=(64, 64), max_work_bytes=100 * 64 * 64 * (8 if dtype == np.float32 else 16))):
        plan = xb.CorrelationPlan(n, R, pairs, dtype=tdt, n_blocks=1, **kw)
        assert plan.dims == (64, 64) and plan.n_groups > 0, (plan.dims, plan.n_groups)
        got = plan(x).cpu().numpy()
        idx, val = plan.correlate_argmax()
        torch.cuda.synchronize()
        for q, (i, j) in enumerate(pairs):
            ref = orc.cc(rec[i].astype(dtype).astype(np.float64), rec[j].astype(dtype).astype(np.float64), unbiased=False)
            assert np.max(np.abs(got[q, 0] - ref)) <= tol * np.max(np.abs(ref)), (kw, q)
            assert int(idx[q, 0]) == int(np.argmax(got[q, 0])), (kw, q)      # fused arg-max == arg-max of the stored CCF
    # a lone pair keeps the plain warp kernel (no groups)
    plan = xb.CorrelationPlan(n, 2, [(0, 1)], dtype=tdt, dims=(64, 64), n_blocks=1)
    assert plan.n_groups == 0
    got = plan(x[:2]).cpu().numpy()
    ref = orc.cc(rec[0].astype(dtype).astype(np.float64), rec[1].astype(dtype).astype(np.float64), unbiased=False)
    assert np.max(np.abs(got[0, 0] - ref)) <= tol * np.max(np.abs(ref))


def test_graphed_step_equals_direct_launches(xb, orc):
    """Config 1 shape in small: the CUDA-graph replay of forward + correlate gives bit-identical CCFs, also after
    the input buffer is refilled in place."""
    import torch
    n = 200_000
    rec = orc.synthetic_receivers(2, n, seed=77, delay=137)
    x = torch.from_numpy(rec).cuda()
    plan = xb.CorrelationPlan(n, 2, [(0, 1)], dtype=torch.float64, index_min=n - 1, index_max=n - 1 + 60_000)
    direct = plan(x).clone()
    replay, out = plan.graphed(x)
    replay()
    torch.cuda.synchronize()
    assert torch.equal(out, direct)
    rec2 = orc.synthetic_receivers(2, n, seed=78, delay=55)
    x.copy_(torch.from_numpy(rec2).cuda())
    replay()
    torch.cuda.synchronize()
    ref = orc.cc(rec2[0], rec2[1], unbiased=False)
    got = out[0, 0].cpu().numpy()
    assert np.argmax(got) == 55 and np.max(np.abs(got - ref[n - 1:n - 1 + 60_000])) <= 1e-12 * np.max(np.abs(ref))
