"""Pin the CPU oracle against vectors produced by the reference itself
(tests/golden/make_golden.py executes /root/reference/xcorr/util/cc.py)."""
import json

import numpy as np
import pytest

from oracle import cc_oracle as orc


def _cases(golden):
    return list(zip(golden["_cases"].tolist(), golden["_kwargs"].tolist()))


def test_cc_matches_reference_vectors(golden):
    n_checked = 0
    for name, kw in _cases(golden):
        kw = json.loads(kw)
        if "dtype" in kw:
            kw["dtype"] = np.dtype(kw["dtype"])
        x, y, r = golden[name + "_x"], golden[name + "_y"], golden[name + "_r"]
        with np.errstate(all="ignore"):
            got = orc.cc(x, y, unbiased=False, **kw)
        assert got.shape == r.shape and got.dtype == r.dtype, name
        # same numpy, same operations in the same order -> bit identical
        assert np.array_equal(got, r, equal_nan=True), name
        n_checked += 1
    assert n_checked >= 30


def test_lag_weight_shift_match_reference(golden):
    for n, d in ((5, 0.5), (6, 0.02), (1000, 0.004)):
        assert np.array_equal(orc.lag(n, d, True), golden[f"lag_n{n}_pad"])
        assert np.array_equal(orc.lag(n, d, False), golden[f"lag_n{n}_nopad"])
    for n in (5, 6, 1000):
        ref_w = golden[f"weight_n{n}"]
        assert len(ref_w) == 2 * n - 3          # the reference's short vector
        assert np.array_equal(orc.weight_reference_length(n), ref_w)
        assert np.array_equal(orc.weight_reference_length(n, clip=3.0),
                              golden[f"weight_n{n}_clip"])
        w = orc.weight(n)
        assert len(w) == 2 * n - 1
        assert np.allclose(w, w[::-1]) and w[n - 1] == 1.0 and w[0] == n
    r = golden["shift7_r"]
    assert orc.extract_shift(r) == golden["shift7_extract"][0] == -7
    assert orc.extract_shift(r, 0.02) == golden["shift7_extract"][1]
    s, m = orc.extract_shift_and_max(r, 0.02)
    assert s == golden["shift7_extract_max"][0] and m == golden["shift7_extract_max"][1]
    x, y = golden["shift7_x"], golden["shift7_y"]
    assert orc.compute_shift(x, y, unbiased=False) == golden["shift7_compute"][0]


def test_smooth_length_oracle_agrees_with_reference_path(golden):
    for name in ("synth_01", "synth_02", "rand_f64_n4096", "rand_f64_n257"):
        x, y, r = golden[name + "_x"], golden[name + "_y"], golden[name + "_r"]
        got = orc.cc_smooth(x, y)
        assert np.argmax(got) == np.argmax(r)
        assert np.max(np.abs(got - r)) <= 1e-13 * np.max(np.abs(r))


def test_known_answers():
    # delta pulses: x=2*d[30], y=-3*d[10] -> single peak -1 at lag +20
    a = np.zeros(50); b = np.zeros(50); a[30] = 2.0; b[10] = -3.0
    r = orc.cc(a, b, unbiased=False)
    assert np.argmin(r) - 49 == 20 and abs(r.min() + 1.0) < 1e-15
    assert np.sum(np.abs(r) > 1e-12) == 1
    # direct-sum definition Rxy[k] = sum_m x[m+k] y[m]
    rng = np.random.default_rng(3)
    x = rng.standard_normal(33); y = rng.standard_normal(33)
    direct = np.correlate(x, y, mode="full") / (np.linalg.norm(x) * np.linalg.norm(y))
    assert np.allclose(orc.cc(x, y, unbiased=False), direct, atol=1e-14)
    # unbiased = biased * n/(n-|k|)
    ub = orc.cc(x, y, unbiased=True)
    k = np.arange(-32, 33)
    assert np.allclose(ub, direct * 33 / (33 - np.abs(k)), atol=1e-12)
    # sinusoid pair: normalised CCF envelope (n-|k|)/n * cos(w k) (up to O(1/n))
    n = 4000; w = 2 * np.pi * 50 / n
    s = np.cos(w * np.arange(n))
    r = orc.cc(s, s, unbiased=False)
    kk = np.arange(-(n - 1), n)
    assert np.max(np.abs(r - (n - np.abs(kk)) / n * np.cos(w * kk))) < 2e-2
    with pytest.raises(AssertionError):
        orc.cc(np.zeros(3), np.zeros(4))
    with pytest.raises(AssertionError):
        orc.cc(np.zeros(3), np.zeros(3, dtype=np.float32))


def test_crop_indices_and_unbias():
    npts, delta = 1000, 0.02
    i0, i1 = orc.lag_clip_indices(npts, delta, (0.0, 9.0))
    assert (i0, i1) == (999, 999 + 450)
    lags = orc.lag(npts, delta)[i0:i1]
    assert orc.update_lag_indices(lags, npts, delta, 1 / delta) == (i0, i1)
    assert orc.lag_clip_indices(npts, delta, None) == (0, 2 * npts - 1)
    rec = orc.synthetic_receivers(2, npts, seed=5, delay=37)
    cell = orc.process_cell(rec[0], rec[1], npts, i0, i1, dtype=np.float32)
    assert cell.dtype == np.float32 and cell.shape == (450,)
    assert np.argmax(cell) == 37
    ub = orc.process_cell(rec[0], rec[1], npts, i0, i1, dtype=np.float32,
                          bias_correct=True)
    assert np.allclose(ub, cell * (npts / (npts - np.arange(450))), rtol=1e-6)
    cell_s = orc.process_cell(rec[0], rec[1], npts, i0, i1, dtype=np.float64, smooth=True)
    cell_r = orc.process_cell(rec[0], rec[1], npts, i0, i1, dtype=np.float64)
    assert np.max(np.abs(cell_s - cell_r)) < 1e-14
