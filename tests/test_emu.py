"""CPU checks of the CUDA pass bodies through the host thread-emulator
(tests/host_emu): the same xcb_passes.h / xcb_plan.h code the kernels compile,
run thread-by-thread, compared with the oracle."""
import numpy as np
import pytest

from oracle import cc_oracle as orc
from tests.emu_util import choose_dims, emu_cc
from tests.algo_model import forward_packed, half_spectrum, freq_of_address

MENU = [64, 80, 96, 100, 120, 128, 144, 150, 160, 180, 192, 200, 225, 240, 256, 300, 320, 400]


def _check(x, pairs, dims, tol, **kw):
    out, ss, _ = emu_cc(x, pairs, dims=dims, **kw)
    n = x.shape[1]
    i0 = kw.get("i0") or 0
    i1 = kw.get("i1") or 2 * n - 1
    for q, (i, j) in enumerate(pairs):
        ref = orc.cc(x[i].astype(np.float64), x[j].astype(np.float64),
                     normalize=kw.get("normalize", True), unbiased=kw.get("unbiased", False))
        peak = np.max(np.abs(ref))
        ref = ref[i0:i1]
        assert np.max(np.abs(out[q] - ref)) <= tol * peak, (dims, q)
        if not kw.get("unbiased", False) and i0 <= np.argmax(np.abs(orc.cc(x[i].astype(np.float64), x[j].astype(np.float64), unbiased=False))) < i1:
            assert np.argmax(np.abs(out[q])) == np.argmax(np.abs(ref))
    # float32 mode squares in float32 (exact promotion costs a conversion per sample)
    assert np.allclose(ss, np.sum(x.astype(np.float64) ** 2, axis=1), rtol=1e-13 if x.dtype == np.float64 else 1e-6)


@pytest.mark.parametrize("N", MENU)
def test_every_menu_size_single_pass_f64(N):
    rng = np.random.default_rng(N)
    n = N - 3
    x = orc.synthetic_receivers(3, n, seed=N, delay=5)
    _check(x, [(0, 1), (2, 0)], (N,), 1e-13)


@pytest.mark.parametrize("N", MENU)
def test_every_menu_size_as_column_pass_f64(N):
    # N leads (column pass, real input / lag epilogue); inner = 100 is not a
    # multiple of the 8-column tile -> exercises partial tiles too
    n = N * 100 - 7
    x = orc.synthetic_receivers(2, n, seed=N + 1, delay=11)
    _check(x, [(0, 1)], (N, 100), 1e-13)


@pytest.mark.parametrize("dims", [(64, 64, 64), (100, 96, 80), (64, 80, 64, 64)])
def test_multi_pass_f64(dims):
    M = int(np.prod(dims))
    n = min(M, 300_000) - 1
    if len(dims) == 4:
        n = 70_000          # heavily padded: most of the transform is zeros
    x = orc.synthetic_receivers(2, n, seed=7, delay=137)
    _check(x, [(0, 1), (1, 0)], dims, 1e-13)


@pytest.mark.parametrize("N", [64, 150, 240, 300, 400])
def test_f32_single_and_two_pass(N):
    x = orc.synthetic_receivers(3, N * 64 - 1, seed=N, delay=29).astype(np.float32)
    _check(x, [(0, 1), (1, 2)], (N, 64), 1e-5)
    x = orc.synthetic_receivers(2, N - 1, seed=N, delay=3).astype(np.float32)
    _check(x, [(0, 1)], (N,), 1e-5)


def test_crop_unbiased_nonorm_and_odd_lengths():
    for n in (1, 2, 3, 63, 64, 1000, 4095, 4096):
        x = orc.synthetic_receivers(2, n, seed=n, delay=1)
        _check(x, [(0, 1)], None, 1e-13)
    n = 5000
    x = orc.synthetic_receivers(2, n, seed=1, delay=137)
    i0, i1 = orc.lag_clip_indices(n, 0.02, (0.0, 30.0))
    _check(x, [(0, 1)], None, 1e-13, i0=i0, i1=i1)
    _check(x, [(0, 1)], None, 1e-13, i0=100, i1=4000)
    _check(x, [(0, 1)], None, 1e-13, normalize=False)
    _check(x, [(0, 1)], None, 1e-12, unbiased=True, i0=i0, i1=i1)
    out, _, _ = emu_cc(x, [(0, 1)], i0=i0, i1=i1)
    assert np.argmax(out[0]) == 137


def test_zero_vector_gives_nan_like_reference():
    x = np.zeros((2, 50))
    x[1] = np.random.default_rng(0).standard_normal(50)
    out, _, _ = emu_cc(x, [(0, 1)])
    with np.errstate(all="ignore"):
        ref = orc.cc(x[0], x[1], unbiased=False)
    assert np.all(np.isnan(out)) and np.all(np.isnan(ref))


@pytest.mark.parametrize("dims", [(64, 80), (225,), (100, 64, 96)])
def test_spectrum_is_the_half_spectrum_in_digit_transposed_order(dims):
    M = int(np.prod(dims))
    x = np.random.default_rng(2).standard_normal((1, min(5000, 2 * M - 3)))
    _, _, spec = emu_cc(x[:, :min(x.shape[1], M)], [(0, 0)], dims=dims, want_spec=True)
    x = x[:, :min(x.shape[1], M)]
    X = half_spectrum(forward_packed(x[0], dims), dims)
    assert np.max(np.abs(spec[0] - X)) < 1e-12 * np.max(np.abs(X))
    # ... which is numpy's rfft of the zero-padded signal, frequency k at its digit-transposed address
    R = np.fft.rfft(x[0], 2 * M)
    k = freq_of_address(dims)
    ref = R[k].copy()
    ref[0] = R[0].real + 1j * R[M].real
    assert np.max(np.abs(spec[0] - ref)) < 1e-12 * np.max(np.abs(ref))


def test_planner_named_shapes():
    assert int(np.prod(choose_dims(21_600_000))) == 21_600_000
    # 24 h @ 50 Hz: a 256-point row axis is preferred over the exact 180*150*160 (2.4 % more points)
    d = choose_dims(4_320_000)
    assert 4_320_000 <= int(np.prod(d)) <= 1.03 * 4_320_000 and d[-1] == 256
    d = choose_dims(900_000)
    assert 900_000 <= int(np.prod(d)) <= 1.03 * 900_000
    assert choose_dims(5) == (64,)


def test_spectral_stacking_sums_windows():
    """n_stack windows per receiver: the cross-spectra are summed before one
    inverse transform == sum of the per-window CCFs (each normalised by its own norms)."""
    n, R, W = 3000, 3, 4
    rec = np.stack([orc.synthetic_receivers(R, n, seed=50 + w, delay=17) for w in range(W)], axis=1)   # [R, W, n]
    pairs = [(0, 1), (1, 2), (0, 2)]
    out, _, _ = emu_cc(rec.reshape(R * W, n), pairs, n_stack=W)
    for q, (i, j) in enumerate(pairs):
        ref = sum(orc.cc(rec[i, w], rec[j, w], unbiased=False) for w in range(W))
        assert np.max(np.abs(out[q] - ref)) <= 1e-13 * np.max(np.abs(ref))
        assert np.argmax(out[q]) == np.argmax(ref)


@pytest.mark.parametrize("B,dims,n,K,lag0", [
    (2, (64, 64), 6000, 1500, 0),          # F = 8192 >= 3000 + 1500
    (3, (64, 80), 12000, 5000, 0),         # F = 10240 >= 4000 + 5000
    (2, (64, 64, 64), 400_000, 100_000, 0),
    (2, (64, 64), 6000, 1500, 700),        # first kept lag > 0
    (4, (100,), 500, 60, 0),               # single-pass transform, F = 200 >= 125 + 60
    (2, (64, 80), 6000, 3001, -1500),      # symmetric window: segments start before the signal
    (3, (64, 64), 5001, 1000, -4000),      # negative lags only, odd length, odd offset
    (2, (80, 80), 6000, 2500, -5999),      # window reaching the most negative lag
])
def test_block_mode_matches_full_transform(B, dims, n, K, lag0):
    """Cropped non-negative lags through B blocks of the second signal and
    transforms of length ~n/B + K instead of 2n: same lags, same values."""
    x = orc.synthetic_receivers(3, n, seed=n + B, delay=41)
    pairs = [(0, 1), (2, 1), (0, 2)]
    i0 = n - 1 + lag0
    lb = -(-n // B)
    out, ss, _ = emu_cc(x, pairs, dims=dims, i0=i0, i1=i0 + K, n_blocks=B, block_len=lb)
    assert emu_cc.last_blocks == B
    assert np.allclose(ss, np.sum(x ** 2, axis=1), rtol=1e-13)
    for q, (i, j) in enumerate(pairs):
        ref = orc.cc(x[i], x[j], unbiased=False)
        peak = np.max(np.abs(ref))
        assert np.max(np.abs(out[q] - ref[i0:i0 + K])) <= 1e-13 * peak, (q,)
    # unbiased weights use the full window length
    out_u, _, _ = emu_cc(x, pairs[:1], dims=dims, i0=i0, i1=i0 + K, n_blocks=B, block_len=lb, unbiased=True)
    ref_u = orc.cc(x[0], x[1], unbiased=True)
    assert np.max(np.abs(out_u[0] - ref_u[i0:i0 + K])) <= 1e-12 * np.max(np.abs(ref_u[i0:i0 + K]))


def test_block_planner_picks_blocks_for_cropped_day():
    # with many pairs per receiver the inverse transforms dominate and blocks pay off
    R = 6
    x = orc.synthetic_receivers(R, 40_000, seed=3, delay=41).astype(np.float32)
    pairs = [(i, j) for i in range(R) for j in range(i + 1, R)]
    out, _, _ = emu_cc(x, pairs, dims=None, i0=39_999, i1=39_999 + 15_000, n_blocks=-1)
    assert emu_cc.last_blocks >= 2
    for q in (0, 7, 14):
        i, j = pairs[q]
        ref = orc.cc(x[i].astype(np.float64), x[j].astype(np.float64), unbiased=False)
        assert np.max(np.abs(out[q] - ref[39_999:39_999 + 15_000])) <= 1e-5 * np.max(np.abs(ref))
        assert np.argmax(out[q]) == 41 * (j - i)
    # a single pair of two receivers: the classic transform stays cheaper
    emu_cc(x[:2], [(0, 1)], dims=None, i0=39_999, i1=39_999 + 15_000, n_blocks=-1)
    assert emu_cc.last_blocks == 1


# ---- warp-synchronous row pass (row axis 16 x R2: 128, 160, 192, 240, 256) ----
@pytest.mark.parametrize("dims", [(64, 128), (80, 240), (100, 160), (96, 192), (64, 64, 256), (150, 256)])
def test_warp_row_pass_as_last_axis_f64(dims):
    M = int(np.prod(dims))
    n = min(M, 200_000) - 5
    x = orc.synthetic_receivers(3, n, seed=M % 1000, delay=137)
    _check(x, [(0, 1), (2, 1)], dims, 1e-13)
    i0 = n - 1 - 1000
    _check(x, [(0, 2)], dims, 1e-12, i0=i0, i1=i0 + 4001, unbiased=True)


@pytest.mark.parametrize("dims", [(64, 256), (80, 240), (64, 64, 128)])
def test_warp_row_pass_f32(dims):
    M = int(np.prod(dims))
    x = orc.synthetic_receivers(3, min(M, 150_000) - 1, seed=11, delay=29).astype(np.float32)
    _check(x, [(0, 1), (1, 2)], dims, 1e-5)


def test_warp_row_pass_blocks_and_stacking():
    n, K = 30_000, 9_000
    x = orc.synthetic_receivers(3, n, seed=77, delay=41)
    pairs = [(0, 1), (2, 1), (0, 2)]
    for lag0 in (0, -4500):
        i0 = n - 1 + lag0
        out, ss, _ = emu_cc(x, pairs, dims=(64, 256), i0=i0, i1=i0 + K, n_blocks=2, block_len=15_008)
        assert emu_cc.last_blocks == 2
        for q, (i, j) in enumerate(pairs):
            ref = orc.cc(x[i], x[j], unbiased=False)
            assert np.max(np.abs(out[q] - ref[i0:i0 + K])) <= 1e-13 * np.max(np.abs(ref)), (lag0, q)
    R, W, nw = 3, 3, 9000
    rec = np.stack([orc.synthetic_receivers(R, nw, seed=90 + w, delay=17) for w in range(W)], axis=1)
    out, _, _ = emu_cc(rec.reshape(R * W, nw), pairs, dims=(64, 160), n_stack=W)
    for q, (i, j) in enumerate(pairs):
        ref = sum(orc.cc(rec[i, w], rec[j, w], unbiased=False) for w in range(W))
        assert np.max(np.abs(out[q] - ref)) <= 1e-13 * np.max(np.abs(ref))


@pytest.mark.parametrize("max_jobs,max_rows", [(8, 9), (4, 4)])
def test_pair_groups_cover_every_job_once(max_jobs, max_rows):
    """Planner grouping for the staged pair product: every job in exactly one group,
    group sizes within the kernel's limits, and spectra shared within groups."""
    import ctypes
    from tests.emu_util import lib
    rng = np.random.default_rng(5)
    cases = [[(i, j) for i in range(16) for j in range(i + 1, 16)],              # all pairs of 16 receivers
             [(0, 1)], [(2, 2), (0, 1), (1, 0)],
             [tuple(p) for p in rng.integers(0, 9, size=(40, 2))]]               # arbitrary list with repeats
    for pairs in cases:
        arr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32))
        grp = np.zeros(len(pairs), dtype=np.int32)
        rows = ctypes.c_int()
        n = lib().emu_pair_groups(arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(pairs), max_jobs, max_rows,
                                  grp.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctypes.byref(rows))
        assert n >= 1, n
        assert np.all(grp >= 0)
        assert np.bincount(grp).max() <= max_jobs
        if len(pairs) == 120:
            # sharing: clearly fewer staged rows than the 2 per job of the ungrouped kernel
            assert rows.value <= (0.9 if max_jobs == 8 else 1.2) * len(pairs), rows.value
