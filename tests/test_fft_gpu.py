"""GPU parity of the DFT wrappers (xcb_fft_*, xcb_dft_bluestein) against numpy.fft --
which IS the reference implementation behind signal/fft.py:15-18 -- for lengths of
every kind: tiny, prime, smooth, odd, even, and a multi-million-point trace."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {np.float32: 2e-6, np.float64: 1e-13}


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import importlib
    xf = importlib.import_module("xcorr_b200.signal.fft")     # (the package also exports a function named fft)
    return torch, xf


def _rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("K", [1, 2, 3, 7, 64, 100, 101, 1000, 4099, 50_000, 65_537])
def test_rfft_irfft_any_length(env, K, dtype):
    torch, xf = env
    rng = np.random.default_rng(K)
    x = rng.standard_normal((3, K)).astype(dtype)
    xd = torch.from_numpy(x).cuda()
    X = xf.rfft(xd).cpu().numpy()
    ref = np.fft.rfft(x.astype(np.float64), axis=-1)
    assert X.shape == ref.shape
    assert _rel(X, ref) <= TOL[dtype] * max(1.0, np.log2(K + 1))
    y = xf.irfft(torch.from_numpy(ref.astype(np.complex64 if dtype == np.float32 else np.complex128)).cuda(), n=K)
    assert y.shape[-1] == K
    assert _rel(y.cpu().numpy(), np.fft.irfft(ref, n=K, axis=-1)) <= TOL[dtype] * max(1.0, np.log2(K + 1))


@pytest.mark.parametrize("K", [5, 360, 1021])
def test_complex_fft_ifft(env, K):
    torch, xf = env
    rng = np.random.default_rng(K)
    z = rng.standard_normal((2, K)) + 1j * rng.standard_normal((2, K))
    zd = torch.from_numpy(z).cuda()
    assert _rel(xf.fft(zd).cpu().numpy(), np.fft.fft(z, axis=-1)) <= 1e-13
    assert _rel(xf.ifft(zd).cpu().numpy(), np.fft.ifft(z, axis=-1)) <= 1e-13
    # n larger than the data: zero padding like numpy
    assert _rel(xf.rfft(zd.real.contiguous(), n=K + 11).cpu().numpy(), np.fft.rfft(z.real, n=K + 11, axis=-1)) <= 1e-13


def test_day_long_trace_of_awkward_length(env):
    """40 Hz x 24 h + 5 % buffer: 3 628 801 samples (odd, large prime factors)."""
    torch, xf = env
    K = 3_628_801
    rng = np.random.default_rng(1)
    x = rng.standard_normal(K)
    X = xf.rfft(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = np.fft.rfft(x)
    assert _rel(X, ref) <= 1e-12
    back = xf.irfft(torch.from_numpy(ref).cuda(), n=K).cpu().numpy()
    assert np.max(np.abs(back - x)) <= 1e-12 * np.max(np.abs(x))


def test_plain_plan_round_trip(env):
    """xcb_fft_forward / xcb_fft_inverse_conj on a smooth length: spectrum in digit-transposed
    order, so check Parseval and the round trip."""
    import ctypes
    torch, xf = env
    from xcorr_b200 import _lib
    lib = _lib.load()
    M = xf.good_length(30_000)
    h = ctypes.c_void_p()
    _lib.check(lib.xcb_fft_create(ctypes.byref(h), _lib.XCB_F64, M, 2), "create")
    rng = np.random.default_rng(3)
    z = rng.standard_normal((2, M)) + 1j * rng.standard_normal((2, M))
    zd = torch.from_numpy(z).cuda()
    spec = torch.empty_like(zd)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.xcb_fft_forward(h, zd.data_ptr(), M, M, 2, spec.data_ptr(), st), "forward")
    S = spec.cpu().numpy()
    ref = np.fft.fft(z, axis=-1)
    assert np.allclose(np.sort(np.abs(S), axis=-1), np.sort(np.abs(ref), axis=-1), rtol=1e-9, atol=1e-9)
    sc = torch.conj_physical(spec).contiguous()      # (conj() alone is a lazy view)
    out = torch.empty_like(zd)
    _lib.check(lib.xcb_fft_inverse_conj(h, sc.data_ptr(), 2, out.data_ptr(), M, M, st), "inverse")
    assert np.max(np.abs(out.cpu().numpy() - z)) <= 1e-12 * np.max(np.abs(z))
    lib.xcb_fft_destroy(h)
