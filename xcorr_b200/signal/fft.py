r"""
:mod:`signal.fft` -- DFTs of any length on the device
=====================================================

Device counterpart of the reference's ``signal/fft.py`` (``fft`` / ``ifft`` /
``rfft`` / ``irfft`` wrappers around numpy / pyfftw, ``signal/fft.py:15-18,
109-112``).  The reference transforms exactly ``npts`` points whatever their
factorisation; here a length that is not a product of the compiled pass sizes
goes through Bluestein's chirp-z algorithm on a smooth plan of at least
``2 npts - 1`` points (``xcb_dft_bluestein``): three of the hand-written
multi-pass transforms (one cached per length) and two chirp multiplies, all
in ``libxcorr_b200.so``.  torch only owns the buffers.

Arrays are CUDA tensors ``[..., n]`` transformed along the last axis; results
match ``numpy.fft`` to the working precision (float64 ~1e-14, float32 ~1e-6 of
the largest coefficient).
"""
import ctypes

import torch

from .. import _lib

__all__ = ['DFT', 'fft', 'ifft', 'rfft', 'irfft']

_REAL = {torch.float32: _lib.XCB_F32, torch.float64: _lib.XCB_F64,
         torch.complex64: _lib.XCB_F32, torch.complex128: _lib.XCB_F64}
_CPLX = {_lib.XCB_F32: torch.complex64, _lib.XCB_F64: torch.complex128}
_RDT = {_lib.XCB_F32: torch.float32, _lib.XCB_F64: torch.float64}


def good_length(n_min):
    """Smallest transform length >= n_min the pass kernels support directly."""
    out = ctypes.c_int64()
    _lib.check(_lib.load().xcb_fft_good_length(int(n_min), ctypes.byref(out)), "xcb_fft_good_length")
    return out.value


class DFT:
    """DFT of exactly ``K`` points for batches of up to ``max_batch`` signals.

    ``forward`` / ``backward`` follow numpy's sign and scaling conventions
    (``fft``: exp(-2 pi i nk/K), unscaled; ``ifft``: exp(+...)/K).
    """

    def __init__(self, K, dtype=torch.float64, max_batch=16, device=None):
        if dtype not in (torch.float32, torch.float64):
            raise TypeError("dtype must be float32 or float64")
        self.K = int(K)
        if self.K < 1:
            raise ValueError("K must be positive")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise TypeError("xcorr_b200 has no CPU path")
        self.code = _REAL[dtype]
        self.M = good_length(max(2 * self.K - 1, 1))
        self.max_batch = int(max_batch)
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.xcb_fft_create(ctypes.byref(h), self.code, self.M, self.max_batch), "xcb_fft_create")
        self._h = h
        cd = _CPLX[self.code]
        self._filter = {s: torch.empty(self.M, dtype=cd, device=self.device) for s in (-1, 1)}
        self._ready = {-1: False, 1: False}
        self._work = torch.empty(self.max_batch * (self.K + self.M), dtype=cd, device=self.device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.xcb_fft_destroy(h)
            self._h = None

    def _run(self, x, mode, n_in, sign, n_out, real_out, scale):
        """x: [B, stride] contiguous; returns [B, n_out]."""
        B = x.shape[0]
        out = torch.empty((B, n_out), dtype=_RDT[self.code] if real_out else _CPLX[self.code], device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for b0 in range(0, B, self.max_batch):
            nb = min(self.max_batch, B - b0)
            xb, ob = x[b0:b0 + nb], out[b0:b0 + nb]
            _lib.check(self._lib.xcb_dft_bluestein(
                self._h, self.K, sign, xb.data_ptr(), mode, xb.stride(0), n_in, nb, ob.data_ptr(), int(real_out),
                ob.stride(0), n_out, float(scale), self._filter[sign].data_ptr(), int(self._ready[sign]),
                self._work.data_ptr(), stream), "xcb_dft_bluestein")
            self._ready[sign] = True
        return out

    def _prep(self, x, complex_in):
        if not torch.is_tensor(x) or not x.is_cuda:
            raise TypeError("x must be a CUDA tensor (xcorr_b200 has no CPU path)")
        want = _CPLX[self.code] if complex_in else _RDT[self.code]
        lead = x.shape[:-1]
        x2 = x.to(want).reshape(-1, x.shape[-1]).contiguous()
        return x2, lead

    def rfft(self, x):
        """numpy.fft.rfft(x, n=K) along the last axis (x real, at most K samples)."""
        x2, lead = self._prep(x, False)
        n_in = min(x2.shape[1], self.K)
        y = self._run(x2, 0, n_in, -1, self.K // 2 + 1, False, 1.0)
        return y.reshape(*lead, -1)

    def irfft(self, X, scale=1.0):
        """numpy.fft.irfft(X, n=K) * scale: X holds (at least) K//2 + 1 coefficients."""
        X2, lead = self._prep(X, True)
        if X2.shape[1] < self.K // 2 + 1:
            X2 = torch.nn.functional.pad(X2, (0, self.K // 2 + 1 - X2.shape[1]))
        y = self._run(X2, 2, self.K, +1, self.K, True, float(scale) / self.K)
        return y.reshape(*lead, -1)

    def fft(self, x):
        """numpy.fft.fft(x, n=K) (x real or complex)."""
        cplx = torch.is_complex(x)
        x2, lead = self._prep(x, cplx)
        y = self._run(x2, 1 if cplx else 0, min(x2.shape[1], self.K), -1, self.K, False, 1.0)
        return y.reshape(*lead, -1)

    def ifft(self, X):
        """numpy.fft.ifft(X, n=K)."""
        X2, lead = self._prep(X, True)
        y = self._run(X2, 1, min(X2.shape[1], self.K), +1, self.K, False, 1.0 / self.K)
        return y.reshape(*lead, -1)


_cache = {}


def _plan(K, dtype, device, batch):
    real = torch.float32 if dtype in (torch.float32, torch.complex64) else torch.float64
    key = (int(K), real, str(device))
    p = _cache.get(key)
    # signals per launch: the scratch is max_batch * (K + M) complex numbers -- keep it under 2 GiB
    per_signal = (int(K) + good_length(max(2 * int(K) - 1, 1))) * (8 if real == torch.float32 else 16)
    want = max(1, min(batch, 16, (2 << 30) // per_signal))
    if p is None or p.max_batch < want:
        if len(_cache) >= 4:                      # plans hold O(M) device memory
            _cache.pop(next(iter(_cache)))
        p = DFT(K, real, max_batch=want, device=device)
        _cache[key] = p
    return p


def _batch(x):
    n = 1
    for d in x.shape[:-1]:
        n *= d
    return n


def rfft(x, n=None):
    """``signal/fft.py`` rfft: real input, last axis, ``n`` points (default: all)."""
    return _plan(n or x.shape[-1], x.dtype, x.device, _batch(x)).rfft(x)


def irfft(X, n=None):
    """``signal/fft.py`` irfft; ``n`` defaults to ``2 (m - 1)`` like numpy."""
    return _plan(n or 2 * (X.shape[-1] - 1), X.dtype, X.device, _batch(X)).irfft(X)


def fft(x, n=None):
    return _plan(n or x.shape[-1], x.dtype, x.device, _batch(x)).fft(x)


def ifft(X, n=None):
    return _plan(n or X.shape[-1], X.dtype, X.device, _batch(X)).ifft(X)
