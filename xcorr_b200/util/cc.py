r"""
:mod:`util.cc` -- Cross-correlation (B200)
==========================================

Drop-in for the reference's ``xcorr/util/cc.py``: same function names,
signatures and conventions.  ``cc`` runs on the GPU through the C ABI
(``xcb_cc_host``); ``lag``, ``weight`` and the shift helpers are the same
host-side numpy bookkeeping as in the reference.

Deviations from the reference, all where the reference cannot produce a result
(SURVEY appendix B):

* ``cc(..., unbiased=True)`` returns the intended estimate
  ``Rxy[k] * n / (n - |k|)`` (the reference raises: ``util/cc.py:79`` asks for
  weights of the wrong length and ``np.int`` no longer exists);
* ``weight`` returns all ``2n-1`` weights (the reference's ``arange`` stop at
  ``util/cc.py:138`` drops the last two); values agree wherever both exist.
"""
import ctypes

import numpy as np

from .. import _lib

__all__ = ['cc', 'lag', 'weight', 'extract_shift', 'extract_shift_and_max',
           'compute_shift', 'compute_shift_and_max']

_XCB_DTYPE = {np.dtype(np.float32): _lib.XCB_F32, np.dtype(np.float64): _lib.XCB_F64}


def _device_cc(x, y, normalize, unbiased, index_min=0, index_max=0):
    """Linear cross-correlation of two equal-length host vectors on the GPU."""
    n = len(x)
    k = (index_max - index_min) if (index_min or index_max) else 2 * n - 1
    out = np.empty(k, dtype=x.dtype)
    rc = _lib.load().xcb_cc_host(
        _XCB_DTYPE[x.dtype], n, x.ctypes.data_as(ctypes.c_void_p),
        y.ctypes.data_as(ctypes.c_void_p), int(bool(normalize)), int(bool(unbiased)),
        int(index_min), int(index_max), out.ctypes.data_as(ctypes.c_void_p))
    _lib.check(rc, "xcb_cc_host")
    return out


def cc(
    x: np.ndarray, y: np.ndarray, normalize: bool = True,
    pad: bool = True, unbiased: bool = True, dtype: np.dtype = None
):
    """
    Cross-correlate two vectors `x` and `y` in the frequency domain.

    Same contract as the reference ``util/cc.py:22-79``: ``Rxy[i]`` is the
    correlation at lag ``i - (n-1)`` samples (``pad=True``, length ``2n-1``),
    ``Rxy[k] = sum_m x[m+k] y[m]``, divided by ``||x|| ||y||`` when
    `normalize`.  The inputs are cast to `dtype` before the transform
    (``util/cc.py:68-71``); float32 and float64 are computed natively, any
    other dtype in float64.

    Parameters
    ----------
    x, y : :class:`numpy.ndarray`
        Data vectors of equal length and dtype.
    normalize : bool, optional
        Normalize by the norms of `x` and `y` (default).
    pad : bool, optional
        Linear correlation of length 2*N-1 if True (default); circular
        correlation of length N otherwise.
    unbiased : bool, optional
        Scale lag ``k`` by ``N / (N - |k|)`` (``pad=True`` only).
    dtype : :class:`numpy.dtype`, optional
        dtype of the estimate.  Defaults to the dtype of `x`.

    Returns
    -------
    Rxy : :class:`numpy.ndarray`
        Cross-correlation estimate of `x` and `y`.
    """
    n = len(x)
    x = np.asarray(x)
    y = np.asarray(y)
    dtype = np.dtype(dtype or x.dtype)
    assert n == len(y), "Vectors `x` and `y` should have the same length!"
    assert x.dtype == y.dtype, (
        "Vectors `x` and `y` should have the same dtype!"
    )
    if not pad:
        dtype = x.dtype                      # the reference ignores `dtype` here
        if unbiased:
            raise ValueError("unbiased=True needs pad=True")
    work = dtype if dtype in _XCB_DTYPE else np.dtype(np.float64)
    xs = np.ascontiguousarray(x, dtype=work)
    ys = np.ascontiguousarray(y, dtype=work)
    lin = _device_cc(xs, ys, normalize, unbiased and pad)
    if pad:
        return lin
    # circular correlation of length n from the linear one:
    # c[k] = r[k] + r[k-n]; then fftshift as util/cc.py:78 does
    circ = lin[n - 1:].copy()
    circ[1:] += lin[:n - 1]
    return np.fft.fftshift(circ)


def lag(
    n: int, delta: float, pad: bool = True
):
    """
    Construct the cross-correlation lag vector (``util/cc.py:82-108``).

    Returns ``(-(n-1) ... (n-1)) * delta`` seconds for ``pad=True``.
    """
    nn = n*2-1 if pad else n
    return np.fft.fftshift(np.fft.fftfreq(nn, 1/(nn*delta)))


def weight(
    n: int, pad: bool = True, clip: float = None
):
    """
    Construct the unbiased cross-correlation weight vector
    (``util/cc.py:111-141``, with the length fixed to ``nn``).
    """
    nn = n*2-1 if pad else n
    n = int((nn+1)/2)
    w = n / (n - np.abs(np.arange(1-n, nn-n+1, 1, np.float64)))
    if clip is not None:
        w[np.where(w > clip)] = clip
    return w


def extract_shift(
    Rxy: np.ndarray, delta: float = None
):
    """
    Sample (or time) shift at the maximum of `Rxy` (``util/cc.py:144-171``).
    """
    zero_index = int(len(Rxy) / 2)
    shift = np.argmax(Rxy) - zero_index
    return shift * (delta or 1)


def extract_shift_and_max(
    Rxy: np.ndarray, delta: float = None
):
    """
    Shift at the maximum of `Rxy` and the maximum itself
    (``util/cc.py:174-206``).
    """
    zero_index = int(len(Rxy) / 2)
    index_max = np.argmax(Rxy)
    shift = index_max - zero_index
    return shift * (delta or 1), Rxy[index_max]


def compute_shift(
    x: np.ndarray, y: np.ndarray, delta: float = None, **kwargs
):
    """
    Shift at the maximum of ``cc(x, y, **kwargs)`` (``util/cc.py:209-242``).
    """
    Rxy = cc(x, y, **kwargs)
    return extract_shift_and_max(Rxy, delta)[0]


def compute_shift_and_max(
    x: np.ndarray, y: np.ndarray, delta: float = None, **kwargs
):
    """
    Shift and maximum of ``cc(x, y, **kwargs)`` (``util/cc.py:245-280``).
    """
    Rxy = cc(x, y, **kwargs)
    return extract_shift_and_max(Rxy, delta)
