r"""
:mod:`signal.normalize` -- vector / matrix norms on the device
==============================================================

``norm1d`` / ``norm2d`` of the reference (``signal/normalize.py:33-155``) divide
by ``np.linalg.norm`` along the last (two) axes.  Here the sum of squares of each
row is the fixed-order float64 block reduction of ``xcb_window_stats`` (bit
reproducible) and the division runs in ``xcb_l2_normalize_rows``; both through
the C ABI.
"""
import torch

from .. import _lib

__all__ = ['norm1d', 'norm2d']

_XCB = {torch.float32: _lib.XCB_F32, torch.float64: _lib.XCB_F64}


def _norm(x, ndim):
    if not torch.is_tensor(x) or not x.is_cuda:
        raise TypeError('x must be a CUDA tensor (xcorr_b200 has no CPU path)')
    if x.dtype not in _XCB:
        x = x.double()
    lib = _lib.load()
    n = 1
    for d in x.shape[-ndim:]:
        n *= int(d)
    y = x.contiguous().clone().reshape(-1, n)           # rows = the reduced axes flattened
    rows = y.shape[0]
    stats = torch.empty((rows, 3), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.xcb_window_stats(_XCB[y.dtype], y.data_ptr(), n, rows, 0, 0, 0, n, stats.data_ptr(), stream),
                   'xcb_window_stats')
        _lib.check(lib.xcb_l2_normalize_rows(_XCB[y.dtype], y.data_ptr(), n, n, rows, stats.data_ptr(), stream),
                   'xcb_l2_normalize_rows')
    return y.reshape(x.shape)


def norm1d(x):
    """x / ||x||_2 along the last axis (``signal/normalize.py:33-92``)."""
    return _norm(x, 1)


def norm2d(x):
    """x / Frobenius norm over the last two axes (``signal/normalize.py:95-155``)."""
    return _norm(x, 2)
