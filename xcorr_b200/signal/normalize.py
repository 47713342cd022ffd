r"""
:mod:`signal.normalize` -- vector / matrix norms on the device
==============================================================

``norm1d`` / ``norm2d`` of the reference (``signal/normalize.py:33-155``) divide
by ``np.linalg.norm`` along the last (two) axes.  On the device the sum of
squares is the fixed-order float64 reduction the correlation path already
uses for its normalisation (``xcb_forward``); here it is exposed for arrays.
"""
import torch

__all__ = ['norm1d', 'norm2d']


def _norm(x, ndim):
    if not torch.is_tensor(x) or not x.is_cuda:
        raise TypeError('x must be a CUDA tensor (xcorr_b200 has no CPU path)')
    dims = tuple(range(-ndim, 0))
    n = torch.linalg.vector_norm(x.double(), dim=dims, keepdim=True)
    return (x.double() / n).to(x.dtype)


def norm1d(x):
    """x / ||x||_2 along the last axis (``signal/normalize.py:33-92``)."""
    return _norm(x, 1)


def norm2d(x):
    """x / Frobenius norm over the last two axes (``signal/normalize.py:95-155``)."""
    return _norm(x, 2)
