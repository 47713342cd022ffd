r"""
:mod:`stream.resample` -- sampling-rate changes of the op-lists on the device
=============================================================================

``('resample', {'sampling_rate': r})`` and ``('interpolate', {'sampling_rate': r,
'method': 'lanczos', 'a': 20})`` of the reference's op-lists
(``stream/process.py:51-52``, examples ``examples/core/process.py:71-75``) are obspy
``Trace`` methods; here they run on CUDA tensors ``(n_traces, npts)``:

* :func:`resample_fourier` -- obspy ``Trace.resample``: rfft of exactly ``npts`` points
  (``signal.fft.DFT``: Bluestein on the hand-written multi-pass FFT, any length),
  Hann-windowed spectrum re-gridded by ``xcb_resample_spectrum``, inverse rfft of
  ``num`` points, scaled by ``num / npts``.
* :func:`interpolate_lanczos` -- obspy ``lanczos_interpolation`` (``xcb_lanczos``).
"""
import torch

from .. import _lib
from ..signal.fft import _plan

__all__ = ['resample_fourier', 'interpolate_lanczos']

_CODE = {torch.float32: _lib.XCB_F32, torch.float64: _lib.XCB_F64}
_CPLX = {torch.float32: torch.complex64, torch.float64: torch.complex128}


def resample_fourier(x, npts, fs, num, new_fs, window='hann'):
    """x: CUDA tensor (n_traces, npts) float32 / float64 -> (n_traces, num)."""
    if x.dim() != 2 or x.shape[1] != npts or not x.is_cuda:
        raise ValueError(f'x must be a CUDA tensor (n_traces, {npts})')
    lib = _lib.load()
    x = x.contiguous()
    nsig = x.shape[0]
    X = _plan(npts, x.dtype, x.device, nsig).rfft(x)                      # (nsig, npts//2 + 1)
    Y = torch.empty((nsig, num // 2 + 1), dtype=_CPLX[x.dtype], device=x.device)
    delta = 1.0 / fs
    df = 1.0 / (npts * delta)
    dlf = 1.0 / num * new_fs
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(lib.xcb_resample_spectrum(_CODE[x.dtype], X.data_ptr(), X.stride(0), Y.data_ptr(), Y.stride(0), npts,
                                         num, df, dlf, 0 if window is None else 1, nsig, stream),
               'xcb_resample_spectrum')
    return _plan(num, x.dtype, x.device, nsig).irfft(Y, scale=float(num) / float(npts))


def interpolate_lanczos(x, npts, fs, npts_new, new_fs, a, start_rel=0.0):
    """x: CUDA tensor (n_traces, npts) -> (n_traces, npts_new) sampled at start_rel + k / new_fs."""
    if x.dim() != 2 or x.shape[1] != npts or not x.is_cuda:
        raise ValueError(f'x must be a CUDA tensor (n_traces, {npts})')
    lib = _lib.load()
    if x.stride(1) != 1:
        x = x.contiguous()
    y = torch.empty((x.shape[0], npts_new), dtype=x.dtype, device=x.device)
    old_dt, new_dt = 1.0 / fs, 1.0 / new_fs
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(lib.xcb_lanczos(_CODE[x.dtype], x.data_ptr(), x.stride(0), npts, y.data_ptr(), y.stride(0), npts_new,
                               new_dt / old_dt, start_rel / old_dt, int(a), x.shape[0], stream), 'xcb_lanczos')
    return y
