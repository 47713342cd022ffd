r"""
:mod:`signal.postprocess_ops` -- CCF post-processing operators on the device
=============================================================================

Array-valued counterparts of the reference's ``xcorr.signal`` operators that
``core/postprocess.py:133-141`` chains on every CCF.  They act on
CUDA tensors ``[..., lag]`` along the last axis; the lag coordinate's
``sampling_rate`` attribute the reference reads from ``x[dim].attrs`` is an
explicit argument here.

============  =====================================================================
``unbias``    ``signal/unbias.py:30-103``: ``x * w`` with ``w = weight(npts)[i0:i1]``
``demean``    ``signal/detrend.py:113`` (``x - x.mean(dim)``)
``taper``     ``signal/taper.py:25-85`` / ``signal/window.py:27-116``
              (``scipy.signal.windows.<wtype>(2 nmax, False)`` halves at both ends)
``timeshift`` ``signal/timeshift.py:24-117`` (rfft, phase ramp, irfft; ``fast=True``)
``filter``    ``signal/filter.py:27-120`` (``butter`` + ``sosfiltfilt``, odd padding,
              ``sosfilt_zi`` initial states)
``snr``       ``signal/snr.py:30-168`` (max |x| in the signal window over the rms of
              the noise window)
============  =====================================================================

Host side: filter / window *design* with the same scipy calls the reference makes.
Device side (``libxcorr_b200.so``): ``xcb_stats`` / ``xcb_pointwise_apply`` (means,
windows, weights), ``xcb_dft_bluestein`` + ``xcb_phase_ramp`` (time shift of exactly
``n`` points), ``xcb_sosfilt`` forward and backward with initial states, and
``xcb_window_stats``.
"""
import ctypes

import numpy as np
import scipy.signal
import torch

from .. import _lib
from ..stream.process import _carry_matrices, _carry_window, _SCAN_CHUNK
from ..util.cc import weight
from .fft import _plan

__all__ = ['unbias', 'unbias_weights', 'demean', 'detrend', 'rms', 'taper', 'window', 'timeshift', 'filter', 'snr']

_CODE = {torch.float32: _lib.XCB_F32, torch.float64: _lib.XCB_F64}


def _rows(x):
    if not torch.is_tensor(x) or not x.is_cuda:
        raise TypeError('x must be a CUDA tensor (xcorr_b200 has no CPU path)')
    if x.dtype not in _CODE:
        raise TypeError('x must be float32 or float64')
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    return x2, lead


def _stream(x):
    return torch.cuda.current_stream(x.device).cuda_stream


def _pointwise(x2, detrend=0, left=None, right=None, out_dtype=None):
    """y = w(i) * (x - mean) with w given by its first / last values."""
    lib = _lib.load()
    nsig, n = x2.shape
    y = torch.empty((nsig, n), dtype=out_dtype or x2.dtype, device=x2.device)
    pw = _lib.xcb_pointwise()
    pw.detrend = detrend
    keep = []
    if left is not None and len(left):
        t = torch.from_numpy(np.ascontiguousarray(left, dtype=np.float64)).to(x2.device)
        keep.append(t)
        pw.taper_left, pw.ramp_left = len(left), t.data_ptr()
    if right is not None and len(right):
        t = torch.from_numpy(np.ascontiguousarray(right, dtype=np.float64)).to(x2.device)
        keep.append(t)
        pw.taper_right, pw.ramp_right = len(right), t.data_ptr()
    stats_ptr = None
    if detrend:
        nb = ctypes.c_size_t()
        _lib.check(lib.xcb_stats_workspace_bytes(n, nsig, ctypes.byref(nb)), 'xcb_stats_workspace_bytes')
        ws = torch.empty(nb.value, dtype=torch.uint8, device=x2.device)
        stats = torch.empty((nsig, 4), dtype=torch.float64, device=x2.device)
        keep += [ws, stats]
        stats_ptr = stats.data_ptr()
        _lib.check(lib.xcb_stats(_CODE[x2.dtype], x2.data_ptr(), n, x2.stride(0), nsig, stats_ptr, ws.data_ptr(),
                                 _stream(x2)), 'xcb_stats')
    _lib.check(lib.xcb_pointwise_apply(_CODE[x2.dtype], x2.data_ptr(), x2.stride(0), _CODE[y.dtype], y.data_ptr(),
                                       y.stride(0), n, nsig, ctypes.byref(pw), stats_ptr, _stream(x2)),
               'xcb_pointwise_apply')
    return y


def unbias_weights(npts, index_min, index_max):
    """``signal/unbias.py:106-145``: ``weight(npts, pad=True)[index_min:index_max]``."""
    return weight(npts, pad=True)[index_min:index_max]


def unbias(x, w):
    """``x * w.astype(x.dtype)`` along the last axis (``signal/unbias.py:85``)."""
    x2, lead = _rows(x)
    w = np.asarray(w)
    if w.shape != (x2.shape[1],):
        raise ValueError('w must have one value per lag')
    npdt = np.float32 if x.dtype == torch.float32 else np.float64
    return _pointwise(x2, left=w.astype(npdt).astype(np.float64)).reshape(*lead, -1)


def demean(x):
    """``x - x.mean(dim)`` (``signal/detrend.py:113``)."""
    x2, lead = _rows(x)
    return _pointwise(x2, detrend=1).reshape(*lead, -1)


def detrend(x, type=None):
    """``signal/detrend.py:30-111``: ``type='constant'`` (default) removes the mean, ``'linear'`` the
    least-squares line over the (regularly spaced) last axis -- what ``scipy.stats.linregress`` fits there."""
    type = type or 'constant'
    if type not in ('constant', 'linear'):
        raise KeyError('type should be any of "constant|linear"')
    x2, lead = _rows(x)
    return _pointwise(x2, detrend=1 if type == 'constant' else 2).reshape(*lead, -1)


def rms(x):
    """``signal/rms.py``: root-mean-square along the last axis (host array, float64)."""
    x2, lead = _rows(x)
    out = torch.empty((x2.shape[0], 3), dtype=torch.float64, device=x2.device)
    n = x2.shape[1]
    _lib.check(_lib.load().xcb_window_stats(_CODE[x2.dtype], x2.data_ptr(), x2.stride(0), x2.shape[0], 0, 1, 0, n,
                                            out.data_ptr(), _stream(x2)), 'xcb_window_stats')
    return np.sqrt(out.cpu().numpy()[:, 2] / float(n)).reshape(lead)


def window(n, sampling_rate, wtype=None, max_percentage=None, max_length=None, side=None):
    """Host: the taper window of ``signal/window.py:27-116`` as a float64 array."""
    wtype = wtype or 'hann'
    if wtype not in scipy.signal.windows.__all__:
        raise ValueError('wtype should be a :func:`scipy.signal.windows`!')
    side = side or 'both'
    if side not in ('left', 'right', 'both'):
        raise ValueError('side should be either "left", "right" or "both"!')
    nmax = 0.5 * n
    if max_percentage:
        if not isinstance(max_percentage, float):
            raise TypeError('max_percentage should be of type `float`.')
        if max_percentage < 0. or max_percentage > .5:
            raise ValueError('max_percentage should be within [0, .5].')
        nmax = max_percentage * n
    if max_length:
        if not isinstance(max_length, float):
            raise TypeError('max_length should be of type `float`.')
        nmax = max_length * sampling_rate
    nmax = int(round(nmax))
    win = getattr(scipy.signal.windows, wtype)(nmax * 2, False)
    w = np.ones(n, dtype=np.float64)
    # (the reference applies both halves whatever `side` says, signal/window.py:94-95)
    w[0:nmax] = win[0:nmax]
    w[-nmax:] = win[-nmax:]
    return w, nmax


def taper(x, sampling_rate, wtype=None, max_percentage=None, max_length=None, side=None):
    """``x * window(...).astype(x.dtype)`` (``signal/taper.py:62-66``)."""
    x2, lead = _rows(x)
    w, nmax = window(x2.shape[1], sampling_rate, wtype, max_percentage, max_length, side)
    npdt = np.float32 if x.dtype == torch.float32 else np.float64
    w = w.astype(npdt).astype(np.float64)
    nmax = min(nmax, x2.shape[1] // 2)
    return _pointwise(x2, left=w[:nmax], right=w[len(w) - nmax:]).reshape(*lead, -1)


def timeshift(x, delay, sampling_rate, fast=True):
    """Shift ``x[..., lag]`` by ``delay`` seconds in the frequency domain
    (``signal/timeshift.py:86-99``).  `delay`: float or array broadcastable to
    ``x.shape[:-1]``.  With an odd number of samples one sample is trimmed, like the
    reference's rfft (``signal/fft.py:357``)."""
    if not fast:
        raise NotImplementedError('only the rfft / irfft path (fast=True) runs on the device')
    x2, lead = _rows(x)
    lib = _lib.load()
    size = x2.shape[1]
    n = size - 1 if size & 1 else size
    delta = 1.0 / sampling_rate
    fN = np.round(n // 2 / ((n - 1) * delta), decimals=0)          # signal/fft.py:358
    df = fN / (n // 2)
    nsig = x2.shape[0]
    plan = _plan(n, x2.dtype, x2.device, nsig)
    X = plan.rfft(x2[:, :n])
    d = np.broadcast_to(np.asarray(delay, dtype=np.float64), lead if lead else ()).reshape(-1)
    dd = torch.from_numpy(np.array(d, dtype=np.float64)).to(x2.device)
    _lib.check(lib.xcb_phase_ramp(_CODE[x2.dtype], X.data_ptr(), X.stride(0), X.shape[1], nsig, dd.data_ptr(),
                                  float(df), _stream(x2)), 'xcb_phase_ramp')
    return plan.irfft(X).reshape(*lead, -1)


def _sosfilt(x2, sos, reverse, init):
    """One causal pass of scipy.signal.sosfilt(sos, x, zi=init) (backwards if `reverse`), float64 out."""
    lib = _lib.load()
    nsig, n = x2.shape
    f = _lib.xcb_sos()
    f.n_sections = len(sos)
    f.decimate = 1
    f.reverse = int(reverse)
    f.chunk = _SCAN_CHUNK
    for k, sec in enumerate(sos):
        a0 = sec[3]
        for m, v in enumerate((sec[0] / a0, sec[1] / a0, sec[2] / a0, sec[4] / a0, sec[5] / a0)):
            f.coef[k][m] = float(v)
    P, PG = _carry_matrices(sos, _SCAN_CHUNK)
    f.carry_window = _carry_window(P)
    dP, dPG = torch.from_numpy(P).to(x2.device), torch.from_numpy(PG).to(x2.device)
    f.carry = ctypes.cast(dP.data_ptr(), ctypes.POINTER(ctypes.c_double))
    f.carry_group = ctypes.cast(dPG.data_ptr(), ctypes.POINTER(ctypes.c_double))
    f.init_state = ctypes.cast(init.data_ptr(), ctypes.POINTER(ctypes.c_double))
    nb = ctypes.c_size_t()
    _lib.check(lib.xcb_sosfilt_workspace_bytes(n, nsig, ctypes.byref(f), ctypes.byref(nb)), 'xcb_sosfilt_workspace_bytes')
    ws = torch.empty(nb.value, dtype=torch.uint8, device=x2.device)
    y = torch.empty((nsig, n), dtype=torch.float64, device=x2.device)
    _lib.check(lib.xcb_sosfilt(_CODE[x2.dtype], x2.data_ptr(), x2.stride(0), _lib.XCB_F64, y.data_ptr(), y.stride(0),
                               n, nsig, ctypes.byref(f), None, None, ws.data_ptr(), _stream(x2)), 'xcb_sosfilt')
    return y


def filter(x, sampling_rate, frequency, btype='highpass', order=2):
    """Zero-phase Butterworth filter: ``scipy.signal.butter(output='sos', fs=...)`` +
    ``scipy.signal.sosfiltfilt`` (``signal/filter.py:86-96``): odd extension by
    ``3 * ntaps`` samples, forward and backward ``sosfilt`` from the steady-state
    initial conditions ``sosfilt_zi(sos) * first sample``."""
    if not (isinstance(frequency, float) or (isinstance(frequency, tuple) and len(frequency) == 2)):
        raise ValueError('Corner frequency should be a `float` or tuple-pair with (min, max)!')
    x2, lead = _rows(x)
    btype = btype or 'highpass'
    sos = scipy.signal.butter(N=order, Wn=frequency, btype=btype, output='sos', fs=sampling_rate)
    n_sections = sos.shape[0]
    ntaps = 2 * n_sections + 1
    ntaps -= min((sos[:, 2] == 0).sum(), (sos[:, 5] == 0).sum())
    edge = 3 * ntaps
    n = x2.shape[1]
    if n <= edge:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % edge)
    zi = torch.from_numpy(scipy.signal.sosfilt_zi(sos).reshape(-1)).to(x2.device)        # (2S,)
    xd = x2.double()
    # odd extension at both ends (scipy.signal._arraytools.odd_ext) -- buffer assembly only
    left = 2.0 * xd[:, :1] - xd[:, 1:edge + 1].flip(1)
    right = 2.0 * xd[:, -1:] - xd[:, n - edge - 1:n - 1].flip(1)
    ext = torch.cat((left, xd, right), dim=1).contiguous()
    y = _sosfilt(ext, sos, False, (ext[:, :1] * zi[None, :]).contiguous())
    y = _sosfilt(y, sos, True, (y[:, -1:] * zi[None, :]).contiguous())
    return y[:, edge:edge + n].to(x.dtype).reshape(*lead, -1)


def snr(x, signal_window, noise_window, power=False, decibels=False, extend=False):
    """``signal/snr.py:107-121`` with index windows ``(begin, end)`` along the last
    axis instead of boolean masks: ``s = max |x|`` over the signal window, ``n`` = rms
    over the noise window, ``snr = s / n`` (squared for `power`, ``20 log10`` for
    `decibels`).  Returns ``snr`` or ``(snr, s, n, index of s)`` if `extend`."""
    x2, lead = _rows(x)
    lib = _lib.load()
    s0, s1 = (int(v) for v in signal_window)
    n0, n1 = (int(v) for v in noise_window)
    size = x2.shape[1]
    if not (0 <= s0 < s1 <= size and 0 <= n0 < n1 <= size):
        raise ValueError('windows must lie inside the lag axis')
    out = torch.empty((x2.shape[0], 3), dtype=torch.float64, device=x2.device)
    _lib.check(lib.xcb_window_stats(_CODE[x2.dtype], x2.data_ptr(), x2.stride(0), x2.shape[0], s0, s1, n0, n1,
                                    out.data_ptr(), _stream(x2)), 'xcb_window_stats')
    res = out.cpu().numpy()
    s = res[:, 0].reshape(lead)
    idx = res[:, 1].astype(np.int64).reshape(lead)
    noise = np.sqrt(res[:, 2] / float(n1 - n0)).reshape(lead)
    ratio = s / noise
    if decibels:
        ratio = 20 * np.log10(ratio)
    elif power:
        ratio = np.square(ratio)
    return (ratio, s, noise, idx) if extend else ratio
