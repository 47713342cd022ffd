"""numpy / scipy restatement of the reference's CCF post-processing chain
(TEST INFRASTRUCTURE; not imported by the product package ``xcorr_b200``).

The reference implements these in ``xcorr/signal/*.py`` on xarray objects
(xarray is not installed here, so the modules cannot be imported); the
arithmetic inside is plain numpy / scipy, restated here function by function on
arrays ``[..., lag]``:

* ``window`` / ``taper``       -- ``signal/window.py:74-106``, ``signal/taper.py:62-66``
* ``demean``                   -- ``signal/detrend.py:113``
* ``timeshift``                -- ``signal/timeshift.py:86-99`` with the rfft / irfft grids of
                                  ``signal/fft.py:356-359, 476-479``
* ``filter``                   -- ``signal/filter.py:86-96`` (``butter`` + ``sosfiltfilt``)
* ``snr``                      -- ``signal/snr.py:107-121`` (``rms``: ``signal/rms.py``)
* ``postprocess``              -- ``core/postprocess.py:133-141``
* ``spectrogram``              -- ``signal/spectrogram.py:116-145``

Pinned by construction: every function is a direct call of the same numpy /
scipy routine the reference calls, with the same arguments.
"""
import numpy as np
import scipy.signal


def window(n, sampling_rate, wtype="hann", max_percentage=None, max_length=None):
    nmax = 0.5 * n
    if max_percentage:
        nmax = max_percentage * n
    if max_length:
        nmax = max_length * sampling_rate
    nmax = int(round(nmax))
    win = getattr(scipy.signal.windows, wtype)(nmax * 2, False)
    w = np.ones(n, dtype=np.float64)
    w[0:nmax] = win[0:nmax]
    w[-nmax:] = win[-nmax:]
    return w


def taper(x, sampling_rate, wtype="hann", max_percentage=None, max_length=None):
    w = window(x.shape[-1], sampling_rate, wtype, max_percentage, max_length).astype(x.dtype)
    return x * w


def demean(x):
    return x - x.mean(axis=-1, keepdims=True)


def timeshift(x, delay, lag0, delta, fast=True):
    """x[..., lag]; delay scalar or array broadcastable to x[..., 0]; lag0: first lag (s)."""
    size = x.shape[-1]
    if not fast:
        raise NotImplementedError
    n = size - 1 if size & 1 else size
    fN = np.round(n // 2 / ((n - 1) * delta), decimals=0)
    f = np.linspace(0.0, fN, n // 2 + 1)
    X = np.fft.rfft(x, n=n, axis=-1)
    X = X * np.exp(-1j * 2 * np.pi * np.asarray(delay)[..., None] * f)
    y = np.fft.irfft(X, axis=-1)
    return y.astype(x.dtype)


def filter(x, sampling_rate, frequency, btype="highpass", order=2):
    sos = scipy.signal.butter(N=order, Wn=frequency, btype=btype, output="sos", fs=sampling_rate)
    return scipy.signal.sosfiltfilt(sos, x, axis=-1).astype(x.dtype)


def snr(x, signal_mask, noise_mask):
    """max |x| over the signal window / rms over the noise window."""
    n = np.sqrt(np.mean(np.square(x[..., noise_mask]), axis=-1))
    xs = np.abs(x[..., signal_mask])
    idx = np.argmax(xs, axis=-1)
    s = np.take_along_axis(xs, idx[..., None], axis=-1)[..., 0]
    return s / n, s, n, np.flatnonzero(signal_mask)[idx]


def postprocess(cc, weights, delay, lag0, sampling_rate, frequency=3.0, btype="highpass", order=2):
    """core/postprocess.py:133-141 on cc[..., lag] already cut to the lag window."""
    y = cc * weights.astype(cc.dtype)
    y = demean(y)
    y = taper(y, sampling_rate, max_length=5.0)
    y = timeshift(y, delay, lag0, 1.0 / sampling_rate)
    y = filter(y, sampling_rate, frequency, btype, order)
    y = taper(y, sampling_rate, max_length=3 / 2)
    return y


def spectrogram(x, sampling_rate, duration, padding_factor=2, scaling="density", **kwargs):
    """signal/spectrogram.py:116-145: scipy.signal.spectrogram, one segment per sample, padded to the lag axis."""
    win_len = int(duration * sampling_rate)
    nfft = int(win_len * padding_factor)
    edge = int(np.rint(win_len / 2))
    _f, _t, Sxx = scipy.signal.spectrogram(x=x, fs=sampling_rate, nperseg=win_len, noverlap=win_len - 1, nfft=nfft,
                                           scaling=scaling, mode="psd", axis=-1, return_onesided=True, **kwargs)
    npad = [(0, 0)] * (len(Sxx.shape) - 1) + [(edge, x.shape[-1] - Sxx.shape[-1] - edge)]
    return np.pad(Sxx, pad_width=npad, mode="constant", constant_values=0), np.linspace(0.0, sampling_rate / 2, int(nfft / 2 + 1))
