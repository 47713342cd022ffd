r"""
:mod:`signal.spectrogram` -- sliding-window power spectra of CCFs on the device
===============================================================================

Mirror of ``xcorr/signal/spectrogram.py:27-197``: ``scipy.signal.spectrogram`` with
one segment per lag sample (``noverlap = win_len - 1``), ``nfft = win_len *
padding_factor``, ``mode='psd'``, centred by padding ``rint(win_len / 2)`` zero
columns on the left (the timelapse script feeds these to ``correlate2d``,
``scripts/timelapse.py:207-209``).  Window design is host-side scipy
(default ``('tukey', 0.25)``, periodic, like scipy); every per-sample operation
is ``xcb_spectrogram``.
"""
import numpy as np
import scipy.signal
import torch

from .. import _lib

__all__ = ['spectrogram']

_CODE = {torch.float32: _lib.XCB_F32, torch.float64: _lib.XCB_F64}


def spectrogram(x, sampling_rate, duration=None, padding_factor=2, scaling=None, window=('tukey', .25),
                detrend='constant'):
    """x: CUDA tensor ``[..., lag]`` -> ``(psd[..., freq, lag], freq)``.

    ``scaling``: 'density' (default, V**2/Hz) or 'spectrum' (V**2), as the reference.
    """
    if not torch.is_tensor(x) or not x.is_cuda or x.dtype not in _CODE:
        raise TypeError('x must be a float32 / float64 CUDA tensor (xcorr_b200 has no CPU path)')
    scaling = scaling or 'density'
    if scaling not in ('density', 'spectrum'):
        raise ValueError('scaling should be either "density" or "spectrum"!')
    if not isinstance(duration, float):
        raise TypeError('duration should be float seconds')
    if not isinstance(padding_factor, int) or padding_factor < 1:
        raise ValueError('padding_factor should be a positive integer')
    if detrend not in ('constant', False, None):
        raise NotImplementedError("detrend: 'constant' or False")
    win_len = int(duration * sampling_rate)
    if win_len < 16:
        raise ValueError('Change duration to have at least 16 sample points!')
    nfft = int(win_len * padding_factor)
    lead = x.shape[:-1]
    n = x.shape[-1]
    if win_len > n:
        raise ValueError('window longer than the signal')
    x2 = x.reshape(-1, n)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    w = scipy.signal.get_window(window, win_len)
    scale = 1.0 / (sampling_rate * (w * w).sum()) if scaling == 'density' else 1.0 / w.sum() ** 2
    wd = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float64)).to(x.device)
    n_bins = nfft // 2 + 1
    y = torch.empty((x2.shape[0], n_bins, n), dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().xcb_spectrogram(_CODE[x.dtype], x2.data_ptr(), x2.stride(0), n, x2.shape[0], wd.data_ptr(),
                                           win_len, nfft, float(scale), int(detrend == 'constant'), y.data_ptr(),
                                           torch.cuda.current_stream(x.device).cuda_stream), 'xcb_spectrogram')
    freq = np.linspace(0., sampling_rate / 2, int(nfft / 2 + 1))
    return y.reshape(*lead, n_bins, n), freq
