"""Mirror of the part of the reference's ``xcorr.signal`` that sits on the
FFT-correlation path: ``correlate1d`` / ``correlate2d`` (the timelapse kernel,
``signal/correlate.py``), the norms they use (``signal/normalize.py``), the DFT
wrappers (``signal/fft.py``) and the CCF post-processing operators chained by
``core/postprocess.py`` (unbias, demean, taper, timeshift, filter, snr)."""
from .correlate import correlate1d, correlate2d, Correlate2dPlan, timelapse_pairs  # noqa: F401
from .normalize import norm1d, norm2d  # noqa: F401
from .fft import DFT, fft, ifft, rfft, irfft  # noqa: F401
from .postprocess_ops import (unbias, unbias_weights, demean, detrend, rms, taper, window, timeshift, filter,  # noqa: F401
                              snr)
from .spectrogram import spectrogram  # noqa: F401
