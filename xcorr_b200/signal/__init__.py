"""Mirror of the part of the reference's ``xcorr.signal`` that sits on the
FFT-correlation path: ``correlate1d`` / ``correlate2d`` (the timelapse kernel,
``signal/correlate.py``) and the norms they use (``signal/normalize.py``)."""
from .correlate import correlate1d, correlate2d, Correlate2dPlan  # noqa: F401
from .normalize import norm1d, norm2d  # noqa: F401
