"""Mirror of the reference's ``xcorr.stream`` namespace for the hot path:
``xcorr_b200.stream.process.process(waveforms, operations, ...)``."""
from . import process  # noqa: F401
