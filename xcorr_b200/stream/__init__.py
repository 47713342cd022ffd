"""Mirror of the reference's ``xcorr.stream`` namespace for the hot path."""
from . import process as _process_module  # noqa: F401
from .process import process, PreprocessProgram  # noqa: F401
