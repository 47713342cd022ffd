"""Mirror of the reference's ``xcorr.util`` namespace for the hot path."""
from . import cc  # noqa: F401
