"""Mirror of the reference's ``xcorr.core`` namespace for the hot path:
``init`` (CF-1.9 variable layout) and the batched ``process`` driver."""
from .init import init, CCDataset, Variable  # noqa: F401
from .process import process  # noqa: F401
