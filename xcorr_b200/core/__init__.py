"""Mirror of the reference's ``xcorr.core`` namespace for the hot path:
``init`` (CF-1.9 variable layout), the batched ``process`` driver, the day-batched task loop of
``core/lazy.py`` and the predefined ``postprocess`` chain."""
from .init import init, CCDataset, Variable  # noqa: F401
from .process import process  # noqa: F401
from .postprocess import postprocess  # noqa: F401
from .lazy import process_day, day_tasks, lazy_process, select_pairs  # noqa: F401
