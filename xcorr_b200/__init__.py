"""xcorr_b200 -- B200-native cross-correlation hot path of psmsmets/xcorr.

``xcorr_b200.util.cc`` mirrors the reference's ``xcorr.util.cc`` (``cc``,
``lag``, ``weight``, shift helpers); :class:`CorrelationPlan` is the batched
device entry point that replaces the serial loop of ``core/process.py``.
All numerics run in hand-written sm_100a CUDA (``libxcorr_b200.so``); there is
no CPU fallback.
"""
from . import util  # noqa: F401
from .util.cc import (cc, lag, weight, extract_shift, extract_shift_and_max,  # noqa: F401
                      compute_shift, compute_shift_and_max)

__version__ = "0.1.0"


def __getattr__(name):
    # torch is imported lazily so that the numpy-only helpers stay light
    if name in ("CorrelationPlan", "all_pairs"):
        from . import plan
        return getattr(plan, name)
    raise AttributeError(name)
