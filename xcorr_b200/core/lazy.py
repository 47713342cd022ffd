r"""
:mod:`core.lazy` -- Lazy (day-batched)
======================================

The reference distributes one Dask task per (pair, day) file
(``xcorr/core/lazy.py:26-82`` ``single_threaded_process``: read the file if it exists, ``init`` otherwise,
``process(retry_missing=True)``, ``write``), so the receivers a day's pairs share are fetched,
preprocessed and transformed once per pair.  Here the unit a scheduler hands out is the DAY: all (pair,
day) tasks of one day go through ONE batched :func:`xcorr_b200.core.process` call -- every receiver
transformed once, all pairs correlated on the GPU -- and the results are split back into the
reference's per-pair files.  Dask (or any scheduler) keeps its I/O and bookkeeping role: it maps
:func:`process_day` over the days returned by :func:`day_tasks`.
"""
import os

import numpy as np
import pandas as pd

from .init import init, CCDataset, Variable
from .process import process
from ..io import ncfile, read, write

__all__ = ['process_day', 'day_tasks', 'lazy_process', 'select_pairs']


def day_tasks(pairs, times, done=None):
    """Group the reference's (pair, day) tasks by day: ``{day: [pairs still to do]}``.
    `done(pair, day) -> bool` (optional) filters finished tasks (``core/lazy.py:130-176`` uses the
    availability / preprocessing status arrays for that)."""
    out = {}
    for t in pd.DatetimeIndex(times):
        todo = [p for p in pairs if not (done and done(p, t))]
        if todo:
            out[pd.Timestamp(t)] = todo
    return out


def select_pairs(ds, index):
    """The sub-set of a multi-pair variable set holding the pairs `index` (views, no copy of ``cc``)."""
    index = list(index)
    out = CCDataset()
    out.attrs = dict(ds.attrs)
    for k, v in ds.coords.items():
        data = v.data[index] if v.dims == ('pair',) else v.data
        out.coords[k] = Variable(v.dims, data, dict(v.attrs), dict(v.encoding))
    for k, v in ds.variables.items():
        data = v.data[index] if v.dims and v.dims[0] == 'pair' else v.data
        out.variables[k] = Variable(v.dims, data, dict(v.attrs), dict(v.encoding))
    return out


def process_day(pairs, time, init_args, waveforms, root, freq='1D', force_fresh=False, verb=0, **kwargs):
    """All (pair, day) tasks of one day (``core/lazy.py:26-82`` for every pair at once).

    Files that exist and are complete are skipped; partially processed files are read and their
    finished cells kept (``retry_missing=True`` reprocesses the rest); everything else is initialised.
    One :func:`process` call fills all pairs; every pair with at least one processed cell is written to
    ``{root}/{YYYY}/{pair}/{pair}.{YYYY}.{DOY}.nc``.  Returns ``{pair: processed}`` like the reference's
    task result.
    """
    time = pd.Timestamp(time)
    end = time + pd.Timedelta(freq)
    result, existing, todo = {}, {}, []
    for pair in pairs:
        nc = ncfile(pair, time, root)
        data = None
        if not force_fresh and os.path.exists(nc):
            data = read(nc, fast=True)
            if data is not None and np.all(data.variables['status'].data == 1):
                result[pair] = True
                continue
        if data is not None:
            existing[pair] = data
        todo.append(pair)
    if not todo:
        return result
    ds = init(pair=todo, starttime=time, endtime=end, **init_args)
    for q, pair in enumerate(todo):                 # resume: keep the cells a previous run finished
        old = existing.get(pair)
        if old is None or old.variables['cc'].data.shape[1:] != ds.variables['cc'].data.shape[1:]:
            continue
        for name in ('cc', 'status', 'pair_offset', 'time_offset'):
            ds.variables[name].data[q] = old.variables[name].data[0]
    process(ds, waveforms, retry_missing=True, verb=verb, **kwargs)
    for q, pair in enumerate(todo):
        one = select_pairs(ds, [q])
        if np.any(one.variables['status'].data == 1):
            write(one, ncfile(pair, time, root), verb=-1 if verb < 1 else verb)
            result[pair] = True
        else:
            result[pair] = False
    return result


def lazy_process(pairs, times, init_args, waveforms, root, map_fn=map, **kwargs):
    """``core/lazy.py:179-...``: run :func:`process_day` for every day.  `map_fn` is where a scheduler
    plugs in (``dask.bag`` / ``client.map`` / a process pool with one GPU per worker); the default runs
    the days in order on the current device.  Returns ``{day: {pair: processed}}``."""
    tasks = day_tasks(pairs, times)
    days = list(tasks)
    res = map_fn(lambda d: process_day(tasks[d], d, init_args, waveforms, root, **kwargs), days)
    return dict(zip(days, res))
