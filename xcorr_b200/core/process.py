r"""
:mod:`core.process` -- Process
==============================

Batched replacement of the reference's serial ``pair x time`` loop
(``xcorr/core/process.py:82-165``).  The reference fetches, preprocesses and
transforms both receivers again for every pair and every window; here every
(receiver, window) trace is fetched once, preprocessed on the GPU as one batch
per channel, transformed once, and reused by all the pairs it belongs to.

Windowing follows the reference's client (``stream/client.py:499-519, 996-1050``):
a trace may start before the nominal window (the client's 5 % buffer); the op-list
runs on the buffered trace with the window's start / end injected into ``trim``
(``stream/process.py:86-89,153-196``); the result is trimmed to the window
(``client.py:1017``), checked for sampling rate and strict duration (``:1020-1050``)
and its first ``npts`` samples are correlated (``core/process.py:156-157``).

Host <-> device traffic is pipelined: the raw traces of batch ``b+1`` are gathered
into page-locked staging buffers and copied in while batch ``b`` computes, and the
correlation functions of batch ``b`` are copied out (straight into ``dataset.cc``
when :func:`xcorr_b200.core.init` placed it in page-locked memory) while batch
``b+1`` runs.
"""
import os
import time

import numpy as np
import pandas as pd

from . import init as _init  # noqa: F401
from ..stream import process as stream_process

__all__ = ['process', 'compile_operations', 'clear_caches']


def _channel(receiver):
    return receiver.split('.')[-1]


def compile_operations(ops, channels, window_length):
    """Validate every channel's op-list ONCE, before any window is touched
    (``stream/process.py:29-94`` whitelist, argument checks of the compiler): a
    configuration error raises here instead of turning every cell into "missing data".
    Returns ``{channel: [(op, params), ...]}`` with ``trim`` completed by the window
    (``inject_dynamic_parameters``)."""
    out = {}
    for chan in channels:
        if chan not in ops:
            raise KeyError(f"preprocess operations for channel '{chan}' are missing")
        lst = []
        for item in ops[chan]:
            if not isinstance(item, (tuple, list)) or len(item) != 2:
                raise TypeError('Provided operation should be a tuple or list with length 2 (method:str,params:dict).')
            op, params = item
            if not stream_process.is_operation(op):
                raise ValueError(f'Provided operation "{op}" is invalid.')
            if op not in stream_process.device_operations():
                raise NotImplementedError(
                    f'operation "{op}" acts on obspy objects / station metadata and is not part of the '
                    'array-valued GPU path; apply it with the reference client before the hand-off')
            if not isinstance(params, dict):
                raise TypeError(f'parameters of operation "{op}" should be a dictionary')
            lst.append((op, stream_process.inject_dynamic_parameters(op, params, starttime=0.0,
                                                                     endtime=float(window_length))))
        out[chan] = lst
    return out


# Plans, preprocessing programs and the page-locked staging buffers outlive one call: the reference's unit of
# work is one (pair, day) file (core/lazy.py:26-82), so process() is called over and over with the same shapes,
# and creating a plan (tables, workspace) or page-locking memory costs more than a window of correlations.
_PLAN_CACHE, _PROGRAM_CACHE, _STAGERS = {}, {}, {}
last_stats = {}                # of the most recent process() call: {'h2d_bytes': ...}
_PLAN_CACHE_MAX = 3


def clear_caches():
    """Release the cached plans, programs and staging buffers."""
    _PLAN_CACHE.clear()
    _PROGRAM_CACHE.clear()
    _STAGERS.clear()


def _cached_plan(key, make):
    if key in _PLAN_CACHE:
        _PLAN_CACHE[key] = _PLAN_CACHE.pop(key)           # most recently used last
        return _PLAN_CACHE[key]
    while len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
        _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
    _PLAN_CACHE[key] = make()
    return _PLAN_CACHE[key]


def _shared_base(arrs):
    """If all arrays are contiguous 1-D views of one longer contiguous 1-D array and overlap enough that
    their union is smaller than their sum, return ``(base, element offsets, lo, hi)``; else None."""
    if len(arrs) < 2:
        return None
    base = arrs[0].base
    if base is None or not isinstance(base, np.ndarray) or base.ndim != 1 or not base.flags.c_contiguous:
        return None
    p0, item = base.__array_interface__['data'][0], base.itemsize
    offs = []
    for a in arrs:
        if a.base is not base or a.ndim != 1 or not a.flags.c_contiguous or a.dtype != base.dtype:
            return None
        offs.append((a.__array_interface__['data'][0] - p0) // item)
    lo, hi = min(offs), max(o + len(a) for o, a in zip(offs, arrs))
    if hi - lo >= sum(len(a) for a in arrs):
        return None
    return base, offs, lo, hi


class _Stager:
    """Two page-locked host buffers (+ device twins) reused round-robin."""

    def __init__(self, torch, dev):
        self.torch, self.dev = torch, dev
        self.host, self.device, self.free_ev = {}, {}, {}

    def get(self, slot, key, shape, dtype, host=True, device=True, zero=False):
        torch = self.torch
        k = (slot, key)
        n = int(np.prod(shape))
        h = d = None
        if device:
            if k not in self.device or self.device[k].numel() < n or self.device[k].dtype != dtype:
                self.device[k] = (torch.zeros if zero else torch.empty)(n, dtype=dtype, device=self.dev)
            d = self.device[k][:n].view(*shape)
        if host:
            if k not in self.host or self.host[k].numel() < n or self.host[k].dtype != dtype:
                self.host[k] = torch.empty(n, dtype=dtype, pin_memory=True)
            h = self.host[k][:n].view(*shape)
        return h, d


def process(dataset, waveforms, retry_missing: bool = False, test_run: bool = False,
            windows_per_batch: int = 8, device=None, verb: int = 1):
    """
    Process the variable set created by :func:`xcorr_b200.core.init`.

    Parameters
    ----------
    dataset : :class:`xcorr_b200.core.CCDataset`
        Filled in place: ``cc``, ``status``, ``pair_offset``, ``time_offset``.

    waveforms : callable ``(receiver: str, time: pd.Timestamp) -> (data, sampling_rate, start_offset) | None``
        Raw trace for the window centred on `time` (length ``window_length``) as a
        1-D float array, its sampling rate, and the offset in seconds of its first
        sample from the nominal window start -- negative when the trace carries a
        buffer in front of the window like the reference client's 5 %
        (``stream/client.py:499-519``).  ``None`` marks missing data: the cell gets
        status -1 and its ``cc`` stays zero (``core/process.py:122-132``).  (The
        reference gets this from ``client.get_pair_processed_waveforms``; I/O stays
        on the host.)

    retry_missing : `bool`
        Also process cells with status -1 (``core/process.py:93-104``).

    windows_per_batch : `int`
        Windows transformed per GPU batch.
    """
    import torch
    from ..plan import CorrelationPlan
    from ..io.hasher import hash_dataset

    dataset.attrs['history'] += f", Process started @ {pd.Timestamp.now()}"
    lag = dataset.coords['lag']
    npts = lag.attrs['npts']
    fs = lag.attrs['sampling_rate']
    # util/time.py:94-108 (update_lag_indices)
    lag_max = (npts - 1) * lag.attrs['delta']
    i0 = int(round((lag.data[0] + lag_max) * fs))
    i1 = int(round((lag.data[-1] + lag_max) * fs + 1))
    lag.attrs['index_min'], lag.attrs['index_max'] = i0, i1
    K = i1 - i0

    pairs = [str(p) for p in dataset.coords['pair'].data]
    times = pd.DatetimeIndex(dataset.coords['time'].data)
    window_length = float(dataset.coords['time'].attrs['window_length'])
    cc = dataset.variables['cc']
    status = dataset.variables['status'].data
    receivers = sorted({r for p in pairs for r in p.split('-')})
    ridx = {r: k for k, r in enumerate(receivers)}
    pair_idx = [(ridx[p.split('-')[0]], ridx[p.split('-')[1]]) for p in pairs]
    R, P = len(receivers), len(pairs)
    # configuration errors raise before the loop (they are not "missing data")
    ops = compile_operations(dataset.coords['pair'].attrs['preprocess'], sorted({_channel(r) for r in receivers}),
                             window_length)
    dtype = torch.float32 if cc.data.dtype == np.float32 else torch.float64
    dev = torch.device(device or 'cuda')
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    cc_t = torch.from_numpy(cc.data)
    cc_pinned = cc_t.is_pinned() and cc_t.is_contiguous()

    todo = [w for w in range(len(times))
            if any(status[q, w] == 0 or (retry_missing and status[q, w] == -1) for q in range(P))]
    if test_run:
        todo = todo[:1]
    batches = [todo[b:b + windows_per_batch] for b in range(0, len(todo), windows_per_batch)]

    ops_key = stream_process.operations_to_json(dict(dataset.coords['pair'].attrs['preprocess']))
    programs = _PROGRAM_CACHE.setdefault((ops_key, window_length), {})
    with torch.cuda.device(dev):
        s_in, s_run, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        stager = _STAGERS.setdefault(str(dev), _Stager(torch, dev))
        x_dev = [None, None]
        ccf_dev = [None, None]
        ev_run = [None, None]                 # compute of the batch that last used slot s
        ev_out = [None, None]                 # its device -> host copy
        h2d_bytes = [0]
        trace = [] if os.environ.get('XCB_TIMELINE') else None      # per-batch device / host time stamps (diagnostic)

        def mark(stream):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            return ev

        def program_for(chan, n_raw, rate, off):
            key = (chan, n_raw, rate, off)
            if key not in programs:
                try:
                    programs[key] = stream_process.PreprocessProgram(ops[chan], n_raw, rate, trace_start=off)
                except ValueError as error:   # the trace does not cover the window / cannot take the ops: missing data
                    if verb > 0:
                        print(f'{chan}: {n_raw} samples at {rate} Hz cannot be processed ({error})')
                    programs[key] = None      # (TypeError / NotImplementedError / ArithmeticError = configuration: raise)
            return programs[key]

        def launch(slot, batch):
            """Gather + async H2D + preprocess + correlate + async D2H of one batch."""
            W = len(batch)
            tl = {'host_launch0': time.perf_counter()} if trace is not None else None
            raw, offsets = {}, np.full((R, W), np.nan)
            for k, r in enumerate(receivers):
                for j, w in enumerate(batch):
                    got = waveforms(r, times[w])
                    if got is None:
                        continue
                    data, rate, off = got
                    data = np.asarray(data)
                    if data.dtype not in (np.float32, np.float64):
                        data = data.astype(np.float64)            # obspy promotes integer data to float64
                    raw[(k, j)] = (data, float(rate), float(off))
            # (persistent across calls like the staging buffers: a fresh cudaMalloc in the first batch of every
            # call stalled the host for ~0.1 s; traces that stay missing keep the zeros of the allocation)
            _, x_dev[slot] = stager.get(slot, ('x', R, W, npts), (R, W, npts), dtype, host=False, zero=True)
            _, ccf_dev[slot] = stager.get(slot, ('ccf', P, W, K), (P, W, K), dtype, host=False)
            x = x_dev[slot]
            ok = np.zeros((R, W), dtype=bool)
            groups = {}
            for (k, j), (data, rate, off) in raw.items():
                groups.setdefault((_channel(receivers[k]), len(data), rate, off, data.dtype), []).append((k, j))
            if ev_run[slot] is not None:
                s_in.wait_event(ev_run[slot])                     # x_dev[slot] and the staging twins are free again
                ev_run[slot].synchronize()                        # ... and so are the page-locked buffers on the host
            if tl is not None:
                tl['host_gathered'] = time.perf_counter()
                tl['h2d0'] = mark(s_in)
            staged = []
            for gi, ((chan, n_raw, rate, off, ndt), members) in enumerate(groups.items()):
                prog = program_for(chan, n_raw, rate, off)
                if prog is None:
                    continue
                # client.py:1017: trim the processed trace to the window, then the checks of :1020-1050
                skip, keep = stream_process.trim_samples(prog.npts_out, prog.fs_out, prog.trace_start_out,
                                                         0.0, window_length)
                if not np.isclose(prog.fs_out, fs, atol=0.0) or window_length * prog.fs_out - keep > 0:
                    continue
                start_final = prog.trace_start_out + skip / prog.fs_out
                tdt = torch.float32 if ndt == np.float32 else torch.float64
                arrs = [raw[m][0] for m in members]
                _, devbuf = stager.get(slot, gi, (len(members), n_raw), tdt, host=False)
                # overlapping windows (core/init.py:28, window_overlap = 0.875: eight windows a day share 87.5 % of
                # their samples) handed over as views of ONE longer trace per receiver: the union of a receiver's
                # windows goes to the device once and the windows are cut there
                by_base = {}
                for m, a in enumerate(arrs):
                    by_base.setdefault(id(a.base) if a.base is not None else ('own', m), []).append(m)
                loose = []
                for bi, (key, ms) in enumerate(by_base.items()):
                    shared = _shared_base([arrs[m] for m in ms])
                    if shared is None:
                        loose += ms
                        continue
                    base, offs, lo, hi = shared
                    src = torch.from_numpy(base[lo:hi])
                    hb, devbase = stager.get(slot, ('base', gi, bi), (hi - lo,), tdt, host=not src.is_pinned())
                    with torch.cuda.stream(s_in):
                        if hb is not None:
                            hb.numpy()[...] = base[lo:hi]
                            src = hb
                        devbase.copy_(src, non_blocking=True)
                        for m, o in zip(ms, offs):
                            devbuf[m].copy_(devbase[o - lo:o - lo + n_raw], non_blocking=True)
                    h2d_bytes[0] += (hi - lo) * devbase.element_size()
                if loose:
                    srcs = [torch.from_numpy(arrs[m]) for m in loose]
                    if all(t.is_pinned() and t.is_contiguous() for t in srcs):
                        # the caller's I/O layer already reads into page-locked memory: DMA straight from it
                        with torch.cuda.stream(s_in):
                            for m, t in zip(loose, srcs):
                                devbuf[m].copy_(t, non_blocking=True)
                    else:
                        host, devtmp = stager.get(slot, ('loose', gi), (len(loose), n_raw), tdt,
                                                  device=len(loose) != len(members))
                        hnp = host.numpy()
                        for k2, m in enumerate(loose):
                            hnp[k2] = arrs[m]
                        with torch.cuda.stream(s_in):
                            if len(loose) == len(members):
                                devbuf.copy_(host, non_blocking=True)
                            else:
                                devtmp.copy_(host, non_blocking=True)
                                devbuf[torch.tensor(loose, device=dev)] = devtmp
                    h2d_bytes[0] += len(loose) * n_raw * devbuf.element_size()
                staged.append((prog, devbuf, members, skip, start_final))
            ev_in = torch.cuda.Event(enable_timing=trace is not None)
            ev_in.record(s_in)
            if tl is not None:
                tl['host_h2d_queued'] = time.perf_counter()
            with torch.cuda.stream(s_run):
                s_run.wait_event(ev_in)
                if ev_out[slot] is not None:
                    s_run.wait_event(ev_out[slot])                # ccf_dev[slot] drained
                if tl is not None:
                    tl['h2d1'] = ev_in
                    tl['run0'] = mark(s_run)
                whole = [(k, j) for k in range(R) for j in range(W)]
                for prog, devbuf, members, skip, start_final in staged:
                    y = prog.run(devbuf, out_dtype=dtype, copy=False)       # devbuf is ours: no defensive copy
                    for (k, j) in members:
                        ok[k, j] = True
                        offsets[k, j] = start_final
                    if len(staged) == 1 and members == whole and y.dtype == dtype and y.stride(1) == 1:
                        # every trace of the batch came out of one program, in order: the plan reads the
                        # processed traces where they are (rows of y, core/process.py:156-157 slices [:npts])
                        x = y[:, skip:skip + npts]
                    else:
                        for m, (k, j) in enumerate(members):       # (device -> device, asynchronous: an index
                            x[k, j].copy_(y[m, skip:skip + npts], non_blocking=True)   # tensor would cost a synchronous upload)
                if tl is not None:
                    tl['host_preprocessed'] = time.perf_counter()
                normalize, unbiased = cc.attrs['normalize'] == 1, cc.attrs['bias_correct'] == 1
                plan = _cached_plan(
                    (str(dev), npts, R, tuple(pair_idx), W, str(dtype), i0, i1, bool(normalize), bool(unbiased)),
                    lambda: CorrelationPlan(npts, R, pair_idx, n_windows=W, dtype=dtype, index_min=i0, index_max=i1,
                                            normalize=normalize, unbiased=unbiased, device=dev))
                plan(x, out=ccf_dev[slot])
                ev_run[slot] = torch.cuda.Event(enable_timing=trace is not None)
                ev_run[slot].record(s_run)
            if tl is not None:
                tl['host_run_queued'] = time.perf_counter()
            # cells of this batch that get a result
            cells = []
            for q, (a, b) in enumerate(pair_idx):
                for j, w in enumerate(batch):
                    if not (status[q, w] == 0 or (retry_missing and status[q, w] == -1)):
                        continue
                    cells.append((q, j, w, bool(ok[a, j] and ok[b, j])))
            direct = cc_pinned and all(good for *_, good in cells) and len(cells) == P * W and \
                batch == list(range(batch[0], batch[0] + W))
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_run[slot])
                if tl is not None:
                    tl['run1'] = ev_run[slot]
                    tl['d2h0'] = mark(s_out)
                if direct:
                    # rows of the page-locked dataset array are contiguous per pair: one DMA per pair
                    if W == cc_t.shape[1]:
                        cc_t.copy_(ccf_dev[slot], non_blocking=True)
                    else:
                        for q in range(P):
                            cc_t[q, batch[0]:batch[0] + W].copy_(ccf_dev[slot][q], non_blocking=True)
                    host_out = None
                else:
                    host_out, _ = stager.get(slot, 'out', (P, W, K), dtype, device=False)
                    host_out.copy_(ccf_dev[slot], non_blocking=True)
                ev_out[slot] = torch.cuda.Event(enable_timing=trace is not None)
                ev_out[slot].record(s_out)
            if tl is not None:
                tl['d2h1'] = ev_out[slot]
                tl['host_launch1'] = time.perf_counter()
                trace.append(tl)
            return batch, cells, offsets, host_out, ev_out[slot]

        def finish(pending):
            batch, cells, offsets, host_out, ev = pending
            ev.synchronize()
            out_np = host_out.numpy() if host_out is not None else None
            for q, j, w, good in cells:
                a, b = pair_idx[q]
                if not good:
                    status[q, w] = -1
                    if verb > 0:
                        print(pairs[q], str(times[w])[:19], 'Missing data. Set status "-1" and skip.')
                    continue
                if out_np is not None:
                    cc.data[q, w] = out_np[q, j]
                dataset.variables['pair_offset'].data[q, w] = offsets[a, j] - offsets[b, j]
                dataset.variables['time_offset'].data[q, w] = offsets[a, j]      # start(receiver 0) - (time - window_length/2)
                status[q, w] = 1

        pending = None
        for b, batch in enumerate(batches):
            nxt = launch(b % 2, batch)
            if pending is not None:
                finish(pending)
            pending = nxt
        if pending is not None:
            finish(pending)
        torch.cuda.synchronize(dev)
        if trace:
            t0, h0 = trace[0]['h2d0'], trace[0]['host_launch0']
            last_stats['timeline'] = [
                {k: (round(t0.elapsed_time(v), 3) if not isinstance(v, float) else round((v - h0) * 1e3, 3))
                 for k, v in tl.items()} for tl in trace]

    dataset.attrs['history'] += f", Process ended @ {pd.Timestamp.now()}"
    if cc.attrs['bias_correct'] == 1 and cc.attrs['unbiased'] == 0:
        # fused into the inverse transform's epilogue (signal/unbias.py:85)
        cc.attrs['unbiased'] = np.byte(1)
        cc.attrs['long_name'] = 'Unbiased ' + cc.attrs['long_name']
        cc.attrs['standard_name'] = 'unbiased_' + cc.attrs['standard_name']
        dataset.attrs['history'] += f", Unbiased CC @ {pd.Timestamp.now()}"
    last_stats['h2d_bytes'] = int(h2d_bytes[0])                     # bookkeeping: bytes copied host -> device
    # update metadata hash (core/process.py:183)
    dataset.attrs['sha256_hash_metadata'] = hash_dataset(dataset, metadata_only=True)
    return dataset
