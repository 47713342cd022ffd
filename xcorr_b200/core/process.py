r"""
:mod:`core.process` -- Process
==============================

Batched replacement of the reference's serial ``pair x time`` loop
(``xcorr/core/process.py:82-165``).  The reference fetches, preprocesses and
transforms both receivers again for every pair and every window; here every
(receiver, window) trace is fetched once, preprocessed on the GPU as one batch
per channel, transformed once, and reused by all the pairs it belongs to.
"""
import numpy as np
import pandas as pd

from . import init as _init  # noqa: F401
from ..stream import process as stream_process

__all__ = ['process']


def _channel(receiver):
    return receiver.split('.')[-1]


def process(dataset, waveforms, retry_missing: bool = False, test_run: bool = False,
            windows_per_batch: int = 8, device=None, verb: int = 1):
    """
    Process the variable set created by :func:`xcorr_b200.core.init`.

    Parameters
    ----------
    dataset : :class:`xcorr_b200.core.CCDataset`
        Filled in place: ``cc``, ``status``, ``pair_offset``, ``time_offset``.

    waveforms : callable ``(receiver: str, time: pd.Timestamp) -> (data, sampling_rate, start_offset) | None``
        Raw trace covering the window centred on `time` (length
        ``window_length``), as a 1-D float array, with its sampling rate and the
        offset in seconds of its first sample from the nominal window start.
        ``None`` marks missing data -- the cell gets status -1 and its ``cc``
        stays zero, as ``core/process.py:122-132`` does.  (The reference gets
        this from ``client.get_pair_processed_waveforms``; I/O stays on the host.)

    retry_missing : `bool`
        Also process cells with status -1 (``core/process.py:93-104``).

    windows_per_batch : `int`
        Windows transformed per GPU batch.
    """
    import torch
    from ..plan import CorrelationPlan

    dataset.attrs['history'] += f", Process started @ {pd.Timestamp.now()}"
    lag = dataset.coords['lag']
    npts = lag.attrs['npts']
    fs = lag.attrs['sampling_rate']
    # util/time.py:94-108 (update_lag_indices)
    lag_max = (npts - 1) * lag.attrs['delta']
    i0 = int(round((lag.data[0] + lag_max) * fs))
    i1 = int(round((lag.data[-1] + lag_max) * fs + 1))
    lag.attrs['index_min'], lag.attrs['index_max'] = i0, i1

    pairs = [str(p) for p in dataset.coords['pair'].data]
    times = pd.DatetimeIndex(dataset.coords['time'].data)
    ops = dataset.coords['pair'].attrs['preprocess']
    window_length = dataset.coords['time'].attrs['window_length']
    cc = dataset.variables['cc']
    status = dataset.variables['status'].data
    receivers = sorted({r for p in pairs for r in p.split('-')})
    ridx = {r: k for k, r in enumerate(receivers)}
    pair_idx = [(ridx[p.split('-')[0]], ridx[p.split('-')[1]]) for p in pairs]
    dtype = torch.float32 if cc.data.dtype == np.float32 else torch.float64
    dev = torch.device(device or 'cuda')
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())

    todo = [w for w in range(len(times))
            if any(status[q, w] == 0 or (retry_missing and status[q, w] == -1) for q in range(len(pairs)))]
    if test_run:
        todo = todo[:1]
    plans = {}
    for b0 in range(0, len(todo), windows_per_batch):
        batch = todo[b0:b0 + windows_per_batch]
        W = len(batch)
        # ---- gather: every (receiver, window) once ----
        raw, offsets = {}, np.full((len(receivers), W), np.nan)
        for k, r in enumerate(receivers):
            for j, w in enumerate(batch):
                got = waveforms(r, times[w])
                if got is None:
                    continue
                data, rate, off = got
                raw[(k, j)] = (np.asarray(data), float(rate))
                offsets[k, j] = off
        # ---- preprocess on the GPU, one batch per (channel, length, rate) ----
        x = torch.zeros((len(receivers), W, npts), dtype=dtype, device=dev)
        ok = np.zeros((len(receivers), W), dtype=bool)
        groups = {}
        for (k, j), (data, rate) in raw.items():
            groups.setdefault((_channel(receivers[k]), len(data), rate), []).append((k, j))
        for (chan, n_raw, rate), members in groups.items():
            if chan not in ops:
                continue
            stack = torch.from_numpy(np.stack([raw[m][0] for m in members]).astype(np.float64)).to(dev)
            out = stream_process.process(stack, [tuple(o) for o in ops[chan]], sampling_rate=rate,
                                         out_dtype=dtype)
            if out is None:
                continue
            y, rate_out = out
            # sampling-rate and strict duration checks of stream/client.py:1020-1050
            if abs(rate_out - fs) > 1e-9 * fs or y.shape[1] < npts:
                continue
            for m, (k, j) in enumerate(members):
                x[k, j] = y[m, :npts]               # core/process.py:156-157 slices [:npts]
                ok[k, j] = True
        # ---- correlate all pairs of the batch ----
        if W not in plans:
            plans[W] = CorrelationPlan(npts, len(receivers), pair_idx, n_windows=W, dtype=dtype,
                                       index_min=i0, index_max=i1,
                                       normalize=cc.attrs['normalize'] == 1,
                                       unbiased=cc.attrs['bias_correct'] == 1, device=dev)
        ccf = plans[W](x).cpu().numpy()
        # ---- scatter ----
        for q, (a, b) in enumerate(pair_idx):
            for j, w in enumerate(batch):
                if not (status[q, w] == 0 or (retry_missing and status[q, w] == -1)):
                    continue
                if not (ok[a, j] and ok[b, j]):
                    status[q, w] = -1
                    if verb > 0:
                        print(pairs[q], str(times[w])[:19], 'Missing data. Set status "-1" and skip.')
                    continue
                cc.data[q, w] = ccf[q, j]
                dataset.variables['pair_offset'].data[q, w] = offsets[a, j] - offsets[b, j]
                dataset.variables['time_offset'].data[q, w] = offsets[a, j]      # start(receiver 0) - (time - window_length/2)
                status[q, w] = 1
    dataset.attrs['history'] += f", Process ended @ {pd.Timestamp.now()}"
    if cc.attrs['bias_correct'] == 1 and cc.attrs['unbiased'] == 0:
        # fused into the inverse transform's epilogue (signal/unbias.py:85)
        cc.attrs['unbiased'] = np.byte(1)
        cc.attrs['long_name'] = 'Unbiased ' + cc.attrs['long_name']
        cc.attrs['standard_name'] = 'unbiased_' + cc.attrs['standard_name']
        dataset.attrs['history'] += f", Unbiased CC @ {pd.Timestamp.now()}"
    return dataset
