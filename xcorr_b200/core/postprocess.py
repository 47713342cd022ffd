r"""
:mod:`core.postprocess` -- predefined CCF post-processing on the device
=======================================================================

Mirror of ``xcorr/core/postprocess.py:22-150``: extract the valid windows
(``status == 1``) and the lag window of interest, then on every CCF

    unbias -> demean -> taper(max_length=5 s) -> timeshift(-(pair_offset + time_offset))
           -> filter(highpass 3 Hz, order 2, zero phase) -> taper(max_length=1.5 s)

(``core/postprocess.py:133-141``).  The arithmetic runs on the GPU through
:mod:`xcorr_b200.signal` (one batched launch chain for all pair x time CCFs); the
selection and bookkeeping stay numpy, like the reference's xarray indexing.
"""
import warnings

import numpy as np
import pandas as pd

from .init import CCDataset, Variable

__all__ = ['postprocess']


def postprocess(ds: CCDataset, lag_lim: tuple = None, time_lim: tuple = None, clim: tuple = None,
                cmin: float = None, cmax: float = None, time_min=None, time_max=None,
                lag_min: float = None, lag_max: float = None, filter_kwargs: dict = None, device=None):
    """Postprocess a cc variable set (same arguments as the reference).

    Returns a new :class:`CCDataset` holding the valid times only, with ``cc``
    (and ``lag``) cut to the lag window and replaced by the postprocessed CCFs.
    """
    import torch
    from .. import signal as sg

    cc = ds.variables['cc']
    if 'postprocess' in cc.attrs:
        warnings.warn(f"Dataset already postprocessed on {cc.attrs['postprocess']}")
        return ds
    if lag_lim is not None:
        lag_min, lag_max = lag_lim
    if time_lim is not None:
        time_min, time_max = time_lim
    if clim is not None:
        cmin, cmax = clim

    d = ds.variables['distance']
    d_fact = 1000 if (d.attrs.get('units') == 'km' and ((cmin or 0) > 10 and (cmax or 0) > 10)) else 1

    time = ds.coords['time'].data
    time_min = time_min if time_min is not None else time.min()
    time_max = time_max if time_max is not None else time.max()
    if not isinstance(time_min, np.datetime64) or not isinstance(time_max, np.datetime64):
        raise TypeError('time min and max should be of type numpy.datetime64')

    status = ds.variables['status'].data                      # (pair, time)
    tmask = (time >= time_min) & (time <= time_max)
    m = (status == 1) & tmask[None, :]
    if not m.any():
        raise ValueError("No data after extracting valid times")
    keep_t = m.any(axis=0)                                    # where(m, drop=True): times valid for any pair

    filter_kwargs = {'frequency': 3., 'btype': 'highpass', 'order': 2, **(filter_kwargs or dict())}

    delay = -(ds.variables['pair_offset'].data + ds.variables['time_offset'].data)[:, keep_t]

    lag = ds.coords['lag']
    lag_min = lag_min or lag.data.min().item()
    lag_max = lag_max or lag.data.max().item()
    lag_min = max((lag_min, d.data.min() * d_fact / cmax)) if cmax else lag_min
    lag_max = min((lag_max, d.data.max() * d_fact / cmin)) if cmin else lag_max
    lm = (lag.data >= lag_min) & (lag.data <= lag_max)
    if not lm.any():
        raise ValueError("No data after extracting time lag")
    li = np.flatnonzero(lm)
    l0, l1 = int(li[0]), int(li[-1]) + 1

    fs = lag.attrs['sampling_rate']
    npts = lag.attrs['npts']
    i0 = lag.attrs['index_min']
    dev = torch.device(device if device is not None else 'cuda')
    x = torch.from_numpy(np.ascontiguousarray(cc.data[:, keep_t, l0:l1])).to(dev)

    y = x
    if not cc.attrs.get('unbiased', 0):
        y = sg.unbias(y, sg.unbias_weights(npts, i0 + l0, i0 + l1))
    y = sg.demean(y)
    y = sg.taper(y, fs, max_length=5.)                         # timeshift phase wrapping
    y = sg.timeshift(y, delay, fs, fast=True)
    y = sg.filter(y, fs, **filter_kwargs)
    y = sg.taper(y, fs, max_length=3 / 2)                      # filter artefacts
    out_cc = y.cpu().numpy()
    # entries of invalid (pair, time) cells are NaN after ``where(m, drop=True)``
    out_cc[~m[:, keep_t]] = np.nan

    out = CCDataset()
    out.attrs = dict(ds.attrs)
    n_lag = out_cc.shape[-1]
    out.coords['pair'] = ds.coords['pair']
    out.coords['time'] = Variable(('time',), time[keep_t], ds.coords['time'].attrs)
    out.coords['lag'] = Variable(('lag',), lag.data[l0:l0 + n_lag], {**lag.attrs, 'cmin': cmin, 'cmax': cmax})
    for name in ('pair_offset', 'time_offset', 'hash'):
        if name in ds.variables:
            v = ds.variables[name]
            out.variables[name] = Variable(v.dims, v.data[:, keep_t], v.attrs, v.encoding)
    out.variables['distance'] = d
    out.variables['cc'] = Variable(cc.dims, out_cc, {**cc.attrs, 'unbiased': np.byte(1),
                                                      'postprocess': f"{pd.to_datetime('now')}"}, cc.encoding)
    return out
