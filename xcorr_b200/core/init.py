r"""
:mod:`core.init` -- Init
========================

The reference builds an ``xarray.Dataset`` skeleton per receiver pair
(``xcorr/core/init.py:25-325``).  The hot path only needs that *layout* --
coordinates ``pair / time / lag``, variables ``distance, status, pair_offset,
time_offset, cc`` with their attributes and encodings -- so this module builds
the same structure as plain numpy variables (:class:`CCDataset`) and converts
it with :meth:`CCDataset.to_xarray` when xarray is importable (it is not in
the build image).  Several pairs share one dataset, which is what the
reference's ``merge`` produces afterwards (``core/merge.py``).
"""
import json

import numpy as np
import pandas as pd

from ..stream.process import hash_operations
from ..util import cc as _cc

__all__ = ['init', 'CCDataset', 'Variable']


class Variable:
    """``(dims, data, attrs, encoding)`` -- the tuple form xarray accepts."""

    def __init__(self, dims, data, attrs=None, encoding=None):
        self.dims = tuple(dims)
        self.data = data
        self.attrs = dict(attrs or {})
        self.encoding = dict(encoding or {})

    @property
    def values(self):
        return self.data

    def as_tuple(self):
        return (self.dims, self.data, self.attrs, self.encoding)


class CCDataset:
    """Minimal labelled container with the reference's variable layout."""

    def __init__(self):
        self.attrs = {}
        self.coords = {}
        self.variables = {}

    def __getitem__(self, name):
        return self.coords[name] if name in self.coords else self.variables[name]

    def __contains__(self, name):
        return name in self.coords or name in self.variables

    def __getattr__(self, name):
        d = self.__dict__
        if name in d.get('coords', {}):
            return d['coords'][name]
        if name in d.get('variables', {}):
            return d['variables'][name]
        raise AttributeError(name)

    def to_xarray(self):
        import xarray as xr
        ds = xr.Dataset()
        ds.attrs = dict(self.attrs)
        for k, v in self.coords.items():
            ds.coords[k] = (v.dims, v.data, _json_attrs(v.attrs))
        for k, v in self.variables.items():
            ds[k] = v.as_tuple()
        return ds


def _zeros_maybe_pinned(shape, dtype, pin_memory):
    """Zero-filled host array; page-locked when asked for (or by default with a CUDA device), so
    device-to-host copies into it run at full PCIe speed without a staging copy."""
    if pin_memory is None or pin_memory:
        try:
            import torch
            if torch.cuda.is_available():
                t = torch.zeros(shape, dtype=torch.float32 if np.dtype(dtype) == np.float32 else torch.float64,
                                pin_memory=True)
                return t.numpy()
        except Exception:
            if pin_memory:
                raise
    return np.zeros(shape, dtype=dtype)


def _json_attrs(attrs):
    return {k: (json.dumps(v) if isinstance(v, dict) else v) for k, v in attrs.items()}


def init(
    pair, starttime: pd.Timestamp, endtime: pd.Timestamp,
    preprocess: dict, attrs: dict, sampling_rate: float,
    window_length: float = 86400., window_overlap: float = 0.875,
    clip_lag=None, unbiased_cc: bool = False, closed: str = 'left',
    dtype: np.dtype = np.float32, distance=None, hash_waveforms: bool = False,
    pin_memory=None,
):
    """
    Initiate the cross-correlation variable set (``core/init.py:25``).

    Same arguments and semantics as the reference, except that `pair` may be a
    list of ``'{net}.{sta}.{loc}.{cha}-{net}.{sta}.{loc}.{cha}'`` strings (one
    dataset for all pairs) and the geodetic distance, which the reference
    derives from an obspy inventory with pyproj, is passed in (`distance`, km
    per pair, default NaN).  `pin_memory` (default: when a CUDA device is present) places the
    ``cc`` array in page-locked host memory so that :func:`xcorr_b200.core.process` copies the
    correlation functions from the device straight into it.
    """
    if not isinstance(attrs, dict):
        raise TypeError('attrs should be a dictionary')
    for key, hint in (('institution', 'Institution, department'), ('author', 'Name - E-mail'),
                      ('source', 'Data source description')):
        if key not in attrs:
            raise KeyError(f"attrs['{key}'] = '{hint}'!")
    pairs = [pair] if isinstance(pair, str) else list(pair)
    if not pairs or not all(isinstance(p, str) and '-' in p for p in pairs):
        raise TypeError("pair should be receiver pair string!")
    starttime, endtime = pd.Timestamp(starttime), pd.Timestamp(endtime)
    dtype = np.dtype(dtype).type

    # config (core/init.py:119-121)
    delta = 1 / sampling_rate
    npts = int(window_length * sampling_rate)
    encoding = dict(zlib=True, complevel=9, shuffle=True)

    ds = CCDataset()
    ds.attrs = {
        'title': ((attrs['title'] if 'title' in attrs else '') +
                  ' Cross-correlations - {}{}'.format(
                      starttime.strftime('%Y.%j'),
                      ' to {}'.format(endtime.strftime('%Y.%j'))
                      if starttime.strftime('%Y.%j') != endtime.strftime('%Y.%j') else '')).strip(),
        'institution': attrs['institution'],
        'author': attrs['author'],
        'source': attrs['source'],
        'history': 'Created @ {}'.format(pd.Timestamp.now()),
        'references': ('Bendat, J. Samuel, & Piersol, A. Gerald. (1971). '
                       'Random data : analysis and measurement procedures. '
                       'New York (N.Y.): Wiley-Interscience.'),
        'comment': attrs['comment'] if 'comment' in attrs else 'n/a',
        'Conventions': 'CF-1.9',
        'xcorr_version': 'xcorr_b200-0.1.0',
    }

    # pair (core/init.py:154-160)
    ds.coords['pair'] = Variable(('pair',), np.array(pairs, dtype=object), {
        'long_name': 'Cross-correlation receiver pair',
        'standard_name': 'receiver_pair',
        'units': '-',
        'preprocess': hash_operations(preprocess),
    })

    # time (core/init.py:163-173); pandas >= 1.4 spells `closed` as `inclusive`
    inclusive = {None: 'both', 'left': 'left', 'right': 'right'}[closed]
    times = pd.date_range(start=starttime, end=endtime,
                          freq='{0:.0f}s'.format(window_length * (1 - window_overlap)), inclusive=inclusive)
    ds.coords['time'] = Variable(('time',), times.values.astype('datetime64[ns]'), {
        'window_length': window_length,
        'window_overlap': window_overlap,
        'closed': closed,
    })

    # time lag (core/init.py:176-209)
    lag = _cc.lag(npts, delta, pad=True)
    if clip_lag is not None:
        msg = ('``clip_lag`` should be in seconds of type `float` or of type'
               '`tuple` with length 2 specifying start and end.')
        if isinstance(clip_lag, float):
            clip_lag = (-abs(clip_lag), abs(clip_lag))      # the reference's tuple(-c, c) is a TypeError
        elif isinstance(clip_lag, tuple) and len(clip_lag) == 2:
            if not isinstance(clip_lag[0], float) or not isinstance(clip_lag[1], float):
                raise TypeError(msg)
        else:
            raise TypeError(msg)
        nmin = int(np.argmin(abs(lag - clip_lag[0])))
        nmax = int(np.argmin(abs(lag - clip_lag[1])))
    else:
        nmin, nmax = 0, 2 * npts - 1
    ds.coords['lag'] = Variable(('lag',), lag[nmin:nmax], {
        'long_name': 'Time lag',
        'standard_name': 'time_lag',
        'units': 's',
        'sampling_rate': float(sampling_rate),
        'delta': float(delta),
        'npts': int(npts),
        'clip_lag': np.array(clip_lag if clip_lag is not None else []),
        'index_min': int(nmin),
        'index_max': int(nmax),
    })

    P, W, K = len(pairs), len(times), nmax - nmin
    dist = np.full(P, np.nan)
    if distance is not None:
        dist = np.array([float(distance[p]) if isinstance(distance, dict) else float(distance) for p in pairs])
    ds.variables['distance'] = Variable(('pair',), dist, {
        'long_name': 'receiver pair distance',
        'standard_name': 'receiver_pair_distance',
        'units': 'km',
        'description': 'absolute WGS84 geodetic distance',
        'relative_to_poi': 'n/a',
    }, encoding)
    ds.variables['status'] = Variable(('pair', 'time'), np.zeros((P, W), dtype=np.byte), {
        'long_name': 'processing status',
        'standard_name': 'processing_status',
        'units': '-',
        'valid_range': np.array([-1, 1], dtype=np.byte),
        'flag_values': np.array([-1, 0, 1], dtype=np.byte),
        'flag_meanings': 'missing_data not_processed processed',
    }, encoding)
    if hash_waveforms:
        ds.variables['hash'] = Variable(('pair', 'time'), np.array([['n/a'] * W] * P, dtype=object), {
            'long_name': 'pair preprocessed stream hash',
            'standard_name': 'pair_preprocessed_stream_hash',
            'units': '-',
        }, encoding)
    ds.variables['pair_offset'] = Variable(('pair', 'time'), np.zeros((P, W), dtype=np.float64), {
        'long_name': 'receiver pair start sample offset',
        'standard_name': 'receiver_pair_start_sample_offset',
        'units': 's',
        'description': 'offset = receiver[0].starttime - receiver[1].starttime',
    }, encoding)
    ds.variables['time_offset'] = Variable(('pair', 'time'), np.zeros((P, W), dtype=np.float64), {
        'long_name': 'first receiver start sample offset',
        'standard_name': 'first_receiver_start_sample_offset',
        'units': 's',
        'description': 'offset = receiver[0].starttime - time + window_length/2',
    }, encoding)
    ds.variables['cc'] = Variable(('pair', 'time', 'lag'), _zeros_maybe_pinned((P, W, K), dtype, pin_memory), {
        'long_name': 'Cross-correlation Estimate',
        'standard_name': 'cross-correlation_estimate',
        'units': '-',
        'add_offset': dtype(0.),
        'scale_factor': dtype(1.),
        'valid_range': np.array([-1., 1.], dtype=dtype),
        'normalize': np.byte(1),
        'bias_correct': np.byte(unbiased_cc),
        'unbiased': np.byte(0),
    }, encoding)
    # add metadata hash (core/init.py:323)
    from ..io.hasher import hash_dataset
    ds.attrs['sha256_hash_metadata'] = hash_dataset(ds, metadata_only=True)
    return ds
