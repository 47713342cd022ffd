r"""
:mod:`io.write` -- write / read the cc variable set (``xcorr/io/write.py:31-184``)

The reference writes NetCDF-4 (zlib, shuffle) through xarray.  When xarray with a
NetCDF-4 engine is importable the same call is made here
(``CCDataset.to_xarray().to_netcdf``); otherwise -- this image has neither netCDF4
nor h5netcdf -- the identical variables, dimensions and attributes go to a
NetCDF-3 (64-bit offset) file through :class:`scipy.io.netcdf_file`:
uncompressed, strings as character arrays, times as CF ``seconds since 1970-01-01``.
Either way the reference's protocol is kept: metadata hash verified / updated,
empty sets skipped unless `force_write`, data hash stored in ``sha256_hash``,
temporary file then atomic replace.
"""
import json
import os
import shutil
import warnings

import numpy as np
import pandas as pd

from ..core.init import CCDataset, Variable
from .hasher import hash_dataset

__all__ = ['write', 'read']

_EPOCH_UNITS = 'seconds since 1970-01-01 00:00:00'


def _netcdf4_engine():
    try:
        import xarray  # noqa: F401
    except ImportError:
        return None
    for eng in ('h5netcdf', 'netcdf4'):
        try:
            __import__('h5netcdf' if eng == 'h5netcdf' else 'netCDF4')
            return eng
        except ImportError:
            continue
    return None


def _attr_value(v):
    if isinstance(v, dict):
        return json.dumps(v)
    if isinstance(v, (bool, np.bool_)):
        return np.int8(v)
    if isinstance(v, (int, np.integer)):
        return np.int32(v) if -2**31 <= int(v) < 2**31 else np.float64(v)
    if isinstance(v, float):
        return np.float64(v)                 # scipy's writer would store a Python float as NC_FLOAT (32 bit)
    if isinstance(v, (list, tuple)):
        return np.asarray(v)
    if v is None:
        return 'None'
    return v


def _write_netcdf3(ds, path):
    from scipy.io import netcdf_file
    with netcdf_file(path, 'w', version=2) as f:
        for k, v in ds.attrs.items():
            setattr(f, k, _attr_value(v))
        items = list(ds.coords.items()) + list(ds.variables.items())
        for name, var in items:                                     # dimensions first
            data = np.asarray(var.data)
            for d, n in zip(var.dims, data.shape):
                if d not in f.dimensions:
                    f.createDimension(d, int(n))
        for name, var in items:
            data = np.asarray(var.data)
            attrs = dict(var.attrs)
            if data.dtype.kind in 'OUS':                            # strings -> char[n][strlen]
                strs = [str(s) for s in data.reshape(-1)]
                width = max(1, max(len(s) for s in strs))
                dim = f'{name}_strlen'
                f.createDimension(dim, width)
                v = f.createVariable(name, 'c', tuple(var.dims) + (dim,))
                v[:] = np.array([list(s.ljust(width)) for s in strs], dtype='S1').reshape(data.shape + (width,))
            elif data.dtype.kind == 'M':                            # datetime64 -> CF seconds
                v = f.createVariable(name, 'd', tuple(var.dims))
                v[:] = (data.astype('datetime64[ns]').astype(np.int64) / 1e9)
                attrs.setdefault('units', _EPOCH_UNITS)
                attrs.setdefault('calendar', 'proleptic_gregorian')
            else:
                if data.dtype == np.int64:
                    if data.size and (data.min() < -2**31 or data.max() >= 2**31):
                        raise OverflowError(f'variable {name!r}: int64 values do not fit NetCDF-3 (no 64-bit integers); '
                                            'write through xarray + netCDF4 instead')
                    data = data.astype(np.int32)
                code = {'float32': 'f', 'float64': 'd', 'int8': 'b', 'int16': 'h', 'int32': 'i'}[data.dtype.name]
                v = f.createVariable(name, code, tuple(var.dims))
                if data.ndim:
                    v[:] = data
                else:
                    v.assignValue(data)
            for k, a in attrs.items():
                setattr(v, k, _attr_value(a))


def write(data: CCDataset, path: str, force_write: bool = False, hash_data: bool = True, verb: int = 1, **kwargs):
    """Write the variable set to `path` through a temporary file (``io/write.py:31-184``).
    The metadata / data hashes are updated in ``data.attrs`` in place."""
    if not isinstance(data, CCDataset):
        raise TypeError('data should be a CCDataset')
    metadata_hash = hash_dataset(data, metadata_only=True)
    if data.attrs.get('sha256_hash_metadata') not in (None, metadata_hash) and verb > -1:
        warnings.warn('Data metadata sha256 hash is updated.', UserWarning)
    data.attrs['sha256_hash_metadata'] = metadata_hash
    if 'status' in data.variables and not force_write and not np.any(data.variables['status'].data == 1):
        warnings.warn('Dataset contains no data. No need to save it.', UserWarning)
        return None
    abspath, file = os.path.split(os.path.abspath(path))
    os.makedirs(abspath, exist_ok=True)
    tmp = os.path.join(abspath, '{f}.{t}'.format(f=file, t=int(pd.to_datetime('now').timestamp() * 1e3)))
    if hash_data:
        data.attrs['sha256_hash'] = hash_dataset(data, metadata_only=False)
    else:
        data.attrs.pop('sha256_hash', None)
    engine = _netcdf4_engine()
    if engine:
        xds = data.to_xarray()
        xds.to_netcdf(path=tmp, mode='w', format='netcdf4', compute=True, **{'engine': engine, **kwargs})
    else:
        _write_netcdf3(data, tmp)
    shutil.move(tmp, os.path.join(abspath, file))
    if verb > 0:
        print('Write data as "{}": Done.'.format(path))
    return os.path.join(abspath, file)


def _attr_back(v):
    """Inverse of :func:`_attr_value` as far as NetCDF-3 keeps the information: text, native-endian
    arrays, scalars as numpy scalars (JSON-serialised like the Python numbers they were)."""
    if isinstance(v, bytes):
        v = v.decode()
    if isinstance(v, str):
        return None if v == 'None' else v
    a = np.asarray(v)
    if a.dtype.kind in 'fiu' and not a.dtype.isnative:
        a = a.astype(a.dtype.newbyteorder('='))
    if a.ndim == 0 or a.size == 1 and not isinstance(v, np.ndarray):
        return a.reshape(-1)[0]
    return a


def read(path: str, fast: bool = False):
    """Read a file written by :func:`write` with the NetCDF-3 fallback back into a :class:`CCDataset`
    (``io/read.py:32-205``): the ``preprocess`` attribute becomes the operations dictionary again and its
    hash is verified, the stored ``sha256_hash`` is checked against the data (unless `fast`) and removed."""
    from scipy.io import netcdf_file
    from ..stream.process import operations_to_dict, check_operations_hash
    ds = CCDataset()
    with netcdf_file(path, 'r', mmap=False) as f:
        for k, v in f._attributes.items():
            ds.attrs[k] = _attr_back(v)
        for name, var in f.variables.items():
            attrs = {k: _attr_back(a) for k, a in var._attributes.items()}
            data = var.data.copy()
            if data.dtype.kind in 'fiu' and not data.dtype.isnative:
                data = data.astype(data.dtype.newbyteorder('='))     # NetCDF-3 is big-endian on disk
            dims = tuple(var.dimensions)
            if var.typecode() == 'c':
                data = np.array([b''.join(r).decode().rstrip() for r in data.reshape(-1, data.shape[-1])],
                                dtype=object).reshape(data.shape[:-1])
                dims = dims[:-1]
            elif attrs.get('units') == _EPOCH_UNITS:
                data = (np.round(data * 1e9)).astype(np.int64).astype('datetime64[ns]')
                attrs.pop('units'), attrs.pop('calendar', None)
            target = ds.coords if (len(dims) == 1 and dims[0] == name) else ds.variables
            target[name] = Variable(dims, data, attrs)
    if 'pair' in ds.coords and isinstance(ds.coords['pair'].attrs.get('preprocess'), str):
        ops = operations_to_dict(ds.coords['pair'].attrs['preprocess'])
        check_operations_hash(ops, raise_error=True)
        ds.coords['pair'].attrs['preprocess'] = ops
    stored = ds.attrs.pop('sha256_hash', None)          # io/read.py:184-185
    if stored is not None and not fast:
        if hash_dataset(ds, metadata_only=False) != stored:
            raise ValueError(f'{path}: sha256 data hash does not match the stored sha256_hash')
    if ds.attrs.get('sha256_hash_metadata') is not None:
        if hash_dataset(ds, metadata_only=True) != ds.attrs['sha256_hash_metadata']:
            warnings.warn('Dataset metadata sha256 hash is inconsistent.', UserWarning)
    return ds
