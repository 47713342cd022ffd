r"""
:mod:`io.hasher` -- sha256 hashes of the variable set (``xcorr/util/hasher.py:38-319``)

Same serialisation as the reference: every coordinate and variable (sorted by
name) contributes its name, ``repr(dims)`` and its attributes as JSON
(``separators=(',', ':'), sort_keys=True, indent=4``); with ``metadata_only=False``
the global attributes and the raw bytes of every element are hashed too.
"""
import hashlib
import json

import numpy as np

__all__ = ['to_json', 'hash_variable', 'hash_dataset']

_enc = 'utf-8'

# util/hasher.py:22-35: attributes that never enter a hash -- the hashes themselves (else the data hash
# would depend on the previous one) and the packing attributes xarray moves into `encoding` on read
_ignore_keys = ('sha256_hash', 'sha256_hash_metadata', 'add_offset', 'scale_factor')


def _to_serializable(obj):
    if isinstance(obj, (np.int8, np.int16, np.int32, np.int64)):
        return int(obj)
    if isinstance(obj, (np.float32, np.float64)):
        return float(obj)
    return repr(obj)


def to_json(obj):
    """``util/hasher.py:30-48``: JSON of a dictionary without the ``_ignore_keys``."""
    if isinstance(obj, dict):
        obj = {k: v for k, v in obj.items() if k not in _ignore_keys}
    return json.dumps(obj, separators=(',', ':'), sort_keys=True, indent=4, skipkeys=False, default=_to_serializable)


def hash_variable(name, var, metadata_only=True, hashlib_obj=None):
    """``hash_DataArray`` (``util/hasher.py:273-319``) for a :class:`xcorr_b200.core.Variable`."""
    h = hashlib_obj or hashlib.sha256()
    h.update(name.encode(_enc))
    h.update(repr(tuple(var.dims)).encode(_enc))
    h.update(to_json(var.attrs).encode(_enc))
    if not metadata_only:
        data = np.asarray(var.data)
        if data.dtype == np.dtype(object) or data.dtype.kind in 'US':
            for d in data.reshape(-1):
                h.update(str(d).encode(_enc))
        else:
            # np.nditer(...).tobytes() element by element == the C-order buffer
            h.update(np.ascontiguousarray(data).tobytes())
    return None if hashlib_obj else h.hexdigest()


def hash_dataset(ds, metadata_only=True, hashlib_obj=None):
    """``hash_Dataset`` (``util/hasher.py:218-270``) for a :class:`xcorr_b200.core.CCDataset`."""
    h = hashlib_obj or hashlib.sha256()
    if not metadata_only:
        h.update(to_json(ds.attrs).encode(_enc))
    for coord in sorted(ds.coords):
        hash_variable(coord, ds.coords[coord], metadata_only, h)
    for var in sorted(ds.variables):
        hash_variable(var, ds.variables[var], metadata_only, h)
    return None if hashlib_obj else h.hexdigest()
