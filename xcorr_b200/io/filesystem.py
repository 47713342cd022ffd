r"""
:mod:`io.filesystem` -- output paths (``xcorr/io/filesystem.py:28-97``)
"""
import os
import re

import pandas as pd

__all__ = ['ncfile']

_seed = re.compile(r'^[A-Z0-9]{1,2}\.[A-Z0-9]{1,5}\.[A-Z0-9]{0,2}\.[A-Z0-9]{3}$')
_pair_type_error = ('``pair`` should be of type `str` or of type `tuple` with length 2 specifying '
                    'the receiver SEED-ids without wildcards!')


def ncfile(pair, time: pd.Timestamp, root: str = None, verify_receiver: bool = True):
    r"""``{root}/{YYYY}/{pair}/{pair}.{YYYY}.{DOY}.nc`` (``io/filesystem.py:83-97``)."""
    if verify_receiver:
        if isinstance(pair, str):
            recs = pair.split('-')
        elif isinstance(pair, tuple) and len(pair) == 2:
            recs = list(pair)
        else:
            raise TypeError(_pair_type_error)
        if len(recs) != 2 or not all(isinstance(r, str) and _seed.match(r) for r in recs):
            raise TypeError(_pair_type_error)
        pair = '-'.join(recs)
    if not isinstance(time, pd.Timestamp):
        raise TypeError('``time`` should be of type `pd.Timestamp`!')
    if root and not isinstance(root, str):
        raise TypeError('``root`` should be of type `str`!')
    year = f'{time.year:04d}'
    root = os.path.join(root, year, pair) if root else os.path.join(year, pair)
    return os.path.join(root, '{p}.{y:04d}.{d:03d}.nc'.format(p=pair, y=time.year, d=time.dayofyear))
