#!/usr/bin/env python
"""Golden sha256 hashes produced by the REFERENCE's own hasher (run in the build container only).

``/root/reference/xcorr/util/hasher.py`` imports obspy and xarray at module level (both absent
here) but its dataset hashing touches only ``.name / .dims / .attrs / .dtype`` and ``np.nditer``.
So the file is executed unmodified with empty stand-in modules for the two imports, and
``hash_Dataset`` is applied to duck-typed views of a variable set built by ``xcorr_b200.core.init``
with fixed arguments.  ``tests/test_io.py`` rebuilds that set and requires
``xcorr_b200.io.hash_dataset`` to give the same digests.

    python tests/golden/make_golden_hash.py      # -> tests/golden/hash_golden.json
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("XCORR_REFERENCE", "/root/reference")


def load_reference_hasher():
    for name, attrs in (("obspy", ("Stream", "Trace")), ("xarray", ("Dataset", "DataArray"))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("xcorr_reference_hasher", os.path.join(REF, "xcorr/util/hasher.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class DuckArray:
    def __init__(self, name, var):
        self.name, self.dims, self.attrs = name, tuple(var.dims), var.attrs
        self._data = np.asarray(var.data)
        self.dtype = self._data.dtype

    def __array__(self, dtype=None, copy=None):
        return self._data


class DuckDataset:
    def __init__(self, ds):
        self.attrs = ds.attrs
        self.coords = {k: DuckArray(k, v) for k, v in ds.coords.items()}
        self._vars = {k: DuckArray(k, v) for k, v in ds.variables.items()}

    def __iter__(self):                       # xarray: iterating a Dataset yields its data variables
        return iter(self._vars)

    def __getitem__(self, k):
        return self.coords[k] if k in self.coords else self._vars[k]


def golden_dataset():
    """The fixed variable set both sides hash (also imported by tests/test_io.py)."""
    import pandas as pd
    import xcorr_b200.core as core
    ds = core.init(["NL.AAA..HHZ-NL.BBB..HHZ", "NL.AAA..HHZ-NL.CCC.00.HHZ"], pd.Timestamp("2020-01-01"),
                   pd.Timestamp("2020-01-01 01:00"),
                   preprocess={"HHZ": [("detrend", {"type": "demean"}),
                                       ("filter", {"type": "bandpass", "freqmin": 3.0, "freqmax": 12.0})]},
                   sampling_rate=10.0, window_length=600.0, window_overlap=0.5, clip_lag=(0.0, 20.0),
                   attrs=dict(title="t", institution="i", author="a", source="s"), pin_memory=False)
    ds.attrs["history"] = "Created @ golden"
    rng = np.random.default_rng(42)
    ds.variables["cc"].data[...] = rng.uniform(-1, 1, ds.variables["cc"].data.shape).astype(np.float32)
    ds.variables["status"].data[...] = 1
    ds.variables["pair_offset"].data[...] = 0.25
    return ds


if __name__ == "__main__":
    ref = load_reference_hasher()
    ds = golden_dataset()
    duck = DuckDataset(ds)
    out = {
        "metadata": ref.hash_Dataset(duck, metadata_only=True),
        "data": ref.hash_Dataset(duck, metadata_only=False),
        "to_json_sample": ref.to_json({"b": np.int8(1), "a": np.float32(0.5), "sha256_hash": "x", "scale_factor": 1.0}),
        "generated_by": "xcorr/util/hasher.py::hash_Dataset executed from /root/reference",
    }
    # the hashes must not depend on the stored hashes or the packing attributes (util/hasher.py:22-35)
    ds.attrs["sha256_hash"] = "something"
    ds.attrs["sha256_hash_metadata"] = out["metadata"]
    assert ref.hash_Dataset(DuckDataset(ds), metadata_only=False) == out["data"]
    with open(os.path.join(HERE, "hash_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
