"""Host-side output: SDS-like path, sha256 hashes (serialisation of util/hasher.py) and the
writer's protocol (io/write.py:31-184) with the NetCDF-3 fallback of this image."""
import os
import warnings

import numpy as np
import pandas as pd
import pytest


def _dataset(status=1):
    import xcorr_b200.core as core
    ds = core.init(["NL.AAA..HHZ-NL.BBB..HHZ", "NL.AAA..HHZ-NL.CCC.00.HHZ"], pd.Timestamp("2020-01-01"),
                   pd.Timestamp("2020-01-01 01:00"), preprocess={"HHZ": [("detrend", {"type": "demean"})]},
                   sampling_rate=10.0, window_length=600.0, window_overlap=0.0, clip_lag=(0.0, 20.0),
                   attrs=dict(title="t", institution="i", author="a", source="s"))
    rng = np.random.default_rng(0)
    ds.variables["cc"].data[...] = rng.uniform(-1, 1, ds.variables["cc"].data.shape)
    ds.variables["status"].data[...] = status
    return ds


def test_ncfile_path():
    from xcorr_b200.io import ncfile
    p = ncfile("NL.AAA..HHZ-NL.BBB.00.HHZ", pd.Timestamp("2020-02-03"), root="/data")
    assert p == "/data/2020/NL.AAA..HHZ-NL.BBB.00.HHZ/NL.AAA..HHZ-NL.BBB.00.HHZ.2020.034.nc"
    assert ncfile(("NL.AAA..HHZ", "NL.BBB..HHZ"), pd.Timestamp("2021-12-31")).endswith(".2021.365.nc")
    with pytest.raises(TypeError):
        ncfile("NL.A*..HHZ-NL.BBB..HHZ", pd.Timestamp("2020-01-01"))
    with pytest.raises(TypeError):
        ncfile("NL.AAA..HHZ-NL.BBB..HHZ", "2020-01-01")


def test_hashes_follow_reference_serialisation():
    import hashlib
    from xcorr_b200.io import hash_dataset, hash_variable, to_json
    ds = _dataset()
    # hand-built: name, repr(dims), json(attrs) per coordinate then variable, sorted by name
    h = hashlib.sha256()
    for group in (ds.coords, ds.variables):
        for name in sorted(group):
            v = group[name]
            h.update(name.encode()); h.update(repr(tuple(v.dims)).encode()); h.update(to_json(v.attrs).encode())
    assert hash_dataset(ds) == h.hexdigest()
    assert to_json({"b": np.int8(1), "a": np.float32(0.5)}) == '{\n    "a":0.5,\n    "b":1\n}'
    full = hash_dataset(ds, metadata_only=False)
    assert full != hash_dataset(ds)
    ds.variables["cc"].data[0, 0, 0] += 1e-3
    assert hash_dataset(ds, metadata_only=False) != full and hash_dataset(ds) == h.hexdigest()
    assert hash_variable("cc", ds.variables["cc"]) == hash_variable("cc", ds.variables["cc"])


def test_write_protocol_and_round_trip(tmp_path):
    from xcorr_b200.io import write, read, ncfile, hash_dataset
    ds = _dataset()
    path = os.path.join(tmp_path, ncfile(str(ds.coords["pair"].data[0]), pd.Timestamp(ds.coords["time"].data[0])))
    out = write(ds, path, verb=0)
    assert out == path and os.path.exists(path) and os.listdir(os.path.dirname(path)) == [os.path.basename(path)]
    assert ds.attrs["sha256_hash_metadata"] == hash_dataset(ds)
    assert len(ds.attrs["sha256_hash"]) == 64
    back = read(path)
    # read() verifies the stored data hash and drops it (io/read.py:184-185)
    assert back.attrs["Conventions"] == "CF-1.9" and "sha256_hash" not in back.attrs
    assert hash_dataset(back, metadata_only=False) == ds.attrs["sha256_hash"]
    for name in ("cc", "status", "pair_offset", "time_offset", "distance"):
        a, b = ds.variables[name], back.variables[name]
        assert tuple(a.dims) == tuple(b.dims) and np.array_equal(a.data, b.data, equal_nan=True)
        assert b.data.dtype == a.data.dtype
    assert list(back.coords["pair"].data) == list(ds.coords["pair"].data)
    assert np.array_equal(back.coords["time"].data, ds.coords["time"].data.astype("datetime64[ns]"))
    assert np.array_equal(back.coords["lag"].data, ds.coords["lag"].data)
    assert back.coords["lag"].attrs["index_min"] == ds.coords["lag"].attrs["index_min"]
    # an empty set is not written
    empty = _dataset(status=0)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert write(empty, os.path.join(tmp_path, "empty.nc"), verb=0) is None
    assert any("no data" in str(x.message) for x in w) and not os.path.exists(os.path.join(tmp_path, "empty.nc"))
    assert write(empty, os.path.join(tmp_path, "empty.nc"), force_write=True, verb=0)


def test_hashes_equal_the_reference_hasher_golden():
    """tests/golden/hash_golden.json was produced by executing the reference's util/hasher.py
    (hash_Dataset) on the same variable set (tests/golden/make_golden_hash.py)."""
    import json
    from tests.golden.make_golden_hash import golden_dataset
    from xcorr_b200.io import hash_dataset, to_json
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hash_golden.json")))
    ds = golden_dataset()
    assert hash_dataset(ds, metadata_only=True) == gold["metadata"]
    assert hash_dataset(ds, metadata_only=False) == gold["data"]
    assert to_json({"b": np.int8(1), "a": np.float32(0.5), "sha256_hash": "x", "scale_factor": 1.0}) == gold["to_json_sample"]
    # util/hasher.py:22-35: stored hashes and packing attributes never enter a hash
    ds.attrs["sha256_hash"] = "anything"
    ds.attrs["sha256_hash_metadata"] = gold["metadata"]
    assert hash_dataset(ds, metadata_only=False) == gold["data"]


def test_stored_hashes_survive_write_read_and_rewrite(tmp_path):
    """the stored sha256_hash must equal the hash recomputed after write() and after read(),
    writing twice must store the same hash, and read() must give back a set process() can resume."""
    from xcorr_b200.io import write, read, hash_dataset
    ds = _dataset()
    p1, p2 = os.path.join(tmp_path, "a.nc"), os.path.join(tmp_path, "b.nc")
    write(ds, p1, verb=0)
    h_data, h_meta = ds.attrs["sha256_hash"], ds.attrs["sha256_hash_metadata"]
    assert hash_dataset(ds, metadata_only=False) == h_data and hash_dataset(ds) == h_meta
    with warnings.catch_warnings():
        warnings.simplefilter("error")                       # no "metadata hash is updated" on an unchanged set
        write(ds, p2, verb=0)
    assert ds.attrs["sha256_hash"] == h_data and ds.attrs["sha256_hash_metadata"] == h_meta
    with warnings.catch_warnings():
        warnings.simplefilter("error")                       # metadata hash consistent after the round trip
        back = read(p1)
    assert "sha256_hash" not in back.attrs                    # io/read.py:184-185
    assert hash_dataset(back, metadata_only=False) == h_data
    assert hash_dataset(back) == h_meta == back.attrs["sha256_hash_metadata"]
    ops = back.coords["pair"].attrs["preprocess"]
    assert isinstance(ops, dict) and "HHZ" in ops and [o[0] for o in ops["HHZ"]] == ["detrend"]
    assert back.coords["time"].attrs["closed"] == ds.coords["time"].attrs["closed"]
    # a corrupted file is detected by the data hash
    raw = bytearray(open(p1, "rb").read())
    raw[-40] ^= 0xFF
    open(p1, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        read(p1)
    # int64 data that does not fit NetCDF-3 raises instead of being truncated
    ds.variables["big"] = type(ds.variables["cc"])(("pair",), np.array([2**40, 1], dtype=np.int64))
    with pytest.raises(OverflowError):
        write(ds, os.path.join(tmp_path, "c.nc"), verb=0)
