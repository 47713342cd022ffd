"""CF-1.9 variable layout of core.init (core/init.py:121-325), host only."""
import numpy as np
import pandas as pd
import pytest

from xcorr_b200.core import init
from oracle import cc_oracle as orc


def test_layout_matches_reference_contract():
    ops = {"EDH": [("detrend", {"type": "demean"}), ("decimate", {"factor": 5})], "junk": 1}
    ds = init("IM.H10N1..EDH-IM.H03S1..EDH", pd.Timestamp("2015-01-15"), pd.Timestamp("2015-01-16"), ops,
              dict(institution="x", author="y", source="z", title="Monowai"), 50.0,
              window_length=86400.0, window_overlap=21 / 24, clip_lag=(0.0, 32400.0), unbiased_cc=True)
    assert set(ds.coords) == {"pair", "time", "lag"}
    assert set(ds.variables) == {"distance", "status", "pair_offset", "time_offset", "cc"}
    assert ds.cc.dims == ("pair", "time", "lag") and ds.cc.data.dtype == np.float32
    assert ds.status.data.dtype == np.int8 and ds.status.dims == ("pair", "time")
    assert len(ds.time.data) == 8 and ds.time.attrs["window_overlap"] == 21 / 24       # every 3 h
    la = ds.lag.attrs
    assert (la["npts"], la["index_min"], la["index_max"]) == (4_320_000, 4_319_999, 5_939_999)
    assert (la["index_min"], la["index_max"]) == orc.lag_clip_indices(4_320_000, 0.02, (0.0, 32400.0))
    assert ds.lag.data[0] == 0.0 and len(ds.lag.data) == 1_620_000
    assert ds.cc.attrs["bias_correct"] == 1 and ds.cc.attrs["unbiased"] == 0 and ds.cc.attrs["normalize"] == 1
    for v in ds.variables.values():
        assert v.encoding == dict(zlib=True, complevel=9, shuffle=True)
    assert ds.attrs["Conventions"] == "CF-1.9" and ds.attrs["title"].startswith("Monowai Cross-correlations - 2015.015")
    pre = ds.pair.attrs["preprocess"]
    assert set(pre) == {"EDH", "sha256_hash"}          # non-channel keys filtered, hash added
    with pytest.raises(KeyError):
        init("a-b", pd.Timestamp("2015-01-15"), pd.Timestamp("2015-01-16"), ops, dict(author="y"), 50.0)
    with pytest.raises(TypeError):
        init("a-b", pd.Timestamp("2015-01-15"), pd.Timestamp("2015-01-16"), ops,
             dict(institution="x", author="y", source="z"), 50.0, clip_lag=(0, 10))
    full = init(["a.b..EDH-c.d..EDH", "a.b..EDH-e.f..EDH"], pd.Timestamp("2015-01-15"), pd.Timestamp("2015-01-15 01:00"),
                ops, dict(institution="x", author="y", source="z"), 50.0, window_length=600.0, window_overlap=0.0)
    assert full.cc.data.shape == (2, 6, 2 * 30000 - 1)
