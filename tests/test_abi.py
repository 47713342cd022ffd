"""The C-ABI library loads and exports every symbol include/xcorr_b200.h
declares (no compute calls -- no GPU here), and the host-only planner entry
works."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from xcorr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "xcorr_b200", "csrc"), "-j", "8"])
    return _lib


def test_header_symbols_are_exported(built):
    hdr = open(os.path.join(ROOT, "include", "xcorr_b200.h")).read()
    declared = set(re.findall(r"XCB_API[^;(]*?\b(xcb_\w+)\s*\(", hdr))
    assert declared and declared == set(built.SYMBOLS)
    lib = built.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.xcb_version() == 100
    assert lib.xcb_strerror(0) == b"ok" and b"CUDA" in lib.xcb_strerror(-3)


def test_planner_is_host_only(built):
    assert built.choose_transform(5) == (64,)
    d = built.choose_transform(21_600_000)
    p = 1
    for v in d:
        p *= v
    assert p == 21_600_000 and len(d) == 3


def test_no_cpu_fallback(built):
    import numpy as np
    import torch
    import xcorr_b200 as xb
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(built.XcbError):
        xb.cc(np.ones(8), np.ones(8))
    with pytest.raises(built.XcbError):
        xb.CorrelationPlan(100, 2, [(0, 1)])


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "xcorr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)


def test_host_helpers_match_oracle(golden):
    import numpy as np
    import xcorr_b200 as xb
    for n, d in ((5, 0.5), (6, 0.02), (1000, 0.004)):
        assert np.array_equal(xb.lag(n, d, True), golden[f"lag_n{n}_pad"])
        assert np.array_equal(xb.lag(n, d, False), golden[f"lag_n{n}_nopad"])
        assert np.array_equal(xb.weight(n)[:-2], golden[f"weight_n{n}"])
        assert np.array_equal(xb.weight(n, clip=3.0)[:-2], golden[f"weight_n{n}_clip"])
    r = golden["shift7_r"]
    assert xb.extract_shift(r) == -7
    assert xb.extract_shift_and_max(r, 0.02) == tuple(golden["shift7_extract_max"])
