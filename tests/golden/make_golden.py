#!/usr/bin/env python
"""Generate golden vectors by EXECUTING the reference's own ``xcorr/util/cc.py``.

Run in the build container (``/root/reference`` must exist; it does not travel
to the GPU box).  The reference module only needs numpy, so it is loaded by
path -- ``import xcorr`` itself fails here (xarray/obspy absent).  ``np.int``
(removed from numpy) is patched to ``int`` for ``weight()`` only, which is the
minimal change that lets ``util/cc.py:137`` execute; nothing else is altered.

    python tests/golden/make_golden.py

writes ``tests/golden/cc_golden.npz`` (a few hundred kB).  Cases cover
odd/even/tiny lengths, float32/float64, normalize on/off, pad on/off, zero
vectors, delta pulses and the SURVEY-8d synthetic receivers.
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/xcorr/util/cc.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("refcc", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.cc_oracle import synthetic_receivers

    out = {}
    meta = []

    def add(name, x, y, **kw):
        with np.errstate(all="ignore"):
            r = ref.cc(x, y, unbiased=False, **kw)
        out[name + "_x"] = x
        out[name + "_y"] = y
        out[name + "_r"] = r
        meta.append((name, kw))

    rng = np.random.default_rng(20261018)
    for n in (1, 2, 3, 4, 5, 8, 17, 64, 100, 257, 1000, 4096):
        x = rng.standard_normal(n)
        y = rng.standard_normal(n)
        add(f"rand_f64_n{n}", x, y, normalize=True, pad=True)
    for n in (5, 64, 1000):
        x = rng.standard_normal(n)
        y = rng.standard_normal(n)
        add(f"rand_f64_nonorm_n{n}", x, y, normalize=False, pad=True)
        add(f"rand_f64_nopad_n{n}", x, y, normalize=True, pad=False)
        add(f"rand_f32_n{n}", x.astype(np.float32), y.astype(np.float32),
            normalize=True, pad=True)
        # float64 inputs stored as float32 (what core/process.py:155-162 does)
        add(f"rand_f64_to_f32_n{n}", x, y, normalize=True, pad=True,
            dtype=np.float32)
    # shifted copy: argmax must be -7 (sign convention of util/cc.py:75-78)
    x = rng.standard_normal(1000)
    add("shift7", x, np.roll(x, 7), normalize=True, pad=True)
    # delta pulses
    a = np.zeros(50)
    b = np.zeros(50)
    a[30] = 2.0
    b[10] = -3.0
    add("delta", a, b, normalize=True, pad=True)
    # zero vector: NaN policy of util/cc.py:77 (0/0)
    add("zero", np.zeros(16), rng.standard_normal(16), normalize=True, pad=True)
    # SURVEY 8d synthetic receivers
    rec = synthetic_receivers(3, 5000, seed=1000, delay=137)
    add("synth_01", rec[0], rec[1], normalize=True, pad=True)
    add("synth_02", rec[0], rec[2], normalize=True, pad=True)
    add("synth_12_f32", rec[1].astype(np.float32), rec[2].astype(np.float32),
        normalize=True, pad=True)

    # lag(), weight() (np.int patched), shift helpers
    for n, d in ((5, 0.5), (6, 0.02), (1000, 0.004)):
        out[f"lag_n{n}_pad"] = ref.lag(n, d, pad=True)
        out[f"lag_n{n}_nopad"] = ref.lag(n, d, pad=False)
    np.int = int  # numpy >= 1.24 removed the alias used at util/cc.py:137
    try:
        for n in (5, 6, 1000):
            out[f"weight_n{n}"] = ref.weight(n, pad=True)
            out[f"weight_n{n}_clip"] = ref.weight(n, pad=True, clip=3.0)
    finally:
        del np.int
    r = out["shift7_r"]
    out["shift7_extract"] = np.array([ref.extract_shift(r),
                                      ref.extract_shift(r, 0.02)])
    sm = ref.extract_shift_and_max(r, 0.02)
    out["shift7_extract_max"] = np.array([sm[0], sm[1]])
    out["shift7_compute"] = np.array([
        ref.compute_shift(x, np.roll(x, 7), unbiased=False),
        ref.compute_shift_and_max(x, np.roll(x, 7), 0.02, unbiased=False)[1],
    ])
    out["_cases"] = np.array([m[0] for m in meta])
    out["_kwargs"] = np.array([
        json.dumps({k: (np.dtype(v).name if k == "dtype" else v)
                    for k, v in m[1].items()}) for m in meta])
    path = os.path.join(HERE, "cc_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(meta), "cc cases")


if __name__ == "__main__":
    main()
