"""ctypes access to the host thread-emulator of the CUDA pass bodies
(tests/host_emu/emu.cpp).  TEST HARNESS ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emu", "emu.cpp")
LIB = os.path.join(HERE, "host_emu", "libxcb_emu.so")
CSRC = os.path.join(HERE, "..", "xcorr_b200", "csrc")


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in
                    ("xcb_passes.h", "xcb_plan.h", "xcb_types.h", "xcb_radix_gen.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def choose_dims(n):
    d = (ctypes.c_int * 4)()
    nd = ctypes.c_int()
    rc = lib().emu_choose_dims(ctypes.c_int64(n), d, ctypes.byref(nd))
    assert rc == 0
    return tuple(d[i] for i in range(nd.value))


def choose_blocks(n, n_lags, n_recv, n_pairs):
    """(n_blocks, dims, block_len) of the automatic planner (n_blocks == 1: classic transform)."""
    d = (ctypes.c_int * 4)()
    nd = ctypes.c_int()
    bl = ctypes.c_int64()
    nb = lib().emu_choose_blocks(ctypes.c_int64(n), ctypes.c_int64(n_lags), n_recv, n_pairs, d, ctypes.byref(nd),
                                 ctypes.byref(bl))
    return nb, tuple(d[i] for i in range(nd.value)) if nb > 1 else choose_dims(n), bl.value


def emu_cc(x, pairs, dims=None, i0=None, i1=None, normalize=True, unbiased=False,
           want_spec=False, n_stack=1, n_blocks=1, block_len=0):
    """x: (S, n) float32/float64.  Returns (ccf[P, K], sumsq[S], spec or None)."""
    x = np.ascontiguousarray(x)
    S, n = x.shape
    dtype = {np.dtype(np.float32): 0, np.dtype(np.float64): 1}[x.dtype]
    auto_dims = dims is None
    if dims is None:
        dims = choose_dims(n)
    M = int(np.prod(dims))
    i0 = 0 if i0 is None else i0
    i1 = 2 * n - 1 if i1 is None else i1
    pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
    out = np.full((len(pairs), i1 - i0), np.nan, dtype=x.dtype)
    sumsq = np.zeros(S)
    spec = np.zeros((S, M), dtype=np.complex64 if dtype == 0 else np.complex128) if want_spec else None
    d = (ctypes.c_int * 4)(*(list(dims) + [0] * (4 - len(dims))))
    nd = 0 if (auto_dims and n_blocks < 0) else len(dims)
    rc = lib().emu_cc(dtype, ctypes.c_int64(n), d, nd, x.ctypes.data_as(ctypes.c_void_p),
                      ctypes.c_int64(x.strides[0] // x.itemsize), S,
                      pairs.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(pairs),
                      ctypes.c_int64(i0), ctypes.c_int64(i1), int(normalize), int(unbiased),
                      out.ctypes.data_as(ctypes.c_void_p),
                      sumsq.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                      spec.ctypes.data_as(ctypes.c_void_p) if want_spec else None, int(n_stack),
                      int(n_blocks), ctypes.c_int64(block_len))
    assert rc >= 1, rc           # returns the number of blocks used
    emu_cc.last_blocks = rc
    return out, sumsq, spec
