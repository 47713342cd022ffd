"""ctypes binding of ``libxcorr_b200.so`` (C ABI in ``include/xcorr_b200.h``).

The shared library is built in-tree by ``__graft_entry__.build()`` (``make`` in
``xcorr_b200/csrc``).  There is no CPU fallback: if the library is missing the
import of any compute entry point fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# XCB_LIB: another build of the same library (A/B runs of kernel variants)
LIB_PATH = os.environ.get("XCB_LIB") or os.path.join(_HERE, "libxcorr_b200.so")

XCB_F32, XCB_F64 = 0, 1
XCB_OK = 0


class XcbError(RuntimeError):
    """A libxcorr_b200 call returned a non-zero status."""


class xcb_desc(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int64),
        ("dtype", ctypes.c_int32),
        ("nd", ctypes.c_int32),
        ("dims", ctypes.c_int32 * 4),
        ("n_receivers", ctypes.c_int32),
        ("n_windows", ctypes.c_int32),
        ("n_pairs", ctypes.c_int32),
        ("recv_stride", ctypes.c_int32),
        ("win_stride", ctypes.c_int32),
        ("normalize", ctypes.c_int32),
        ("unbiased", ctypes.c_int32),
        ("stack", ctypes.c_int32),
        ("pairs", ctypes.POINTER(ctypes.c_int32)),
        ("lag_index_min", ctypes.c_int64),
        ("lag_index_max", ctypes.c_int64),
        ("max_work_bytes", ctypes.c_int64),
        ("n_blocks", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("block_len", ctypes.c_int64),
    ]


class xcb_plan_info(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int64),
        ("fft_len", ctypes.c_int64),
        ("spectrum_elems", ctypes.c_int64),
        ("n_lags", ctypes.c_int64),
        ("n_signals", ctypes.c_int64),
        ("n_jobs", ctypes.c_int64),
        ("jobs_in_flight", ctypes.c_int64),
        ("workspace_bytes", ctypes.c_size_t),
        ("nd", ctypes.c_int32),
        ("dims", ctypes.c_int32 * 4),
        ("dtype", ctypes.c_int32),
        ("n_blocks", ctypes.c_int32),
        ("block_len", ctypes.c_int64),
        ("n_groups", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class xcb_lag_mask(ctypes.Structure):
    _fields_ = [("row_width", ctypes.c_int64), ("half_width", ctypes.c_int64), ("n_rows", ctypes.c_int64)]


class xcb_pointwise(ctypes.Structure):
    _fields_ = [
        ("detrend", ctypes.c_int32),
        ("taper_left", ctypes.c_int32),
        ("taper_right", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("ramp_left", ctypes.c_void_p),
        ("ramp_right", ctypes.c_void_p),
    ]


class xcb_sos(ctypes.Structure):
    _fields_ = [
        ("n_sections", ctypes.c_int32),
        ("decimate", ctypes.c_int32),
        ("reverse", ctypes.c_int32),
        ("chunk", ctypes.c_int32),
        ("coef", (ctypes.c_double * 5) * 12),
        ("carry", ctypes.POINTER(ctypes.c_double)),
        ("carry_group", ctypes.POINTER(ctypes.c_double)),
        ("init_state", ctypes.POINTER(ctypes.c_double)),
        ("zin_response", ctypes.POINTER(ctypes.c_double)),
        ("carry_window", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


# every symbol include/xcorr_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
SYMBOLS = {
    "xcb_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "xcb_last_cuda_error": (ctypes.c_char_p, []),
    "xcb_version": (ctypes.c_int, []),
    "xcb_choose_transform": (ctypes.c_int, [_i64, ctypes.POINTER(_i32), ctypes.POINTER(_i32)]),
    "xcb_plan_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.POINTER(xcb_desc)]),
    "xcb_plan_destroy": (ctypes.c_int, [_vp]),
    "xcb_plan_get_info": (ctypes.c_int, [_vp, ctypes.POINTER(xcb_plan_info)]),
    "xcb_forward": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "xcb_correlate": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "xcb_correlate_argmax": (ctypes.c_int, [_vp, _vp, _vp, ctypes.POINTER(xcb_lag_mask), _vp, _vp, _vp, _vp]),
    "xcb_argmax": (ctypes.c_int, [_i32, _vp, _i64, _i64, _vp, _vp, _vp]),
    "xcb_stats_workspace_bytes": (ctypes.c_int, [_i64, _i32, ctypes.POINTER(ctypes.c_size_t)]),
    "xcb_stats": (ctypes.c_int, [_i32, _vp, _i64, _i64, _i32, _vp, _vp, _vp]),
    "xcb_pointwise_apply": (ctypes.c_int, [_i32, _vp, _i64, _i32, _vp, _i64, _i64, _i32,
                                           ctypes.POINTER(xcb_pointwise), _vp, _vp]),
    "xcb_sosfilt_workspace_bytes": (ctypes.c_int, [_i64, _i32, ctypes.POINTER(xcb_sos),
                                                   ctypes.POINTER(ctypes.c_size_t)]),
    "xcb_sosfilt": (ctypes.c_int, [_i32, _vp, _i64, _i32, _vp, _i64, _i64, _i32, ctypes.POINTER(xcb_sos),
                                   ctypes.POINTER(xcb_pointwise), _vp, _vp, _vp]),
    "xcb_upfirdn": (ctypes.c_int, [_i32, _vp, _i64, _i64, _i32, _vp, _i64, _i64, _i32, _vp, _i32, _i32, _i32, _i64,
                                   ctypes.POINTER(xcb_pointwise), _vp, _vp]),
    "xcb_profile_enable": (ctypes.c_int, [_vp, _i32]),
    "xcb_profile_read": (ctypes.c_int, [_vp, ctypes.POINTER(_i32), ctypes.POINTER(ctypes.c_double),
                                        ctypes.POINTER(_i64)]),
    "xcb_fft_good_length": (ctypes.c_int, [_i64, ctypes.POINTER(_i64)]),
    "xcb_fft_create": (ctypes.c_int, [ctypes.POINTER(_vp), _i32, _i64, _i32]),
    "xcb_fft_destroy": (ctypes.c_int, [_vp]),
    "xcb_fft_forward": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "xcb_fft_inverse_conj": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i64, _i64, _vp]),
    "xcb_dft_bluestein": (ctypes.c_int, [_vp, _i64, _i32, _vp, _i32, _i64, _i64, _i32, _vp, _i32, _i64, _i64,
                                         ctypes.c_double, _vp, _i32, _vp, _vp]),
    "xcb_phase_ramp": (ctypes.c_int, [_i32, _vp, _i64, _i64, _i32, _vp, ctypes.c_double, _vp]),
    "xcb_resample_spectrum": (ctypes.c_int, [_i32, _vp, _i64, _vp, _i64, _i64, _i64, ctypes.c_double,
                                             ctypes.c_double, _i32, _i32, _vp]),
    "xcb_lanczos": (ctypes.c_int, [_i32, _vp, _i64, _i64, _vp, _i64, _i64, ctypes.c_double, ctypes.c_double, _i32,
                                   _i32, _vp]),
    "xcb_window_stats": (ctypes.c_int, [_i32, _vp, _i64, _i32, _i64, _i64, _i64, _i64, _vp, _vp]),
    "xcb_spectrogram": (ctypes.c_int, [_i32, _vp, _i64, _i64, _i32, _vp, _i32, _i32, ctypes.c_double, _i32, _vp, _vp]),
    "xcb_running_rms": (ctypes.c_int, [_i32, _vp, _i64, _i64, _i32, _i64, _i64, _i64, _vp, _vp]),
    "xcb_normalize_rows": (ctypes.c_int, [_i32, _vp, _i64, _i64, _i32, _vp, ctypes.c_double, _vp]),
    "xcb_l2_normalize_rows": (ctypes.c_int, [_i32, _vp, _i64, _i64, _i32, _vp, _vp]),
    "xcb_cc_host": (ctypes.c_int, [_i32, _i64, _vp, _vp, _i32, _i32, _i64, _i64, _vp]),
    "xcb_cc_host_clear_cache": (ctypes.c_int, []),
}

_lib = None


def load():
    """Load the shared library (once) and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XcbError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C xcorr_b200/csrc`).  xcorr_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what=""):
    if status != XCB_OK:
        lib = load()
        msg = lib.xcb_strerror(status).decode()
        if status == -3:
            msg += ": " + lib.xcb_last_cuda_error().decode()
        raise XcbError(f"{what or 'libxcorr_b200'} failed ({status}): {msg}")


def choose_transform(n):
    """Pass lengths the planner picks for ``n`` samples (host only)."""
    dims = (ctypes.c_int32 * 4)()
    nd = ctypes.c_int32()
    check(load().xcb_choose_transform(int(n), dims, ctypes.byref(nd)), "xcb_choose_transform")
    return tuple(int(dims[i]) for i in range(nd.value))
