"""Batched correlation plan: the B200 replacement for the reference's serial
pair x time loop (``core/process.py:82-165``).

One plan covers ``n_receivers`` signals per window, ``n_windows`` windows and a
list of receiver pairs.  Each receiver-window is transformed once
(:meth:`CorrelationPlan.forward`) and its spectrum reused by every pair it
appears in (:meth:`CorrelationPlan.correlate`), i.e. N transforms plus P
products per window instead of the reference's 2P.

torch is used for device memory and streams only; all arithmetic happens in
``libxcorr_b200.so`` through the C ABI.
"""
import ctypes

import numpy as np
import torch

from . import _lib

__all__ = ["CorrelationPlan", "all_pairs"]

_TORCH_DTYPE = {_lib.XCB_F32: torch.float32, _lib.XCB_F64: torch.float64}
_TORCH_CDTYPE = {_lib.XCB_F32: torch.complex64, _lib.XCB_F64: torch.complex128}


def all_pairs(n_receivers, auto=False):
    """All receiver pairs (i < j), optionally with the auto-correlations."""
    return [(i, j) for i in range(n_receivers) for j in range(i if auto else i + 1, n_receivers)]


def _xcb_dtype(dtype):
    if dtype in (torch.float32, np.float32, "float32", "f32"):
        return _lib.XCB_F32
    if dtype in (torch.float64, np.float64, "float64", "f64"):
        return _lib.XCB_F64
    d = np.dtype(dtype)
    if d == np.float32:
        return _lib.XCB_F32
    if d == np.float64:
        return _lib.XCB_F64
    raise TypeError(f"unsupported correlation dtype {dtype!r} (float32 / float64)")


class CorrelationPlan:
    """Plan + device buffers for ``cc`` over (pair, window) cells.

    Parameters
    ----------
    n : int
        Samples per window at the correlation rate (``lag.attrs['npts']``).
    n_receivers : int
        Signals per window.
    pairs : sequence of (i, j)
        Receiver index pairs; cell (i, j) is ``cc(x_i, x_j)``.
    n_windows : int
        Windows processed per call.
    dtype : float32 or float64
        ``dataset.cc.dtype`` of the reference (``core/init.py:30``).
    index_min, index_max : int
        Lag crop ``[index_min, index_max)`` (``lag.attrs``); default full.
    normalize, unbiased, stack : bool
        ``util/cc.py:76-77``; ``signal/unbias.py:85`` weights; sum over windows.
    """

    def __init__(self, n, n_receivers, pairs, n_windows=1, dtype=torch.float32,
                 index_min=0, index_max=0, normalize=True, unbiased=False,
                 stack=False, dims=None, max_work_bytes=0, device=None, n_blocks=0, block_len=0):
        if not torch.cuda.is_available():
            raise _lib.XcbError("xcorr_b200 needs a CUDA device; there is no CPU fallback")
        self._lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.xdtype = _xcb_dtype(dtype)
        self.dtype = _TORCH_DTYPE[self.xdtype]
        self.pairs = [(int(i), int(j)) for i, j in pairs]
        self._pairs_arr = (ctypes.c_int32 * (2 * len(self.pairs)))(*[v for p in self.pairs for v in p])
        d = _lib.xcb_desc()
        d.n = int(n)
        d.dtype = self.xdtype
        if dims:
            d.nd = len(dims)
            for a, v in enumerate(dims):
                d.dims[a] = int(v)
        d.n_receivers = int(n_receivers)
        d.n_windows = int(n_windows)
        d.n_pairs = len(self.pairs)
        d.normalize = int(bool(normalize))
        d.unbiased = int(bool(unbiased))
        d.stack = int(bool(stack))
        d.pairs = ctypes.cast(self._pairs_arr, ctypes.POINTER(ctypes.c_int32))
        d.lag_index_min = int(index_min)
        d.lag_index_max = int(index_max)
        d.max_work_bytes = int(max_work_bytes)
        d.n_blocks = int(n_blocks)
        d.block_len = int(block_len)
        self._plan = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.xcb_plan_create(ctypes.byref(self._plan), ctypes.byref(d)), "xcb_plan_create")
        info = _lib.xcb_plan_info()
        _lib.check(self._lib.xcb_plan_get_info(self._plan, ctypes.byref(info)), "xcb_plan_get_info")
        self.n = int(info.n)
        self.fft_len = int(info.fft_len)
        self.spectrum_elems = int(info.spectrum_elems)
        self.n_lags = int(info.n_lags)
        self.n_signals = int(info.n_signals)
        self.n_jobs = int(info.n_jobs)
        self.jobs_in_flight = int(info.jobs_in_flight)
        self.dims = tuple(int(info.dims[a]) for a in range(info.nd))
        self.n_blocks, self.block_len = int(info.n_blocks), int(info.block_len)
        self.n_groups = int(info.n_groups)
        self.n_receivers, self.n_windows = int(n_receivers), int(n_windows)
        self.stack = bool(stack)
        self.index_min = int(index_min)
        self.index_max = int(index_max) if (index_min or index_max) else 2 * self.n - 1
        self.workspace = torch.empty(int(info.workspace_bytes) + 256, dtype=torch.uint8, device=self.device)
        self.spec = torch.empty((self.n_signals, self.spectrum_elems), dtype=_TORCH_CDTYPE[self.xdtype],
                                device=self.device)
        self.sumsq = torch.empty(self.n_signals, dtype=torch.float64, device=self.device)

    def __del__(self):
        plan, self._plan = getattr(self, "_plan", None), None
        if plan:
            try:
                self._lib.xcb_plan_destroy(plan)
            except Exception:
                pass

    # -- device entry points -------------------------------------------------
    def forward(self, x):
        """Transform all receiver-windows.  ``x``: CUDA tensor of ``dtype`` shaped
        ``(n_receivers, n_windows, >=n)`` or ``(n_signals, >=n)``, last dim contiguous."""
        if x.device != self.device or x.dtype != self.dtype:
            raise TypeError(f"x must be a {self.dtype} tensor on {self.device}")
        x2 = x.reshape(-1, x.shape[-1])
        if x2.shape[0] != self.n_signals or x2.shape[1] < self.n or x2.stride(1) != 1:
            raise ValueError(f"x must hold {self.n_signals} signals of >= {self.n} contiguous samples")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.xcb_forward(self._plan, x2.data_ptr(), x2.stride(0), self.spec.data_ptr(),
                                         self.sumsq.data_ptr(), self.workspace.data_ptr(), stream), "xcb_forward")
        return self.spec

    def correlate(self, out=None):
        """Correlate every (pair, window).  Returns ``(n_pairs, n_windows, n_lags)``
        (``(n_pairs, n_lags)``, accumulated, when stacking)."""
        shape = (len(self.pairs), self.n_lags) if self.stack else (len(self.pairs), self.n_windows, self.n_lags)
        if out is None:
            out = (torch.zeros if self.stack else torch.empty)(shape, dtype=self.dtype, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != self.dtype or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous {self.dtype} tensor of shape {shape}")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.xcb_correlate(self._plan, self.spec.data_ptr(), self.sumsq.data_ptr(),
                                           out.data_ptr(), self.workspace.data_ptr(), stream), "xcb_correlate")
        return out

    def correlate_argmax(self, mask=None):
        """Arg-max and maximum of every (pair, window) correlation function without storing it (the last
        inverse pass reduces its tile; ``xcb_correlate_argmax``).  `mask` = ``(row_width, half_width,
        n_rows)`` restricts the lags to the 2-D window of a row-embedded image correlation.  Returns
        ``(index, value)`` shaped ``(n_pairs, n_windows)``; `index` counts from ``index_min``."""
        if self.stack:
            raise ValueError("correlate_argmax does not apply to stacking plans")
        idx = torch.empty(self.n_jobs, dtype=torch.int64, device=self.device)
        val = torch.empty(self.n_jobs, dtype=self.dtype, device=self.device)
        m = None
        if mask is not None:
            m = _lib.xcb_lag_mask(int(mask[0]), int(mask[1]), int(mask[2]))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.xcb_correlate_argmax(self._plan, self.spec.data_ptr(), self.sumsq.data_ptr(),
                                                  ctypes.byref(m) if m is not None else None, idx.data_ptr(),
                                                  val.data_ptr(), self.workspace.data_ptr(), stream),
                   "xcb_correlate_argmax")
        # jobs are window-major (job = w * n_pairs + p)
        shape = (self.n_windows, len(self.pairs))
        return idx.view(shape).t(), val.view(shape).t()

    def argmax(self, ccf):
        """First index of the maximum along the last axis, and the maximum."""
        rows = ccf.reshape(-1, ccf.shape[-1])
        if not rows.is_contiguous():
            rows = rows.contiguous()
        idx = torch.empty(rows.shape[0], dtype=torch.int64, device=ccf.device)
        val = torch.empty(rows.shape[0], dtype=ccf.dtype, device=ccf.device)
        with torch.cuda.device(ccf.device):          # xcb_argmax has no plan: it launches on the current device
            stream = torch.cuda.current_stream(ccf.device).cuda_stream
            _lib.check(self._lib.xcb_argmax(_xcb_dtype(ccf.dtype), rows.data_ptr(), rows.shape[0], rows.shape[1],
                                            idx.data_ptr(), val.data_ptr(), stream), "xcb_argmax")
        return idx.reshape(ccf.shape[:-1]), val.reshape(ccf.shape[:-1])

    # -- measurement aid ------------------------------------------------------
    def profile(self, enable=True):
        """Bracket every pass kernel with CUDA events (see ``xcb_profile_enable``)."""
        _lib.check(self._lib.xcb_profile_enable(self._plan, int(bool(enable))), "xcb_profile_enable")

    def profile_read(self):
        """``[(name, total_ms, launches)]`` per pass kind since the last read."""
        nk = ctypes.c_int32()
        ms = (ctypes.c_double * 8)()
        cnt = (ctypes.c_int64 * 8)()
        _lib.check(self._lib.xcb_profile_read(self._plan, ctypes.byref(nk), ms, cnt), "xcb_profile_read")
        nd = len(self.dims)
        names = [f"fwd_pass{a}_N{self.dims[a]}" for a in range(nd)] + \
                [f"inv_pass{a}_N{self.dims[a]}" for a in range(nd)]
        return [(names[k], float(ms[k]), int(cnt[k])) for k in range(nk.value)]

    def __call__(self, x, out=None):
        self.forward(x)
        return self.correlate(out)

    def graphed(self, x, out=None):
        """Capture ``forward(x)`` + ``correlate(out)`` into a CUDA graph on static buffers and return
        ``(replay, out)``: ``replay()`` re-runs the whole step with one launch.  Small windows (config 1: ten
        kernels of ~15 us) are bound by launch gaps, not by the kernels; the captured step removes them.  `x` and
        `out` must keep their addresses (refill `x` in place between replays)."""
        shape = (len(self.pairs), self.n_lags) if self.stack else (len(self.pairs), self.n_windows, self.n_lags)
        if out is None:
            out = (torch.zeros if self.stack else torch.empty)(shape, dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                 # warm-up outside the capture (lazy module loading)
                self.forward(x)
                self.correlate(out=out)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.forward(x)
                self.correlate(out=out)
        self._graph = graph
        return graph.replay, out
