"""scipy restatement of the obspy ``Trace`` operations the reference's op-lists
name (TEST INFRASTRUCTURE; **parity unpinned**).

The reference applies preprocessing by ``eval('waveforms.<op>(**params)')`` on
obspy objects (``xcorr/stream/process.py:199-297``); the arithmetic lives in
obspy (``obspy>=1.2``, ``setup.cfg:41``), which is neither installed here nor
vendored under ``/root/reference``.  This module restates the published obspy
1.2-1.4 algorithms on numpy/scipy, one function per
operation, float64 throughout like obspy.  It is pinned only by analytic
known-answer tests (``tests/test_preprocess_oracle.py``) and by the reference's
call sites (``examples/core/process.py:61-155``); should an obspy install become
available, regenerate fixtures from it.

Not imported by the product package ``xcorr_b200``.
"""
import numpy as np
from scipy import signal

__all__ = ["detrend", "cosine_taper", "taper_window", "taper", "butter_sos", "filter_",
           "cheby2_lowpass_sos", "decimate", "trim", "trim_time", "resample", "interpolate_lanczos", "running_rms", "normalize", "process"]


def detrend(x, type="simple"):
    """obspy ``Trace.detrend``: 'demean'/'constant' -> x - mean;
    'linear' -> least-squares line (scipy.signal.detrend); 'simple' -> line
    through the first and last sample."""
    x = np.asarray(x, dtype=np.float64)
    if type in ("demean", "constant"):
        return signal.detrend(x, type="constant")
    if type == "linear":
        return signal.detrend(x, type="linear")
    if type == "simple":
        n = len(x)
        return x - (x[0] + (x[-1] - x[0]) * np.arange(n) / float(n - 1))
    raise ValueError(f"unsupported detrend type {type!r}")


def cosine_taper(npts, p=0.1):
    """obspy.signal.invsim.cosine_taper (no frequency arguments)."""
    if p == 0.0 or p == 1.0:
        frac = int(npts * p / 2.0)
    else:
        frac = int(npts * p / 2.0 + 0.5)
    idx1, idx2, idx3, idx4 = 0, frac - 1, npts - frac, npts - 1
    if idx1 == idx2:
        idx2 += 1
    if idx3 == idx4:
        idx3 -= 1
    w = np.zeros(npts)
    w[idx1:idx2 + 1] = 0.5 * (1.0 - np.cos(np.pi * (np.arange(idx1, idx2 + 1) - idx1) / (idx2 - idx1)))
    w[idx2 + 1:idx3] = 1.0
    w[idx3:idx4 + 1] = 0.5 * (1.0 + np.cos(np.pi * (idx3 - np.arange(idx3, idx4 + 1)) / (idx4 - idx3)))
    return w


def taper_window(npts, sampling_rate, max_percentage=None, type="hann", max_length=None, side="both"):
    """The multiplicative window obspy ``Trace.taper`` builds."""
    max_half = []
    if max_percentage is not None:
        max_half.append(int(max_percentage * npts))
    if max_length is not None:
        max_half.append(int(max_length * sampling_rate))
    max_half.append(int(npts / 2))
    wlen = min(max_half)
    M = 2 * wlen if 2 * wlen == npts else 2 * wlen + 1
    if type == "cosine":
        win = cosine_taper(M, 1.0)
    else:
        win = getattr(signal.windows, type)(M)
    if side == "left":
        return np.hstack((win[:wlen], np.ones(npts - wlen)))
    if side == "right":
        return np.hstack((np.ones(npts - wlen), win[len(win) - wlen:]))
    return np.hstack((win[:wlen], np.ones(npts - 2 * wlen), win[len(win) - wlen:]))


def taper(x, sampling_rate, max_percentage=None, type="hann", max_length=None, side="both"):
    x = np.asarray(x, dtype=np.float64)
    return x * taper_window(len(x), sampling_rate, max_percentage, type, max_length, side)


def butter_sos(sampling_rate, type, freq=None, freqmin=None, freqmax=None, corners=4):
    """Second-order sections of obspy.signal.filter.{bandpass,highpass,lowpass}."""
    fe = 0.5 * sampling_rate
    if type == "bandpass":
        low, high = freqmin / fe, freqmax / fe
        if high - 1.0 > -1e-6:          # obspy: falls back to a highpass
            return butter_sos(sampling_rate, "highpass", freq=freqmin, corners=corners)
        if low > 1:
            raise ValueError("Selected low corner frequency is above Nyquist.")
        z, p, k = signal.iirfilter(corners, [low, high], btype="band", ftype="butter", output="zpk")
    elif type == "highpass":
        f = freq / fe
        if f > 1:
            raise ValueError("Selected corner frequency is above Nyquist.")
        z, p, k = signal.iirfilter(corners, f, btype="highpass", ftype="butter", output="zpk")
    elif type == "lowpass":
        f = min(freq / fe, 1.0)
        z, p, k = signal.iirfilter(corners, f, btype="lowpass", ftype="butter", output="zpk")
    else:
        raise ValueError(f"unsupported filter type {type!r}")
    return signal.zpk2sos(z, p, k)


def filter_(x, sampling_rate, type, zerophase=False, **opts):
    """obspy ``Trace.filter``: causal ``sosfilt`` from zero state (twice, time
    reversed in between, for ``zerophase``)."""
    x = np.asarray(x, dtype=np.float64)
    if type == "lowpass_cheby_2":
        sos = cheby2_lowpass_sos(sampling_rate, opts["freq"], opts.get("maxorder", 12))
        return signal.sosfilt(sos, x)
    sos = butter_sos(sampling_rate, type, **opts)
    if zerophase:
        firstpass = signal.sosfilt(sos, x)
        return signal.sosfilt(sos, firstpass[::-1])[::-1]
    return signal.sosfilt(sos, x)


def cheby2_lowpass_sos(sampling_rate, freq, maxorder=12):
    """obspy.signal.filter.lowpass_cheby_2 design (rp=1, rs=96)."""
    nyquist = sampling_rate * 0.5
    rp, rs, order = 1, 96, 1e99
    ws = freq / nyquist
    wp = ws
    if ws > 1:
        ws = 1.0
    while True:
        if order <= maxorder:
            break
        wp = wp * 0.99
        order, wn = signal.cheb2ord(wp, ws, rp, rs, analog=0)
    z, p, k = signal.cheby2(order, rs, wn, btype="low", analog=0, output="zpk")
    return signal.zpk2sos(z, p, k)


def decimate(x, sampling_rate, factor, no_filter=False):
    """obspy ``Trace.decimate``: Chebyshev-II anti-alias low-pass at
    ``0.5*fs/factor`` then ``data[::factor]``.  Returns (y, new_rate)."""
    if factor > 16:
        raise ArithmeticError("Automatic filter design is unstable for decimation factors above 16.")
    x = np.asarray(x, dtype=np.float64)
    if not no_filter:
        x = signal.sosfilt(cheby2_lowpass_sos(sampling_rate, sampling_rate * 0.5 / float(factor)), x)
    return x[::factor].copy(), sampling_rate / float(factor)


def resample(x, sampling_rate, new_sampling_rate, window="hann", no_filter=True):
    """obspy ``Trace.resample`` (Fourier method): real FFT of the
    whole trace in scipy.fftpack's packed real layout, times the ifftshift-ed window,
    real and imaginary parts linearly interpolated (``np.interp``) onto the new
    frequency grid, inverse real FFT to ``num = int(npts / factor)`` samples, scaled
    by ``num / npts``.  Returns (data, new_sampling_rate)."""
    x = np.asarray(x, dtype=np.float64)
    npts = len(x)
    factor = sampling_rate / float(new_sampling_rate)
    if not no_filter:
        if factor > 16:
            raise ArithmeticError("Automatic filter design is unstable for resampling factors above 16.")
        x = signal.sosfilt(cheby2_lowpass_sos(sampling_rate, sampling_rate * 0.5 / float(factor)), x)
    X = np.fft.rfft(x)
    x_r = X.real.copy()
    x_i = X.imag.copy()
    x_i[0] = 0.0
    if npts % 2 == 0:
        x_i[-1] = 0.0
    if window is not None:
        large_w = np.fft.ifftshift(signal.get_window(window, npts))
        x_r *= large_w[:npts // 2 + 1]
        x_i *= large_w[:npts // 2 + 1]
    num = int(npts / factor)
    delta = 1.0 / sampling_rate
    df = 1.0 / (npts * delta)
    d_large_f = 1.0 / num * new_sampling_rate
    f = df * np.arange(0, npts // 2 + 1, dtype=np.int32)
    n_large_f = num // 2 + 1
    large_f = d_large_f * np.arange(0, n_large_f, dtype=np.int32)
    y_r = np.interp(large_f, f, x_r)
    y_i = np.interp(large_f, f, x_i)
    y_i[0] = 0.0
    if num % 2 == 0:
        y_i[-1] = 0.0
    y = np.fft.irfft(y_r + 1j * y_i, n=num) * (float(num) / float(npts))
    return y, float(new_sampling_rate)


def interpolate_lanczos(x, sampling_rate, new_sampling_rate, a, time_shift=0.0, start_offset=0.0, npts=None):
    """obspy ``Trace.interpolate(method='lanczos', a=a)`` -> ``lanczos_interpolation`` (C:
    obspy/signal/src/lanczos.c): windowed-sinc kernel of half-width ``a`` input samples
    evaluated at the new sample times.  ``start_offset``: new start minus old start in
    seconds (after ``time_shift``).  Returns (data, new_sampling_rate)."""
    x = np.asarray(x, dtype=np.float64)
    old_dt = 1.0 / sampling_rate
    new_dt = 1.0 / float(new_sampling_rate)
    old_len = (len(x) - 1) * old_dt
    start = start_offset
    if npts is None:
        npts = int(np.floor((old_len - start) / new_dt)) + 1
    dt = new_dt / old_dt
    offset = start / old_dt
    y = np.zeros(npts)
    idx = np.arange(npts)
    pos = dt * idx + offset
    fl = np.floor(pos).astype(np.int64)

    def sinc(u):
        out = np.ones_like(u)
        nz = np.abs(u) >= 1e-10
        out[nz] = np.sin(np.pi * u[nz]) / (np.pi * u[nz])
        return out

    for m in range(-a, a + 1):
        i = fl - m
        ok = (i >= 0) & (i < len(x))
        u = pos - i
        ok &= (u >= -a) & (u <= a)
        ii = np.where(ok, i, 0)
        y += np.where(ok, x[ii] * sinc(u) * sinc(u / a), 0.0)
    return y, float(new_sampling_rate)


def running_rms(x, sampling_rate, window=1.0):
    """``stream/running_rms.py:73-113``: ``int(int(endtime - starttime) / window)`` output samples, each the
    root-mean-square of one obspy ``Trace.slide(window_length=window, step=window)`` window (both end samples
    included: ``window * fs + 1`` samples).  Returns (data, new_sampling_rate = 1 / window)."""
    x = np.asarray(x, dtype=np.float64)
    duration = (len(x) - 1) / sampling_rate
    npts = int(int(duration) / window)
    step = int(round(window * sampling_rate))
    out = np.zeros(npts)
    for k in range(npts):
        w = x[k * step:k * step + step + 1]
        out[k] = np.sqrt(np.sum(np.power(w, 2)) / len(w))
    return out, 1.0 / window


def normalize(x, norm=None):
    """obspy ``Trace.normalize``: divide by ``abs(norm)``; ``norm=None`` -> the trace's largest absolute value."""
    x = np.asarray(x, dtype=np.float64)
    if norm is None:
        norm = np.max(np.abs(x))
    return x / abs(norm)


def trim(x, start=0, length=None):
    """Array form of the time-based trims (``stream/client.py:1017`` and
    ``core/process.py:156-157``): samples ``[start, start+length)``."""
    return x[start:] if length is None else x[start:start + length]


def trim_time(x, sampling_rate, trace_start, starttime=None, endtime=None):
    """obspy ``Trace.trim(starttime, endtime, pad=False, nearest_sample=True)`` on an
    array whose first sample sits at `trace_start` seconds (what ``('trim', {})`` becomes once
    ``stream/process.py:153-196`` injected the window, and ``stream/client.py:1017``).
    Follows ``Trace._ltrim`` then ``Trace._rtrim``.  Returns ``(y, new_trace_start)``."""
    def round_away(v):                      # obspy.core.compatibility.round_away
        return int(np.sign(v) * np.floor(np.abs(v) + 0.5))
    y = np.asarray(x)
    if starttime is not None:               # _ltrim
        delta = round_away((starttime - trace_start) * sampling_rate)
        if delta > 0:
            y = y[delta:]
            trace_start = trace_start + delta / sampling_rate
    if endtime is not None:                 # _rtrim
        if endtime < trace_start:           # "endtime < starttime": obspy leaves an empty trace
            return y[:0], trace_start
        delta = round_away((endtime - trace_start) * sampling_rate) - len(y) + 1
        if delta < 0:
            y = y[:len(y) - abs(delta)]
    return y, trace_start


def process(x, sampling_rate, operations, trace_start=0.0, starttime=None, endtime=None, return_start=False):
    """Apply an op-list ``[(name, params), ...]`` to one trace.  Returns
    ``(y, sampling_rate)`` (and the time of the first output sample with `return_start`).
    ``('trim', {})`` receives `starttime` / `endtime` like ``stream/process.py:153-196``."""
    y = np.asarray(x, dtype=np.float64)
    fs = float(sampling_rate)
    t0 = float(trace_start)
    for name, params in operations:
        params = dict(params)
        if name == "trim" and "start" not in params and "length" not in params:
            y, t0 = trim_time(y, fs, t0, params.get("starttime", starttime), params.get("endtime", endtime))
            continue
        if name == "detrend":
            y = detrend(y, params.get("type", "simple"))
        elif name == "taper":
            y = taper(y, fs, **params)
        elif name == "filter":
            y = filter_(y, fs, **params)
        elif name == "decimate":
            y, fs = decimate(y, fs, **params)
        elif name == "trim":
            y = trim(y, **params)
        elif name == "running_rms":
            y, fs = running_rms(y, fs, **params)
        elif name == "normalize":
            y = normalize(y, **params)
        elif name == "resample":
            params = dict(params)
            y, fs = resample(y, fs, params.pop("sampling_rate"), **params)
        elif name == "resample_poly":
            # the north-star polyphase decimator: scipy's own implementation IS the published algorithm (pinned)
            up, down = int(params.pop("up")), int(params.pop("down"))
            y = signal.resample_poly(y, up, down, **params)
            fs = fs * up / down
        elif name == "interpolate":
            params = dict(params)
            if params.pop("method", None) != "lanczos":
                raise NotImplementedError("only method='lanczos' is restated")
            y, fs = interpolate_lanczos(y, fs, params.pop("sampling_rate"), **params)
        else:
            raise NotImplementedError(f"operation {name!r} is outside the array-valued hot path")
    return (y, fs, t0) if return_start else (y, fs)
