/* xcorr_b200 -- C ABI of the B200 cross-correlation hot path.
 *
 * The reference (psmsmets/xcorr) is pure Python and has no FFI; these entry
 * points are what a binding for its hot path would call.  Each one names the
 * reference interface it replaces (paths relative to the reference's xcorr/).
 *
 * Conventions
 *   - every function returns 0 (XCB_OK) or a negative xcb_status; no C++
 *     exception crosses the boundary; xcb_strerror() explains a code;
 *   - "device" pointers are caller-owned CUDA device memory (the Python host
 *     side passes torch tensors' data_ptr()); `stream` is a cudaStream_t passed
 *     as void* (NULL = default stream); device calls are asynchronous on it;
 *   - a plan is immutable after creation and may be shared by threads; the
 *     workspace belongs to one stream at a time;
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with XCB_ERR_CUDA.
 */
#ifndef XCORR_B200_H
#define XCORR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XCB_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define XCB_API __attribute__((visibility("default")))
#else
#define XCB_API
#endif

typedef enum xcb_status {
  XCB_OK = 0,
  XCB_ERR_ARG = -1,      /* invalid argument / descriptor                      */
  XCB_ERR_SIZE = -2,     /* no supported transform length covers n             */
  XCB_ERR_CUDA = -3,     /* CUDA runtime error (see xcb_last_cuda_error)       */
  XCB_ERR_ALLOC = -4,    /* host or device allocation failed                   */
  XCB_ERR_UNSUPPORTED = -5
} xcb_status;

typedef enum xcb_dtype { XCB_F32 = 0, XCB_F64 = 1 } xcb_dtype;

typedef struct xcb_plan xcb_plan;

/* Describes one batch of the reference's hot loop (core/process.py:82-165):
 * n_receivers signals per window, n_windows windows, n_pairs receiver pairs,
 * one cc() per (pair, window) cropped to [lag_index_min, lag_index_max). */
typedef struct xcb_desc {
  int64_t n;               /* samples per window at the correlation rate (lag.attrs['npts'], core/init.py:120) */
  int32_t dtype;           /* xcb_dtype: dataset.cc.dtype (core/init.py:30); inputs are cast to it before the FFT (util/cc.py:68-71) */
  int32_t nd;              /* 0: choose the transform automatically; else number of passes */
  int32_t dims[4];         /* pass lengths when nd > 0 (testing / tuning) */
  int32_t n_receivers;
  int32_t n_windows;
  int32_t n_pairs;
  int32_t recv_stride;     /* signal index of (receiver r, window w) = r*recv_stride + w*win_stride; 0,0 -> [receiver][window] */
  int32_t win_stride;
  int32_t normalize;       /* util/cc.py:76-77 */
  int32_t unbiased;        /* intended util/cc.py:79 == signal/unbias.py:85 weights n/(n-|k|) */
  int32_t stack;           /* 1: ccf[pair][lag] += sum over the windows (cross-spectra summed before ONE inverse per pair) */
  const int32_t* pairs;    /* host, [n_pairs][2] receiver indices (i, j): Rxy[k] = sum_m x_i[m+k] x_j[m] */
  int64_t lag_index_min;   /* lag.attrs['index_min'] (core/init.py:191-208, util/time.py:94-108) */
  int64_t lag_index_max;   /* exclusive; both 0 -> full 2n-1 lags */
  int64_t max_work_bytes;  /* cap for the inverse-transform scratch; 0 -> 16 GiB */
  int32_t n_blocks;        /* 0: automatic; 1: one transform of the whole window; >1 (with nd, dims, block_len): block mode
                            * (cropped lags, shorter transforms) */
  int32_t reserved;
  int64_t block_len;       /* samples per block when n_blocks > 1 */
} xcb_desc;

typedef struct xcb_plan_info {
  int64_t n, fft_len;      /* L = 2M >= 2n-1 */
  int64_t spectrum_elems;  /* complex elements per signal in `spec` (M, or 2*n_blocks*M in block mode) */
  int64_t n_lags;          /* K = lag_index_max - lag_index_min */
  int64_t n_signals;       /* n_receivers * n_windows */
  int64_t n_jobs;          /* inverse transforms per call: n_pairs * n_windows (n_pairs when stacking) */
  int64_t jobs_in_flight;
  size_t workspace_bytes;
  int32_t nd;
  int32_t dims[4];
  int32_t dtype;
  int32_t n_blocks;        /* 1 = classic */
  int64_t block_len;
  int32_t n_groups;        /* job groups per launch of the staged pair product (0: the plain row kernel runs) */
  int32_t reserved;
} xcb_plan_info;

XCB_API const char* xcb_strerror(int status);
XCB_API const char* xcb_last_cuda_error(void);   /* thread-local text of the last CUDA failure */
XCB_API int xcb_version(void);

/* Transform length the planner picks for n samples (host only, no GPU needed). */
XCB_API int xcb_choose_transform(int64_t n, int32_t dims[4], int32_t* nd);

/* Creates twiddle tables and the job table on the current CUDA device. */
XCB_API int xcb_plan_create(xcb_plan** plan, const xcb_desc* desc);
XCB_API int xcb_plan_destroy(xcb_plan* plan);
XCB_API int xcb_plan_get_info(const xcb_plan* plan, xcb_plan_info* info);

/* Per receiver-window: cast/zero-pad, forward transform, sum of squares.
 * Replaces the two fft() calls and both np.linalg.norm() of util/cc.py:66-77,
 * done once per receiver-window instead of once per pair.
 *   x      device, dtype[n_signals][x_stride] (x_stride >= n)
 *   spec   device, complex<dtype>[n_signals][spectrum_elems] (half spectrum rfft(x)[k], k < M, in
 *          digit-transposed order; slot 0 = X[0] + i X[M])
 *   sumsq  device, double[n_signals] */
XCB_API int xcb_forward(const xcb_plan* plan, const void* x, int64_t x_stride, void* spec, double* sumsq,
                void* workspace, void* stream);

/* Per (pair, window): conjugate product, inverse transform, fftshift, crop,
 * scaling.  Replaces util/cc.py:75-79 and the [index_min:index_max] slice of
 * core/process.py:155-162 (+ signal/unbias.py:85 when desc.unbiased).
 *   ccf    device, dtype[n_pairs][n_windows][n_lags]  (or [n_pairs][n_lags] when stacking;
 *          the caller zero-fills it before the first stacked call) */
XCB_API int xcb_correlate(const xcb_plan* plan, const void* spec, const double* sumsq, void* ccf,
                  void* workspace, void* stream);

/* Like xcb_correlate, for callers that keep only the maximum of every correlation function: the kept lags
 * [lag_index_min, lag_index_max) are reduced inside the last inverse pass and never stored.  Replaces
 * `cc2.argmax()` + the value look-up of scripts/timelapse.py:255-268 (and util/cc.py:144-206 for 1-D).
 *   mask   NULL, or the 2-D validity of a row-embedded image correlation (signal/correlate.py:228-244 run on
 *          1-D kernels): lag = dF * row_width + dT is kept iff |dT| < half_width and |dF| < n_rows
 *   index  device int64[n_jobs]: first maximal index within the kept lags (np.argmax: a NaN wins), job = w*n_pairs + p
 *   value  device dtype[n_jobs] or NULL
 * Plans with at least two passes, no stacking. */
typedef struct xcb_lag_mask {
  int64_t row_width;   /* W: samples per embedded image row (>= 2*half_width - 1) */
  int64_t half_width;  /* T: image columns */
  int64_t n_rows;      /* F: image rows; n = F * W */
} xcb_lag_mask;
XCB_API int xcb_correlate_argmax(const xcb_plan* plan, const void* spec, const double* sumsq, const xcb_lag_mask* mask,
                                 int64_t* index, void* value, void* workspace, void* stream);

/* First index of the maximum of each row (np.argmax semantics, util/cc.py:144-206).
 *   rows   device, dtype[n_rows][row_len];  index device int64[n_rows];  value device dtype[n_rows] or NULL */
XCB_API int xcb_argmax(int32_t dtype, const void* rows, int64_t n_rows, int64_t row_len, int64_t* index,
               void* value, void* stream);

/* ------------------------------------------------------------------------
 * Stream preprocessing: the array-valued obspy Trace operations named by the
 * reference's op-lists (stream/process.py:29-94, applied at :280 through
 * eval('waveforms.<op>(**params)')).  Buffers are device arrays
 * dtype[n_signals][stride]; arithmetic is float64 whatever the storage dtype.
 * ------------------------------------------------------------------------ */

/* detrend + taper as one pointwise map  y[i] = w(i) * (x[i] - (c0 + c1*i)).
 * Replaces Trace.detrend(type) and Trace.taper(...): the ramps are the first /
 * last `wlen` values of the obspy window, computed on the host. */
typedef struct xcb_pointwise {
  int32_t detrend;           /* 0 none, 1 'demean'/'constant', 2 'linear', 3 'simple' */
  int32_t taper_left;        /* ramp lengths in samples (0 = none) */
  int32_t taper_right;
  int32_t reserved;
  const double* ramp_left;   /* device, taper_left values  */
  const double* ramp_right;  /* device, taper_right values */
} xcb_pointwise;

/* Causal cascade of second-order sections == scipy.signal.sosfilt(sos, x) from
 * zero state, i.e. Trace.filter(bandpass|highpass|lowpass|lowpass_cheby_2) and,
 * with decimate > 1, Trace.decimate(factor) (low-pass + data[::factor]). */
typedef struct xcb_sos {
  int32_t n_sections;        /* 1..12 (several filters may be concatenated) */
  int32_t decimate;          /* store every decimate-th output sample; >= 1 */
  int32_t reverse;           /* 1: filter from the last sample backwards (second half of zerophase=True) */
  int32_t chunk;             /* samples per scan chunk (e.g. 1024) */
  double coef[12][5];        /* b0 b1 b2 a1 a2 per section (a0 = 1) */
  const double* carry;       /* device, [2S][2S] row-major: state transition over `chunk` samples */
  const double* carry_group; /* device, [2S][2S]: carry^256 (one carry group = 256 chunks) */
  const double* init_state;  /* device, [n_signals][2S] or NULL: sosfilt(..., zi=) state before the first processed
                              * sample (scipy layout zi[section][2]); used by sosfiltfilt (signal/filter.py:96) */
  const double* zin_response; /* device, [chunk / decimate][2S] or NULL: zero-input response of the kept samples of a
                              * chunk to unit start states.  Given (decimate >= 2, chunk % decimate == 0, forward), the
                              * kept outputs are formed as zero-state output + response * start state instead of
                              * rerunning the cascade over every input sample (exact by linearity). */
  int32_t carry_window;      /* K >= 1: the caller certifies ||carry^K||_inf <= 2^-80, i.e. a chunk's start state no
                              * longer depends (to far below float64 rounding) on anything older than K chunks; the
                              * start states are then formed in ONE parallel launch, every run of 32 chunks warmed up
                              * over the K chunks before it, instead of fold / sweep / expand over the whole trace.
                              * 0: unknown or slowly decaying filter (e.g. a 0.05 Hz high-pass at 250 Hz) -> the
                              * hierarchical, unconditionally exact form */
  int32_t reserved;
} xcb_sos;

/* stats[s] = {sum x, sum i*x, x[0], x[n-1]} (device double[n_signals][4]) */
XCB_API int xcb_stats_workspace_bytes(int64_t n, int32_t n_signals, size_t* bytes);
XCB_API int xcb_stats(int32_t dtype, const void* x, int64_t n, int64_t stride, int32_t n_signals, double* stats,
                      void* workspace, void* stream);
XCB_API int xcb_pointwise_apply(int32_t dtype_in, const void* x, int64_t x_stride, int32_t dtype_out, void* y,
                                int64_t y_stride, int64_t n, int32_t n_signals, const xcb_pointwise* pw,
                                const double* stats, void* stream);
/* y has ceil(n / decimate) samples per signal; `pre` (optional) is applied to x on the fly. */
XCB_API int xcb_sosfilt_workspace_bytes(int64_t n, int32_t n_signals, const xcb_sos* f, size_t* bytes);
XCB_API int xcb_sosfilt(int32_t dtype_in, const void* x, int64_t x_stride, int32_t dtype_out, void* y,
                        int64_t y_stride, int64_t n, int32_t n_signals, const xcb_sos* f, const xcb_pointwise* pre,
                        const double* stats, void* workspace, void* stream);

/* Polyphase FIR rate change (the north-star decimator; scipy.signal.upfirdn / resample_poly semantics, offered as
 * op 'resample_poly' beside the obspy-faithful ('decimate', {...}) of stream/process.py:34-37):
 *   y[k] = sum_i x[i] * taps[(k + k0) * down - i * up],   k < n_out
 * i.e. zero-stuffing by `up`, FIR, keep every `down`-th sample, skip the first k0 outputs -- computed directly,
 * n_taps / up multiply-adds per output, float64 accumulation.  taps: device double[n_taps] (host design:
 * scipy.signal.firwin as resample_poly does it); `pre` (optional) is applied to x on the fly. */
XCB_API int xcb_upfirdn(int32_t dtype_in, const void* x, int64_t x_stride, int64_t n_in, int32_t dtype_out, void* y,
                        int64_t y_stride, int64_t n_out, int32_t n_signals, const double* taps, int32_t n_taps,
                        int32_t up, int32_t down, int64_t k0, const xcb_pointwise* pre, const double* stats,
                        void* stream);

/* ------------------------------------------------------------------------
 * Plain complex transforms on the same pass kernels, and DFTs of ARBITRARY
 * length through Bluestein's algorithm.  They replace the numpy / pyfftw calls
 * of signal/fft.py:15-18,109-112 and the whole-trace FFTs inside obspy
 * Trace.resample (stream/process.py:61) and signal/timeshift.py:90-96, which
 * need the DFT of exactly npts points.
 * ------------------------------------------------------------------------ */
typedef struct xcb_fft xcb_fft;

/* smallest supported transform length >= n_min (product of 1-4 compiled pass sizes) */
XCB_API int xcb_fft_good_length(int64_t n_min, int64_t* length);
XCB_API int xcb_fft_create(xcb_fft** fft, int32_t dtype, int64_t length, int32_t max_batch);
XCB_API int xcb_fft_destroy(xcb_fft* fft);
/* spec[b][length] = DFT of x[b][0..n_valid) zero-padded to `length` (sign -1), in the plan's
 * digit-transposed order: pointwise products of spectra of the same plan need no reordering.
 *   x device complex<dtype>[batch][x_stride]; spec device complex<dtype>[batch][length] */
XCB_API int xcb_fft_forward(const xcb_fft* fft, const void* x, int64_t x_stride, int64_t n_valid, int32_t batch,
                            void* spec, void* stream);
/* y[b][0..n_out) = (1/length) * inverse DFT of conj(spec_conj[b]); spec_conj (the CONJUGATE of the
 * spectrum, digit-transposed) is transformed in place and destroyed. */
XCB_API int xcb_fft_inverse_conj(const xcb_fft* fft, void* spec_conj, int32_t batch, void* y, int64_t y_stride,
                                 int64_t n_out, void* stream);
/* y[k] = scale * sum_n x[n] exp(sign * 2 pi i n k / K), k < n_out <= K, for ANY K with 2K - 1 <= length.
 *   x_mode 0: x real dtype[batch][x_stride]; 1: complex; 2: half spectrum (K/2 + 1 entries) of a real
 *   signal, extended by conjugate symmetry (irfft; n_in must be K).  n_in <= K samples, zero beyond.
 *   y complex<dtype>[batch][y_stride], or its real part only (y_real).
 *   filter_spec device complex<dtype>[length]: spectrum of the chirp filter of (K, sign); written when
 *   filter_ready == 0, reused afterwards.  work device complex<dtype>[batch * (K + length)]. */
XCB_API int xcb_dft_bluestein(const xcb_fft* fft, int64_t K, int32_t sign, const void* x, int32_t x_mode,
                              int64_t x_stride, int64_t n_in, int32_t batch, void* y, int32_t y_real,
                              int64_t y_stride, int64_t n_out, double scale, void* filter_spec,
                              int32_t filter_ready, void* work, void* stream);

/* X[b][k] *= exp(-2 pi i * delay[b] * df * k), k < m: the phase ramp of signal/timeshift.py:90-96
 * (delay: device double[batch], seconds; df: frequency step in Hz). */
XCB_API int xcb_phase_ramp(int32_t dtype, void* spec, int64_t stride, int64_t m, int32_t batch, const double* delay,
                           double df, void* stream);
/* obspy Trace.resample (stream/process.py:52 -> eval at :280): spectrum of npts samples (npts/2 + 1 bins)
 * times the ifftshift-ed window (0 none, 1 'hann'), linearly interpolated onto the num/2 + 1 bins of the
 * new length (np.interp semantics, frequency steps df -> dlf). */
XCB_API int xcb_resample_spectrum(int32_t dtype, const void* spec_in, int64_t in_stride, void* spec_out,
                                  int64_t out_stride, int64_t npts, int64_t num, double df, double dlf,
                                  int32_t window, int32_t batch, void* stream);
/* obspy Trace.interpolate(method='lanczos', a): y[idx] = sum_i x[i] L(dt*idx + offset - i), L = sinc(u) sinc(u/a)
 * on |u| <= a; dt and offset in units of input samples. */
XCB_API int xcb_lanczos(int32_t dtype, const void* x, int64_t x_stride, int64_t n_in, void* y, int64_t y_stride,
                        int64_t n_out, double dt, double offset, int32_t a, int32_t batch, void* stream);
/* signal/snr.py:107-121: per row max |x| over [sig_begin, sig_end) with its first index, and the sum of
 * squares over [noise_begin, noise_end).  out: device double[n_rows][3] = {max, index, sumsq}. */
XCB_API int xcb_window_stats(int32_t dtype, const void* x, int64_t stride, int32_t n_rows, int64_t sig_begin,
                             int64_t sig_end, int64_t noise_begin, int64_t noise_end, double* out, void* stream);

/* signal/spectrogram.py:125-145: scipy.signal.spectrogram(x, nperseg=win_len, noverlap=win_len-1, nfft, mode='psd',
 * one-sided) of every row -- one segment per sample, each mean-detrended (detrend != 0), windowed and
 * zero-padded to nfft -- placed `rint(win_len / 2)` samples in, zeros at the edges like the reference's np.pad.
 *   window device double[win_len]; scale = 1/(fs * sum w^2) ('density') or 1/(sum w)^2 ('spectrum');
 *   y device dtype[n_signals][nfft/2 + 1][n] */
XCB_API int xcb_spectrogram(int32_t dtype, const void* x, int64_t x_stride, int64_t n, int32_t n_signals,
                            const double* window, int32_t win_len, int32_t nfft, double scale, int32_t detrend,
                            void* y, void* stream);

/* stream/running_rms.py:73-113 (op 'running_rms'): out[s][k] = sqrt(mean(x[s][k*step : k*step + win]^2)),
 * k < n_windows; out device double[n_signals][n_windows]. */
XCB_API int xcb_running_rms(int32_t dtype, const void* x, int64_t stride, int64_t n, int32_t n_signals, int64_t win,
                            int64_t step, int64_t n_windows, double* out, void* stream);

/* obspy Trace.normalize (op 'normalize'): x[s] /= max |x[s]| with window_stats = the output of xcb_window_stats over
 * the whole trace (in place), or x /= |norm| when window_stats is NULL. */
XCB_API int xcb_normalize_rows(int32_t dtype, void* x, int64_t stride, int64_t n, int32_t n_signals,
                               const double* window_stats, double norm, void* stream);

/* signal/normalize.py:33-155 (norm1d, norm2d; used by correlate1d/2d): every row divided by its 2-norm in place,
 * the norm being sqrt(window_stats[s][2]) from xcb_window_stats over the whole row (fixed-order float64 sum). */
XCB_API int xcb_l2_normalize_rows(int32_t dtype, void* x, int64_t stride, int64_t n, int32_t n_signals,
                                  const double* window_stats, void* stream);

/* Measurement aid (bench.py): when enabled, every pass kernel this plan launches
 * is bracketed by CUDA events on its stream.  Kinds 0..nd-1 are the forward
 * passes (axis order), nd..2nd-1 the inverse passes of the same axes.  Not
 * thread-safe; read after synchronising the stream.  total_ms / launches hold
 * 8 entries. */
XCB_API int xcb_profile_enable(xcb_plan* plan, int32_t enable);
XCB_API int xcb_profile_read(xcb_plan* plan, int32_t* n_kinds, double* total_ms, int64_t* launches);

/* cc(x, y, normalize, pad=True, unbiased, dtype) with HOST buffers
 * (util/cc.py:22-79): copies in, runs the device path, copies the
 * [lag_index_min, lag_index_max) lags back, synchronises.  out has
 * (lag_index_max - lag_index_min) elements (both 0 -> 2n-1). */
XCB_API int xcb_cc_host(int32_t dtype, int64_t n, const void* x, const void* y, int32_t normalize,
                int32_t unbiased, int64_t lag_index_min, int64_t lag_index_max, void* out);
/* xcb_cc_host keeps the plan, device buffers and stream of the last few (device, dtype, n, crop, flags)
 * it was called with (the reference calls cc() once per pair and window with the same shape,
 * core/process.py:155-162); this frees them. */
XCB_API int xcb_cc_host_clear_cache(void);

#ifdef __cplusplus
}
#endif
#endif /* XCORR_B200_H */
