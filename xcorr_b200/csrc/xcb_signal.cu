// Small array kernels next to the FFT path: spectral resampling and Lanczos
// interpolation (obspy Trace.resample / Trace.interpolate, named by the
// reference's op-lists, stream/process.py:51-52), the phase ramp of
// signal/timeshift.py:90-96 and the window statistics of signal/snr.py:107-121.
// Arithmetic is float64 whatever the storage dtype.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/xcorr_b200.h"

namespace {

template <class T> struct c2 { T x, y; };

// X[b][k] *= exp(-2 pi i * delay[b] * df * k)
template <class T>
__global__ void k_phase_ramp(c2<T>* __restrict__ X, int64_t stride, int64_t m, const double* __restrict__ delay,
                             double df) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int64_t b = blockIdx.y;
  // phase in turns, reduced before the sincos
  double turns = delay[b] * (df * (double)k);
  turns -= floor(turns);
  double s, c;
  sincospi(-2.0 * turns, &s, &c);
  c2<T>* p = X + b * stride + k;
  const double re = p->x, im = p->y;
  p->x = (T)(re * c - im * s);
  p->y = (T)(re * s + im * c);
}

// obspy Trace.resample: windowed spectrum on the old frequency grid f = df * k,
// k <= npts/2, linearly interpolated (np.interp, clamped at the ends) onto
// large_f = dlf * j, j <= num/2.  window: 0 none, 1 'hann' (periodic,
// ifftshift-ed: large_w[k] = 0.5 - 0.5 cos(2 pi (k + npts/2) / npts)).
// The fftpack real format drops the imaginary parts of the zero bin and (num
// even) of the last bin.
template <class T>
__global__ void k_resample_spectrum(const c2<T>* __restrict__ X, int64_t x_stride, c2<T>* __restrict__ Y,
                                    int64_t y_stride, int64_t npts, int64_t num, double df, double dlf, int window) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m_in = npts / 2 + 1, m_out = num / 2 + 1;
  if (j >= m_out) return;
  const int64_t b = blockIdx.y;
  const c2<T>* x = X + b * x_stride;
  auto win = [&](int64_t k) -> double {
    if (window == 0) return 1.0;
    return 0.5 - 0.5 * cospi(2.0 * (double)((k + npts / 2) % npts) / (double)npts);
  };
  // fftpack layout: x_r[k] = Re X[k]; x_i[0] = 0, x_i[k] = Im X[k], and for even npts x_i[npts/2] = 0
  auto xr = [&](int64_t k) -> double { return (double)x[k].x * win(k); };
  auto xi = [&](int64_t k) -> double {
    if (k == 0 || (npts % 2 == 0 && k == npts / 2)) return 0.0;
    return (double)x[k].y * win(k);
  };
  const double lf = dlf * (double)j;
  double re, im;
  const double f_last = df * (double)(m_in - 1);
  if (lf >= f_last) { re = xr(m_in - 1); im = xi(m_in - 1); }
  else {
    // np.interp: largest k with f[k] <= lf
    int64_t k = (int64_t)floor(lf / df);
    if (k < 0) k = 0;
    if (k > m_in - 2) k = m_in - 2;
    while (k > 0 && df * (double)k > lf) --k;
    while (k < m_in - 2 && df * (double)(k + 1) <= lf) ++k;
    const double f0 = df * (double)k, f1 = df * (double)(k + 1);
    const double r0 = xr(k), r1 = xr(k + 1), i0 = xi(k), i1 = xi(k + 1);
    re = (r1 - r0) / (f1 - f0) * (lf - f0) + r0;
    im = (i1 - i0) / (f1 - f0) * (lf - f0) + i0;
  }
  if (j == 0 || (num % 2 == 0 && j == num / 2)) im = 0.0;
  c2<T> o; o.x = (T)re; o.y = (T)im;
  Y[b * y_stride + j] = o;
}

__device__ __forceinline__ double sinc_pi(double x) {
  if (fabs(x) < 1e-10) return 1.0;
  return sinpi(x) / (M_PI * x);
}

// obspy lanczos_resample (signal/src/lanczos.c): y[idx] = sum_{m=-a..a} x[i] L(pos - i), i = floor(pos) - m,
// pos = dt * idx + offset in units of input samples, L(u) = sinc(u) sinc(u / a) on |u| <= a.
template <class T>
__global__ void k_lanczos(const T* __restrict__ x, int64_t x_stride, int64_t n_in, T* __restrict__ y,
                          int64_t y_stride, int64_t n_out, double dt, double offset, int a) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_out) return;
  const int64_t b = blockIdx.y;
  const T* xs = x + b * x_stride;
  const double pos = dt * (double)idx + offset;
  const int64_t fl = (int64_t)floor(pos);
  double acc = 0.0;
  for (int m = -a; m <= a; ++m) {
    const int64_t i = fl - m;
    if (i < 0 || i >= n_in) continue;
    const double u = pos - (double)i;
    if (u < -(double)a || u > (double)a) continue;
    acc += (double)xs[i] * sinc_pi(u) * sinc_pi(u / (double)a);
  }
  y[b * y_stride + idx] = (T)acc;
}

// Per row: max |x| and its first index over [s0, s1), sum of squares over [n0, n1).  One block per row,
// fixed association order (bit-reproducible).
template <class T>
__global__ void __launch_bounds__(256) k_window_stats(const T* __restrict__ x, int64_t stride, int64_t s0, int64_t s1,
                                                      int64_t n0, int64_t n1, double* __restrict__ out) {
  const T* r = x + (int64_t)blockIdx.x * stride;
  double best = -1.0, ss = 0.0;
  int64_t bi = -1;
  bool nan_seen = false;
  for (int64_t i = s0 + threadIdx.x; i < s1; i += 256) {
    const double v = fabs((double)r[i]);
    if (v != v) { if (!nan_seen) { nan_seen = true; best = v; bi = i; } continue; }
    if (!nan_seen && v > best) { best = v; bi = i; }
  }
  for (int64_t i = n0 + threadIdx.x; i < n1; i += 256) { const double v = (double)r[i]; ss += v * v; }
  __shared__ double sv[256], sq[256];
  __shared__ int64_t si[256];
  __shared__ int sn[256];
  sv[threadIdx.x] = best; si[threadIdx.x] = bi; sq[threadIdx.x] = ss; sn[threadIdx.x] = nan_seen ? 1 : 0;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const int a = threadIdx.x, b = threadIdx.x + o;
      sq[a] += sq[b];
      bool take;
      if (sn[a] != sn[b]) take = sn[b] > sn[a];
      else if (sn[a]) take = si[b] >= 0 && (si[a] < 0 || si[b] < si[a]);
      else take = si[b] >= 0 && (si[a] < 0 || sv[b] > sv[a] || (sv[b] == sv[a] && si[b] < si[a]));
      if (take) { sv[a] = sv[b]; si[a] = si[b]; sn[a] = sn[b]; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double* o = out + (int64_t)blockIdx.x * 3;
    o[0] = sv[0];
    o[1] = (double)si[0];
    o[2] = sq[0];
  }
}

// scipy.signal.spectrogram with noverlap = nperseg - 1 (signal/spectrogram.py:125-139): one segment per
// sample, each detrended (mean), windowed, zero-padded to nfft and transformed; |X|^2 * scale, one-sided.
// A warp owns one frequency bin k and 32 consecutive segments: x[t + n] is read conflict free across
// the lanes, the window and the twiddle of (k, n) are warp-uniform broadcasts, and the stores of a bin
// are contiguous.  The mean is removed through the window's own transform: sum w (x - m) e = sum w x e - m W[k].
constexpr int kSpecSeg = 128;                  // segments per block
template <class T>
__global__ void __launch_bounds__(256) k_spectrogram(const T* __restrict__ x, int64_t x_stride, int64_t n,
                                                     const double* __restrict__ win, int win_len, int nfft,
                                                     double scale, int detrend, int64_t edge, T* __restrict__ y) {
  extern __shared__ unsigned char spec_sm_raw[];
  T* xs = reinterpret_cast<T*>(spec_sm_raw);                       // [kSpecSeg + win_len]
  T* ws = xs + kSpecSeg + win_len;                                 // [win_len]
  T* tc = ws + win_len;                                            // [nfft] cos
  T* tsn = tc + nfft;                                              // [nfft] -sin
  const int64_t n_seg = n - win_len + 1;
  const int64_t s0 = (int64_t)blockIdx.x * kSpecSeg;
  const int64_t sig = blockIdx.y;
  const T* xr = x + sig * x_stride;
  for (int i = threadIdx.x; i < kSpecSeg + win_len; i += 256) xs[i] = (s0 + i < n) ? xr[s0 + i] : T(0);
  for (int i = threadIdx.x; i < win_len; i += 256) ws[i] = (T)win[i];
  for (int i = threadIdx.x; i < nfft; i += 256) {
    double sn, cs;
    sincospi(2.0 * (double)i / (double)nfft, &sn, &cs);
    tc[i] = (T)cs; tsn[i] = (T)(-sn);
  }
  __syncthreads();
  const int n_bins = nfft / 2 + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* yo = y + sig * (int64_t)n_bins * n;
  for (int seg_blk = 0; seg_blk < kSpecSeg; seg_blk += 32) {
    const int sl = seg_blk + lane;                                 // local segment of this lane
    const bool live = s0 + sl < n_seg;
    T mean = T(0);
    if (detrend) {
      T a = T(0);
      for (int m = 0; m < win_len; ++m) a += xs[sl + m];
      mean = a / (T)win_len;
    }
    for (int k = warp; k < n_bins; k += 8) {
      T re = T(0), im = T(0), wre = T(0), wim = T(0);
      int idx = 0;
      for (int m = 0; m < win_len; ++m) {
        const T w = ws[m], c = w * tc[idx], sn = w * tsn[idx];
        const T v = xs[sl + m];
        re = fma(v, c, re); im = fma(v, sn, im);
        wre += c; wim += sn;
        idx += k; if (idx >= nfft) idx -= nfft;
      }
      re -= mean * wre; im -= mean * wim;
      T p = (T)((double)(re * re + im * im) * scale);
      if (k != 0 && !(nfft % 2 == 0 && k == nfft / 2)) p *= T(2);
      if (live) yo[(int64_t)k * n + edge + s0 + sl] = p;
    }
  }
}

// stream/running_rms.py:73-113: root-mean-square over consecutive windows of `win` samples starting every
// `step` samples (obspy Trace.slide: both window ends included).  One warp per (trace, window), fixed order.
template <class T>
__global__ void __launch_bounds__(256) k_running_rms(const T* __restrict__ x, int64_t stride, int64_t win, int64_t step,
                                                     int64_t n_windows, int32_t n_signals, double* __restrict__ out) {
  const int64_t w = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_windows * n_signals) return;
  const int64_t sig = w / n_windows, k = w - sig * n_windows;
  const T* p = x + sig * stride + k * step;
  double a = 0.0;
  for (int64_t i = lane; i < win; i += 32) { const double v = (double)p[i]; a = fma(v, v, a); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if (lane == 0) out[w] = sqrt(a / (double)win);
}

// obspy Trace.normalize(norm=None): x /= max |x| of the trace (stats[s][0] from k_window_stats); norm given: x /= |norm|
// L2: x /= sqrt(stats[s][2]) (sum of squares of the row, fixed-order reduction of k_window_stats) = np.linalg.norm
template <class T, bool L2 = false>
__global__ void k_normalize_rows(T* __restrict__ x, int64_t stride, int64_t n, const double* __restrict__ stats,
                                 double norm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t s = blockIdx.y;
  const double d = L2 ? sqrt(stats[s * 3 + 2]) : (stats ? stats[s * 3] : fabs(norm));
  x[s * stride + i] = (T)((double)x[s * stride + i] / d);
}

int fail(cudaError_t e) { return e == cudaSuccess ? XCB_OK : XCB_ERR_CUDA; }

}  // namespace

extern "C" int xcb_phase_ramp(int32_t dtype, void* spec, int64_t stride, int64_t m, int32_t batch,
                              const double* delay, double df, void* stream) {
  if (!spec || !delay || m < 1 || batch < 1 || stride < m) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((m + 255) / 256), (unsigned)batch);
  if (dtype == XCB_F32) k_phase_ramp<float><<<grid, 256, 0, st>>>(static_cast<c2<float>*>(spec), stride, m, delay, df);
  else if (dtype == XCB_F64) k_phase_ramp<double><<<grid, 256, 0, st>>>(static_cast<c2<double>*>(spec), stride, m, delay, df);
  else return XCB_ERR_ARG;
  return fail(cudaGetLastError());
}

extern "C" int xcb_resample_spectrum(int32_t dtype, const void* spec_in, int64_t in_stride, void* spec_out,
                                     int64_t out_stride, int64_t npts, int64_t num, double df, double dlf,
                                     int32_t window, int32_t batch, void* stream) {
  if (!spec_in || !spec_out || npts < 2 || num < 1 || batch < 1 || window < 0 || window > 1) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((num / 2 + 1 + 255) / 256), (unsigned)batch);
  if (dtype == XCB_F32)
    k_resample_spectrum<float><<<grid, 256, 0, st>>>(static_cast<const c2<float>*>(spec_in), in_stride,
                                                     static_cast<c2<float>*>(spec_out), out_stride, npts, num, df, dlf, window);
  else if (dtype == XCB_F64)
    k_resample_spectrum<double><<<grid, 256, 0, st>>>(static_cast<const c2<double>*>(spec_in), in_stride,
                                                      static_cast<c2<double>*>(spec_out), out_stride, npts, num, df, dlf, window);
  else return XCB_ERR_ARG;
  return fail(cudaGetLastError());
}

extern "C" int xcb_lanczos(int32_t dtype, const void* x, int64_t x_stride, int64_t n_in, void* y, int64_t y_stride,
                           int64_t n_out, double dt, double offset, int32_t a, int32_t batch, void* stream) {
  if (!x || !y || n_in < 1 || n_out < 1 || a < 1 || batch < 1) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((n_out + 255) / 256), (unsigned)batch);
  if (dtype == XCB_F32)
    k_lanczos<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), x_stride, n_in, static_cast<float*>(y), y_stride, n_out, dt, offset, a);
  else if (dtype == XCB_F64)
    k_lanczos<double><<<grid, 256, 0, st>>>(static_cast<const double*>(x), x_stride, n_in, static_cast<double*>(y), y_stride, n_out, dt, offset, a);
  else return XCB_ERR_ARG;
  return fail(cudaGetLastError());
}

extern "C" int xcb_window_stats(int32_t dtype, const void* x, int64_t stride, int32_t n_rows, int64_t sig_begin,
                                int64_t sig_end, int64_t noise_begin, int64_t noise_end, double* out, void* stream) {
  if (!x || !out || n_rows < 1 || sig_begin < 0 || sig_end < sig_begin || noise_begin < 0 || noise_end < noise_begin)
    return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XCB_F32)
    k_window_stats<float><<<(unsigned)n_rows, 256, 0, st>>>(static_cast<const float*>(x), stride, sig_begin, sig_end, noise_begin, noise_end, out);
  else if (dtype == XCB_F64)
    k_window_stats<double><<<(unsigned)n_rows, 256, 0, st>>>(static_cast<const double*>(x), stride, sig_begin, sig_end, noise_begin, noise_end, out);
  else return XCB_ERR_ARG;
  return fail(cudaGetLastError());
}

extern "C" int xcb_spectrogram(int32_t dtype, const void* x, int64_t x_stride, int64_t n, int32_t n_signals,
                               const double* window, int32_t win_len, int32_t nfft, double scale, int32_t detrend,
                               void* y, void* stream) {
  if (!x || !window || !y || n < 1 || n_signals < 1 || win_len < 1 || win_len > n || nfft < win_len) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t b = dtype == XCB_F32 ? 4 : 8;
  if (dtype != XCB_F32 && dtype != XCB_F64) return XCB_ERR_ARG;
  const int n_bins = nfft / 2 + 1;
  const int64_t n_seg = n - win_len + 1, edge = (int64_t)llrint((double)win_len / 2.0);
  // zero frame: the reference pads `edge` columns on the left and the rest on the right (np.pad)
  if (cudaMemsetAsync(y, 0, (size_t)n_signals * n_bins * n * b, st) != cudaSuccess) return XCB_ERR_CUDA;
  const size_t smem = b * ((size_t)kSpecSeg + 2 * (size_t)win_len + 2 * (size_t)nfft);
  if (smem > 200 * 1024) return XCB_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((n_seg + kSpecSeg - 1) / kSpecSeg), (unsigned)n_signals);
  cudaError_t e;
  if (dtype == XCB_F32) {
    e = cudaFuncSetAttribute(k_spectrogram<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return XCB_ERR_CUDA;
    k_spectrogram<float><<<grid, 256, smem, st>>>(static_cast<const float*>(x), x_stride, n, window, win_len, nfft, scale,
                                                  detrend, edge, static_cast<float*>(y));
  } else {
    e = cudaFuncSetAttribute(k_spectrogram<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return XCB_ERR_CUDA;
    k_spectrogram<double><<<grid, 256, smem, st>>>(static_cast<const double*>(x), x_stride, n, window, win_len, nfft,
                                                   scale, detrend, edge, static_cast<double*>(y));
  }
  return fail(cudaGetLastError());
}

extern "C" int xcb_running_rms(int32_t dtype, const void* x, int64_t stride, int64_t n, int32_t n_signals, int64_t win,
                               int64_t step, int64_t n_windows, double* out, void* stream) {
  if (!x || !out || n < 1 || n_signals < 1 || win < 1 || step < 1 || n_windows < 1) return XCB_ERR_ARG;
  if ((n_windows - 1) * step + win > n) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t warps = n_windows * n_signals;
  const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
  if (dtype == XCB_F32)
    k_running_rms<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), stride, win, step, n_windows, n_signals, out);
  else if (dtype == XCB_F64)
    k_running_rms<double><<<grid, 256, 0, st>>>(static_cast<const double*>(x), stride, win, step, n_windows, n_signals, out);
  else return XCB_ERR_ARG;
  return fail(cudaGetLastError());
}

extern "C" int xcb_normalize_rows(int32_t dtype, void* x, int64_t stride, int64_t n, int32_t n_signals,
                                  const double* window_stats, double norm, void* stream) {
  if (!x || n < 1 || n_signals < 1) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)n_signals);
  if (dtype == XCB_F32) k_normalize_rows<float><<<grid, 256, 0, st>>>(static_cast<float*>(x), stride, n, window_stats, norm);
  else if (dtype == XCB_F64) k_normalize_rows<double><<<grid, 256, 0, st>>>(static_cast<double*>(x), stride, n, window_stats, norm);
  else return XCB_ERR_ARG;
  return fail(cudaGetLastError());
}

// signal/normalize.py:33-155 (norm1d / norm2d): every row divided by its 2-norm, sqrt of the row's sum of
// squares in window_stats[s][2] (xcb_window_stats with the noise window [0, n)).
extern "C" int xcb_l2_normalize_rows(int32_t dtype, void* x, int64_t stride, int64_t n, int32_t n_signals,
                                     const double* window_stats, void* stream) {
  if (!x || !window_stats || n < 1 || n_signals < 1) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)n_signals);
  if (dtype == XCB_F32) k_normalize_rows<float, true><<<grid, 256, 0, st>>>(static_cast<float*>(x), stride, n, window_stats, 0.0);
  else if (dtype == XCB_F64) k_normalize_rows<double, true><<<grid, 256, 0, st>>>(static_cast<double*>(x), stride, n, window_stats, 0.0);
  else return XCB_ERR_ARG;
  return fail(cudaGetLastError());
}
