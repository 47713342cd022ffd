// Stream preprocessing on device: the array-valued obspy
// Trace operations the reference's op-lists name (stream/process.py:29-94).
//
//   xcb_stats      sum x, sum i*x, first, last  (detrend coefficients)   -- P1
//   xcb_pointwise  y = taper(i) * (x - (c0 + c1*i))                       -- P1, P2
//   xcb_sosfilt    causal cascade of second-order sections (scipy sosfilt,
//                  DF2T, zero initial state), optional fused pointwise
//                  prologue and decimating store                          -- P3, P4
//
// sosfilt is sequential in time; here it is a three-phase chunked scan that is
// exact (no warm-up approximation, so long-memory filters such as a 0.05 Hz
// high-pass at 250 Hz are fine):
//   1. every chunk (one thread, `chunk` samples) runs the cascade from zero
//      state and keeps its final state z_c (2S doubles);
//   2. the chunk start states follow s_{c+1} = P s_c + z_c with P the cascade's
//      state transition over one chunk (built on the host by simulating unit
//      states); done hierarchically: fold groups of 64 chunks, one short
//      sequential sweep over groups with P^64, expand inside groups;
//   3. every chunk reruns from its true start state and stores the outputs
//      (every `decimate`-th one).
// All arithmetic is float64 regardless of the storage dtype, in the same
// operation order as scipy's sosfilt.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/xcorr_b200.h"

namespace {

constexpr int kGroup = 256;       // chunks per carry group
constexpr int kMaxSections = 12;

struct PointwiseDev {
  int32_t detrend;                // 0 none, 1 demean, 2 linear, 3 simple
  int32_t taper_left, taper_right;
  const double* ramp_left;
  const double* ramp_right;
};

__device__ __forceinline__ void trend_coefficients(const PointwiseDev& pw, const double* st, int64_t n, double& c0,
                                                   double& c1) {
  c0 = 0.0;
  c1 = 0.0;
  if (pw.detrend == 1) {
    c0 = st[0] / (double)n;
  } else if (pw.detrend == 2) {
    const double nn = (double)n, ibar = 0.5 * (nn - 1.0);
    const double sxx = nn * (nn * nn - 1.0) / 12.0;
    c1 = n > 1 ? (st[1] - ibar * st[0]) / sxx : 0.0;
    c0 = st[0] / nn - c1 * ibar;
  } else if (pw.detrend == 3) {
    c0 = st[2];
    c1 = n > 1 ? (st[3] - st[2]) / (double)(n - 1) : 0.0;
  }
}

__device__ __forceinline__ double pointwise_value(const PointwiseDev& pw, double v, int64_t i, int64_t n, double c0,
                                                  double c1) {
  if (pw.detrend) v -= c0 + c1 * (double)i;
  if (i < pw.taper_left) v *= __ldg(pw.ramp_left + i);
  else if (i >= n - pw.taper_right) v *= __ldg(pw.ramp_right + (i - (n - pw.taper_right)));
  return v;
}

// ---------------------------------------------------------------- statistics
template <class T>
__global__ void __launch_bounds__(256) stats_partial(const T* __restrict__ x, int64_t n, int64_t stride, int blocks,
                                                     double* __restrict__ partial) {
  const int sig = blockIdx.y, b = blockIdx.x;
  const T* xs = x + (int64_t)sig * stride;
  const int64_t per = (n + blocks - 1) / blocks;
  const int64_t lo = (int64_t)b * per;
  int64_t hi = lo + per;
  if (hi > n) hi = n;
  double s0 = 0.0, s1 = 0.0;
  // four independent loads per thread and trip (the reduction pass is pure load latency: ncu long_scoreboard
  // 21-32 warps per issue); the order of the additions is fixed, so the sums stay bit reproducible
  for (int64_t i = lo + threadIdx.x; i < hi; i += 1024) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = i + u * 256 < hi ? (double)xs[i + u * 256] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s0 += v[u];
      s1 += v[u] * (double)(i + u * 256);
    }
  }
  __shared__ double r0[256], r1[256];
  r0[threadIdx.x] = s0;
  r1[threadIdx.x] = s1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { r0[threadIdx.x] += r0[threadIdx.x + o]; r1[threadIdx.x] += r1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[((int64_t)sig * blocks + b) * 2 + 0] = r0[0];
    partial[((int64_t)sig * blocks + b) * 2 + 1] = r1[0];
  }
}

template <class T>
__global__ void __launch_bounds__(32) stats_final(const T* __restrict__ x, int64_t n, int64_t stride, int blocks,
                                                  const double* __restrict__ partial, double* __restrict__ stats) {
  const int sig = blockIdx.x, lane = threadIdx.x;
  double s0 = 0.0, s1 = 0.0;
  for (int b = lane; b < blocks; b += 32) {      // fixed association: bit-reproducible
    s0 += partial[((int64_t)sig * blocks + b) * 2 + 0];
    s1 += partial[((int64_t)sig * blocks + b) * 2 + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_down_sync(0xffffffffu, s0, o);
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
  }
  if (lane == 0) {
    const T* xs = x + (int64_t)sig * stride;
    stats[sig * 4 + 0] = s0;
    stats[sig * 4 + 1] = s1;
    stats[sig * 4 + 2] = (double)xs[0];
    stats[sig * 4 + 3] = (double)xs[n - 1];
  }
}

// ----------------------------------------------------------------- pointwise
template <class TI, class TO>
__global__ void __launch_bounds__(256) pointwise_kernel(const TI* __restrict__ x, int64_t xs, TO* __restrict__ y,
                                                        int64_t ys, int64_t n, PointwiseDev pw,
                                                        const double* __restrict__ stats) {
  const int sig = blockIdx.y;
  double c0, c1;
  trend_coefficients(pw, stats ? stats + sig * 4 : nullptr, n, c0, c1);
  const TI* xi = x + (int64_t)sig * xs;
  TO* yo = y + (int64_t)sig * ys;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    yo[i] = (TO)pointwise_value(pw, (double)xi[i], i, n, c0, c1);
}

// ------------------------------------------------------------------- sosfilt
template <int S> struct SosCoef { double c[S][5]; };     // b0 b1 b2 a1 a2

// One sample through the cascade (scipy _sosfilt recurrence, DF2T):
//   y = b0*x + z0;   z0 = b1*x - a1*y + z1;   z1 = b2*x - a2*y
// with fused multiply-adds (5 per section).  Against scipy's separately rounded
// products this differs at the 1e-14 level of the output (measured; a bit-exact
// non-fused variant gave the same parity at twice the FP64 work).
template <int S> struct SosRegs { double c[S][5]; };

// Coefficients go from the kernel-parameter bank into ordinary registers through
// an opaque generic load.  Left to itself the compiler keeps them in the 63-entry
// uniform register file and spills/fills that file on every sample (measured: 176
// instructions per sample instead of ~60, ALU pipe 82 % busy with moves).
template <int S> __device__ __forceinline__ void load_coefficients(const SosCoef<S>& f, SosRegs<S>& r) {
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      double v;
      asm volatile("ld.f64 %0, [%1];" : "=d"(v) : "l"(&f.c[s][k]));
      r.c[s][k] = v;
    }
}

// Uses every loaded coefficient (and start-state) register ONCE before the sample loop.  Their first use would
// otherwise be a multiply-add inside the loop, which then carries the wait on the scoreboard of those loads for
// good -- and the scoreboard is the one the loop's own batch loads signal, so every iteration waited for the
// batch it had just requested (ncu source page: 25 % of the stall samples of phase 1 on one DFMA,
// `long_scoreboard`).  The sum is NaN only if a coefficient or state is NaN / the states hold opposite
// infinities -- then every output of the chunk is NaN anyway, which is what seeding z[0] with it gives.
template <int S> __device__ __forceinline__ void settle_loaded_values(const SosRegs<S>& f, double (&z)[2 * S]) {
  double acc = 0.0;
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int k = 0; k < 5; ++k) acc += f.c[s][k];
#pragma unroll
  for (int k = 0; k < 2 * S; ++k) acc += z[k];
  if (acc != acc) z[0] = acc;
}

template <int S> __device__ __forceinline__ double cascade(const SosRegs<S>& f, double (&z)[2 * S], double v) {
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const double y = fma(f.c[s][0], v, z[2 * s]);
    z[2 * s] = fma(-f.c[s][3], y, fma(f.c[s][1], v, z[2 * s + 1]));
    z[2 * s + 1] = fma(-f.c[s][4], y, f.c[s][2] * v);
    v = y;
  }
  return v;
}

constexpr int kBatchConst = 8;
// ---- batches of kBatch consecutive samples of one thread's chunk, in processing order (dir = +1 / -1) ----
template <class T> struct Vec16;
template <> struct Vec16<float> { using type = float4; static constexpr int V = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int V = 2; };
template <class T> __device__ __forceinline__ void vec_unpack(const float4& t, T (&e)[4]) { e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w; }
template <class T> __device__ __forceinline__ void vec_unpack(const double2& t, T (&e)[2]) { e[0] = t.x; e[1] = t.y; }
__device__ __forceinline__ float4 vec_pack(const float (&e)[4]) { return make_float4(e[0], e[1], e[2], e[3]); }
__device__ __forceinline__ double2 vec_pack(const double (&e)[2]) { return make_double2(e[0], e[1]); }

// is the batch that starts at p (first sample in processing order) one run of aligned 16-byte vectors?
template <class T> __device__ __forceinline__ bool batch_aligned(const T* p, int dir) {
  const uintptr_t lo = reinterpret_cast<uintptr_t>(dir > 0 ? p : p - (kBatchConst - 1));
  return p != nullptr && (lo & 15u) == 0;
}
// The batch comes back in MEMORY order (d[k] = lowest address + k) whatever the direction: the caller picks
// element b of the processing order when it consumes it.  (Reordering here put a select on every loaded
// register right behind the load, i.e. the wait for the NEXT batch in front of the current batch's arithmetic:
// a quarter of all stall samples of phase 1.)
template <class T>
__device__ __forceinline__ void load_batch(const T* __restrict__ p, int dir, bool vec, T (&d)[kBatchConst]) {
  constexpr int V = Vec16<T>::V;
  using VT = typename Vec16<T>::type;
  const T* lo = dir > 0 ? p : p - (kBatchConst - 1);
  if (vec) {
    const VT* q = reinterpret_cast<const VT*>(lo);
#pragma unroll
    for (int k = 0; k < kBatchConst / V; ++k) {
      T e[V];
      vec_unpack(__ldg(q + k), e);
#pragma unroll
      for (int m = 0; m < V; ++m) d[k * V + m] = e[m];
    }
  } else {
#pragma unroll
    for (int b = 0; b < kBatchConst; ++b) d[b] = __ldg(lo + b);
  }
}
template <class T> __device__ __forceinline__ T batch_elem(const T (&d)[kBatchConst], int dir, int b) {
  return dir > 0 ? d[b] : d[kBatchConst - 1 - b];
}
template <class T>
__device__ __forceinline__ void store_batch(T* __restrict__ p, int dir, bool vec, const T (&d)[kBatchConst]) {
  constexpr int V = Vec16<T>::V;
  using VT = typename Vec16<T>::type;
  if (vec) {
    VT* q = reinterpret_cast<VT*>(dir > 0 ? p : p - (kBatchConst - 1));
#pragma unroll
    for (int k = 0; k < kBatchConst / V; ++k) {
      T e[V];
#pragma unroll
      for (int m = 0; m < V; ++m) e[m] = dir > 0 ? d[k * V + m] : d[kBatchConst - 1 - (k * V + m)];
      q[k] = vec_pack(e);
    }
  } else {
#pragma unroll
    for (int b = 0; b < kBatchConst; ++b) p[dir * b] = d[b];
  }
}

constexpr int kPrefetch = 96;   // samples ahead of the current batch that the scan asks into L2
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// resident CTAs the scan kernels are compiled for: the register allocation of the long cascades flips between
// 148 and 188 registers with unrelated edits; three CTAs (<= 168 registers) is what the 10-section chain needs
template <int S> constexpr int kScanMinBlocks = S >= 11 ? 2 : (S >= 5 ? 3 : 4);
constexpr int kIlp = 1;   // chunks per thread (the dispatch below sizes its grids with it)
constexpr int kBatch = kBatchConst; // samples fetched ahead per thread: the loads leave the serial recurrence's critical path

// One chunk of one signal through the cascade, from state z.  Everything that does not change inside a chunk
// is decided once per chunk so that a sample costs its 5 S multiply-adds plus a handful of instructions (the
// first version re-derived 64-bit indices, bounds, taper membership and `t / decimate` per sample: 160
// instructions per sample, two thirds of them integer -- profiles/ncu_r02_preprocess.md):
//   * 32-bit sample indices (a trace holds < 2^31 samples), a running double for the trend's abscissa;
//   * chunks that touch no taper ramp (all but ~15 per trace) skip the taper test;
//   * full batches of kBatch samples run unguarded, the ragged end of the last chunk separately;
//   * kept (decimated) samples are tracked with a counter instead of a division.
// MODE 1: phase 1 -- zero-state pass; writes the kept zero-state outputs to yzs when given.
// MODE 3: phase 3 -- true start state; stores the (kept) outputs to y.
template <int S, class TI, class TO, int MODE, bool RAMPS>
__device__ __forceinline__ void scan_chunk_impl(const TI* __restrict__ xi, int n, int t0, int len,
                                                const SosRegs<S>& fr, double (&z)[2 * S], const PointwiseDev& pw,
                                                double c0, double c1, int reverse, int decim,
                                                double* __restrict__ yz, TO* __restrict__ yo) {
  const int dir = reverse ? -1 : 1;
  const int i = reverse ? n - 1 - t0 : t0;                 // signal index of the chunk's first sample
  const TI* px = xi + i;
  const bool trend = pw.detrend != 0;
  double di = (double)i;
  const double ddir = (double)dir;
  int next_keep = ((t0 + decim - 1) / decim) * decim - t0;  // local index of the first kept sample (once per chunk)
  int64_t ko = ((int64_t)t0 + decim - 1) / decim;           // its output index
  auto sample = [&](double v, double w, int jl) -> double {
    if (trend) v -= fma(c1, di, c0);
    if (RAMPS) v *= w;
    di += ddir;
    const double y = cascade<S>(fr, z, v);
    if (MODE == 1) {
      if (yz && jl == next_keep) { yz[ko++] = y; next_keep += decim; }
    } else if (decim != 1) {                                  // (decim == 1: the caller stores whole batches)
      if (jl == next_keep) { yo[ko++] = (TO)y; next_keep += decim; }
    }
    return y;
  };
  // taper weight of signal index ii (1 between the ramps).  The weights of a batch are fetched one batch ahead
  // like the samples: read inside `sample` they put a dependent L2 load in front of every cascade step, the ~20
  // ramp chunks of a trace ran several times longer than the others, and the one at the very end of the grid
  // (right ramp of the last trace) held the whole launch: most SMs idle for the last quarter of phase 1 (ncu:
  // sm__cycles_active avg 0.75 of max).
  auto taper_w = [&](int ii) -> double {
    if (ii < pw.taper_left) return __ldg(pw.ramp_left + ii);
    if (ii >= n - pw.taper_right) return __ldg(pw.ramp_right + (ii - (n - pw.taper_right)));
    return 1.0;
  };
  // Batches move as 16-byte vectors when the chunk's first batch is 16-byte aligned in its direction (chunk
  // lengths are multiples of kBatch, so then every full batch is).  A thread walks its own chunk, so a warp-wide
  // load touches 32 different lines either way: one L1 wavefront per lane and instruction -- with scalar loads
  // that was one wavefront per SAMPLE and lane, and the L1 data stage (not FP64, not HBM) bound phases 1 and 3:
  // 67 % `L1 wavefronts`, long_scoreboard the top stall, the 4-section cascade as slow as the 10-section one
  // (profiles/ncu_r02_preprocess.md).
  const bool vec_in = batch_aligned<TI>(px, dir);
  TO* const py = (MODE == 3 && decim == 1) ? yo + i : nullptr;
  const bool vec_out = MODE == 3 && decim == 1 && batch_aligned<TO>(py, dir);
  TI buf[kBatch], nxt[kBatch];
  double wb[RAMPS ? kBatch : 1], wn[RAMPS ? kBatch : 1];
  const int full = len / kBatch * kBatch;
  if (full > 0) {
    load_batch<TI>(px, dir, vec_in, buf);
    if (RAMPS) {
#pragma unroll
      for (int b = 0; b < kBatch; ++b) wb[RAMPS ? b : 0] = taper_w(i + dir * b);
    }
  }
  for (int j0 = 0; j0 < full; j0 += kBatch) {
    // The lines a thread will need ~100 samples from now are requested into L2 (no register cost): with the next
    // batch as the only load in flight a quarter of all stall samples sat on its first use (3 warps per scheduler
    // and 32 different DRAM pages per warp-wide load: the latency exceeds the ~1000 cycles a batch computes).
    if ((j0 & (2 * kBatch - 1)) == 0 && j0 + kPrefetch < len) prefetch_l2(px + dir * (j0 + kPrefetch));
    if (j0 + kBatch < full) {
      load_batch<TI>(px + dir * (j0 + kBatch), dir, vec_in, nxt);
      if (RAMPS) {
#pragma unroll
        for (int b = 0; b < kBatch; ++b) wn[RAMPS ? b : 0] = taper_w(i + dir * (j0 + kBatch + b));
      }
    }
    if (MODE == 3 && decim == 1) {
      TO outv[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b) outv[b] = (TO)sample((double)batch_elem(buf, dir, b), wb[RAMPS ? b : 0], j0 + b);
      store_batch<TO>(py + dir * j0, dir, vec_out, outv);
    } else {
#pragma unroll
      for (int b = 0; b < kBatch; ++b) sample((double)batch_elem(buf, dir, b), wb[RAMPS ? b : 0], j0 + b);
    }
#pragma unroll
    for (int b = 0; b < kBatch; ++b) buf[b] = nxt[b];
    if (RAMPS) {
#pragma unroll
      for (int b = 0; b < kBatch; ++b) wb[RAMPS ? b : 0] = wn[RAMPS ? b : 0];
    }
  }
  for (int jl = full; jl < len; ++jl) {
    const double y = sample((double)__ldg(px + dir * jl), RAMPS ? taper_w(i + dir * jl) : 1.0, jl);
    if (MODE == 3 && decim == 1) py[dir * jl] = (TO)y;
  }
}

template <int S, class TI, class TO, int MODE>
__device__ __forceinline__ void scan_chunk(const TI* __restrict__ xi, int n, int t0, int len, const SosRegs<S>& fr,
                                           double (&z)[2 * S], const PointwiseDev& pw, double c0, double c1,
                                           int reverse, int decim, double* __restrict__ yz, TO* __restrict__ yo) {
  // does the chunk touch a taper ramp?
  const int i = reverse ? n - 1 - t0 : t0;
  const int i_lo = reverse ? i - (len - 1) : i, i_hi = reverse ? i : i + (len - 1);
  const bool ramps = i_lo < pw.taper_left || i_hi >= n - pw.taper_right;
  if (ramps) scan_chunk_impl<S, TI, TO, MODE, true>(xi, n, t0, len, fr, z, pw, c0, c1, reverse, decim, yz, yo);
  else scan_chunk_impl<S, TI, TO, MODE, false>(xi, n, t0, len, fr, z, pw, c0, c1, reverse, decim, yz, yo);
}

// phase 1: zero-state final state of every chunk (one thread per chunk)
template <int S, class TI>
__global__ void __launch_bounds__(128, kScanMinBlocks<S>) sos_phase1(const TI* __restrict__ x, int64_t n, int64_t xs, int n_signals,
                                                  int chunk, int64_t n_chunks, const __grid_constant__ SosCoef<S> f,
                                                  PointwiseDev pw, const double* __restrict__ stats, int reverse,
                                                  double* __restrict__ zstate, double* __restrict__ yzs, int decim,
                                                  int64_t n_out) {
  const int64_t gid = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (gid >= (int64_t)n_signals * n_chunks) return;
  const int sig = (int)(gid / n_chunks);
  const int c = (int)(gid - (int64_t)sig * n_chunks);
  double c0, c1;
  trend_coefficients(pw, stats ? stats + sig * 4 : nullptr, n, c0, c1);
  SosRegs<S> fr;
  load_coefficients<S>(f, fr);
  double z[2 * S];
#pragma unroll
  for (int k = 0; k < 2 * S; ++k) z[k] = 0.0;
  settle_loaded_values<S>(fr, z);
  const int t0 = c * chunk;
  const int len = (int)n - t0 < chunk ? (int)n - t0 : chunk;
  scan_chunk<S, TI, double, 1>(x + (int64_t)sig * xs, (int)n, t0, len, fr, z, pw, c0, c1, reverse, decim,
                               yzs ? yzs + (int64_t)sig * n_out : nullptr, (double*)nullptr);
  double* out = zstate + ((int64_t)sig * n_chunks + c) * (2 * S);
#pragma unroll
  for (int k = 0; k < 2 * S; ++k) out[k] = z[k];
}

// phase 3: rerun every chunk from its true start state and store
template <int S, class TI, class TO>
__global__ void __launch_bounds__(128, kScanMinBlocks<S>) sos_phase3(const TI* __restrict__ x, int64_t n, int64_t xs, TO* __restrict__ y,
                                                  int64_t ys, int n_signals, int chunk, int64_t n_chunks,
                                                  const __grid_constant__ SosCoef<S> f, PointwiseDev pw,
                                                  const double* __restrict__ stats, int reverse, int decim,
                                                  const double* __restrict__ sstart) {
  const int64_t gid = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (gid >= (int64_t)n_signals * n_chunks) return;
  const int sig = (int)(gid / n_chunks);
  const int c = (int)(gid - (int64_t)sig * n_chunks);
  double c0, c1;
  trend_coefficients(pw, stats ? stats + sig * 4 : nullptr, n, c0, c1);
  SosRegs<S> fr;
  load_coefficients<S>(f, fr);
  double z[2 * S];
  const double* in = sstart + ((int64_t)sig * n_chunks + c) * (2 * S);
#pragma unroll
  for (int k = 0; k < 2 * S; ++k) z[k] = in[k];
  settle_loaded_values<S>(fr, z);
  const int t0 = c * chunk;
  const int len = (int)n - t0 < chunk ? (int)n - t0 : chunk;
  scan_chunk<S, TI, TO, 3>(x + (int64_t)sig * xs, (int)n, t0, len, fr, z, pw, c0, c1, reverse, decim,
                           (double*)nullptr, y + (int64_t)sig * ys);
}

// phase 3 of a DECIMATING cascade: by linearity the true output is the zero-state output phase 1
// kept (yzs) plus the zero-input response of the chunk's start state, y[r] = yzs[r] + sum_k G[r][k] s0[k],
// with G the response of the kept local samples to unit start states (host: scipy, like the carry
// matrices).  d multiply-adds per kept sample instead of rerunning 5 S per input sample.
// A block takes kZinChunks consecutive chunks of one signal: their kept outputs are contiguous, so
// the threads (one per kept local sample) walk them coalesced.
constexpr int kZinChunks = 64;
template <class TO>
__global__ void __launch_bounds__(128) sos_phase3_zin(const double* __restrict__ yzs, TO* __restrict__ y, int64_t ys,
                                                      int64_t n_out, int rows, int64_t n_chunks, int d,
                                                      const double* __restrict__ G,
                                                      const double* __restrict__ sstart) {
  // A thread owns ONE kept local sample r: its row of G stays in registers (zero padded to 24) while the block
  // walks kZinChunks chunks, whose start states sit in shared memory and are read as warp-wide broadcasts.
  // (First version: G and the states both in shared memory -- two shared loads per multiply-add, 222 us for
  // 308 MB, short_scoreboard the top stall; profiles/ncu_r02_preprocess.md.)
  extern __shared__ double zin_sm[];                     // s0 [kZinChunks][24]
  const int64_t blocks_per_sig = (n_chunks + kZinChunks - 1) / kZinChunks;
  const int sig = (int)(blockIdx.x / blocks_per_sig);
  const int64_t c0 = (blockIdx.x - (int64_t)sig * blocks_per_sig) * kZinChunks;
  const int nc = (int)(n_chunks - c0 < kZinChunks ? n_chunks - c0 : kZinChunks);
  const double* in = sstart + ((int64_t)sig * n_chunks + c0) * d;
  for (int i = threadIdx.x; i < nc * 24; i += blockDim.x) {
    const int c = i / 24, k = i - c * 24;
    zin_sm[i] = k < d ? in[c * d + k] : 0.0;
  }
  __syncthreads();
  const double* src = yzs + (int64_t)sig * n_out;
  TO* dst = y + (int64_t)sig * ys;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    double g[24];
#pragma unroll
    for (int k = 0; k < 24; ++k) g[k] = k < d ? __ldg(G + (int64_t)r * d + k) : 0.0;
    // zero-state outputs are fetched kZinAhead chunks ahead (one ahead left the loads exposed: long_scoreboard
    // 6.5 warps per issue at 14 resident warps per SM)
    constexpr int kZinAhead = 4;
    const int64_t o0 = c0 * rows + r;
    double cur[kZinAhead], nxt[kZinAhead];
#pragma unroll
    for (int u = 0; u < kZinAhead; ++u) {
      const int64_t o = o0 + (int64_t)u * rows;
      cur[u] = (u < nc && o < n_out) ? src[o] : 0.0;
    }
    for (int cb = 0; cb < nc; cb += kZinAhead) {
#pragma unroll
      for (int u = 0; u < kZinAhead; ++u) {
        const int64_t o = o0 + (int64_t)(cb + kZinAhead + u) * rows;
        nxt[u] = (cb + kZinAhead + u < nc && o < n_out) ? src[o] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kZinAhead; ++u) {
        const int c = cb + u;
        const int64_t o = o0 + (int64_t)c * rows;
        if (c < nc && o < n_out) {
          const double2* s0 = reinterpret_cast<const double2*>(zin_sm + c * 24);
          double a0 = cur[u], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
          for (int k = 0; k < 24; k += 4) {
            const double2 p0 = s0[k / 2], p1 = s0[k / 2 + 1];
            a0 = fma(g[k], p0.x, a0);
            a1 = fma(g[k + 1], p0.y, a1);
            a2 = fma(g[k + 2], p1.x, a2);
            a3 = fma(g[k + 3], p1.y, a3);
          }
          dst[o] = (TO)((a0 + a1) + (a2 + a3));
        }
      }
#pragma unroll
      for (int u = 0; u < kZinAhead; ++u) cur[u] = nxt[u];
    }
  }
}

// phase 2: chunk start states, s_{c+1} = P s_c + z_c (d = 2S <= 24).  Lane r owns
// row r of the transition matrix (registers) and component r of the state; the
// state is broadcast with shuffles, the row product runs as four independent
// partial sums, and the next z is prefetched.
__device__ __forceinline__ double carry_step(const double (&prow)[24], double s, double add, int d) {
  double a0 = add, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int k = 0; k < 24; k += 4) {
    const double s0 = __shfl_sync(0xffffffffu, s, k), s1 = __shfl_sync(0xffffffffu, s, k + 1);
    const double s2 = __shfl_sync(0xffffffffu, s, k + 2), s3 = __shfl_sync(0xffffffffu, s, k + 3);
    a0 = fma(prow[k], s0, a0);          // rows are zero padded beyond d
    a1 = fma(prow[k + 1], s1, a1);
    a2 = fma(prow[k + 2], s2, a2);
    a3 = fma(prow[k + 3], s3, a3);
  }
  (void)d;
  return (a0 + a1) + (a2 + a3);
}

// One warp per (signal, group).
// mode 0: fold   e_g = state after the group's chunks from a zero start -> gstate[sig][g]
// mode 2: expand sstart[c] for the group's chunks from gstate[sig][g] (group start state)
__global__ void __launch_bounds__(32) sos_carry_group(const double* __restrict__ zstate, int d, int64_t n_chunks,
                                                      int64_t n_groups, const double* __restrict__ P, int mode,
                                                      double* __restrict__ gstate, double* __restrict__ sstart) {
  const int lane = threadIdx.x;
  const int64_t sg = blockIdx.x;                 // sig * n_groups + g
  const int64_t sig = sg / n_groups, g = sg - sig * n_groups;
  double prow[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) prow[k] = (lane < d && k < d) ? P[lane * d + k] : 0.0;
  double s = (mode == 2 && lane < d) ? gstate[sg * d + lane] : 0.0;
  const int64_t c0 = g * kGroup;
  int64_t c1 = c0 + kGroup;
  if (c1 > n_chunks) c1 = n_chunks;
  const double* zb = zstate + (sig * n_chunks) * d;
  double* sb = sstart + (sig * n_chunks) * d;
  // z is fetched kPf steps ahead: a step is ~100 cycles of arithmetic, DRAM latency several hundred
  constexpr int kPf = 8;
  double cur[kPf], nxt[kPf];
#pragma unroll
  for (int u = 0; u < kPf; ++u) cur[u] = (lane < d && c0 + u < c1) ? zb[(c0 + u) * d + lane] : 0.0;
  for (int64_t cb = c0; cb < c1; cb += kPf) {
#pragma unroll
    for (int u = 0; u < kPf; ++u) nxt[u] = (lane < d && cb + kPf + u < c1) ? zb[(cb + kPf + u) * d + lane] : 0.0;
#pragma unroll
    for (int u = 0; u < kPf; ++u) {
      if (cb + u < c1) {
        if (mode == 2 && lane < d) sb[(cb + u) * d + lane] = s;
        s = carry_step(prow, s, cur[u], d);
      }
    }
#pragma unroll
    for (int u = 0; u < kPf; ++u) cur[u] = nxt[u];
  }
  if (mode == 0 && lane < d) gstate[sg * d + lane] = s;
}

// Start states of a fast-decaying cascade in one launch (xcb_sos::carry_window = K): one warp per (signal, run
// of kRun chunks) starts K chunks before its run -- from the zero state, or from `init` when that is the start of
// the trace -- and steps s <- P s + z through warm-up and run, storing the run's start states.  What the zero
// start drops is P^K s with ||P^K|| <= 2^-80.  (cfg2: fold 56 us + sweep 28 us + expand 58 us of mostly idle
// SMs -- 830 warps x 256 serial steps -- become one launch of 7000 warps x ~40 steps.)
constexpr int kRun = 32;
__global__ void __launch_bounds__(32) sos_carry_window(const double* __restrict__ zstate, int d, int64_t n_chunks,
                                                       int64_t n_runs, int window, const double* __restrict__ P,
                                                       const double* __restrict__ init, double* __restrict__ sstart) {
  __shared__ __align__(16) double sh[32];
  const int lane = threadIdx.x;
  const int64_t sr = blockIdx.x;                 // sig * n_runs + run
  const int64_t sig = sr / n_runs, r = sr - sig * n_runs;
  double prow[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) prow[k] = (lane < d && k < d) ? P[lane * d + k] : 0.0;
  const int64_t r0 = r * kRun;
  int64_t r1 = r0 + kRun;
  if (r1 > n_chunks) r1 = n_chunks;
  const int64_t c0 = r0 > window ? r0 - window : 0;
  double s = (c0 == 0 && init && lane < d) ? init[sig * d + lane] : 0.0;
  const double* zb = zstate + (sig * n_chunks) * d;
  double* sb = sstart + (sig * n_chunks) * d;
  constexpr int kPf = 8;
  double cur[kPf], nxt[kPf];
#pragma unroll
  for (int u = 0; u < kPf; ++u) cur[u] = (lane < d && c0 + u < r1) ? zb[(c0 + u) * d + lane] : 0.0;
  for (int64_t cb = c0; cb < r1; cb += kPf) {
#pragma unroll
    for (int u = 0; u < kPf; ++u) nxt[u] = (lane < d && cb + kPf + u < r1) ? zb[(cb + kPf + u) * d + lane] : 0.0;
#pragma unroll
    for (int u = 0; u < kPf; ++u) {
      if (cb + u < r1) {
        if (cb + u >= r0 && lane < d) sb[(cb + u) * d + lane] = s;
        // the state goes round through shared memory and is read back as 16-byte broadcasts (13 wavefronts per
        // step; the shuffle form of carry_step issues 48 and bound this launch at 87 % L1 wavefronts)
        sh[lane] = s;
        __syncwarp();
        double a0 = cur[u], a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const double2* s2 = reinterpret_cast<const double2*>(sh);
#pragma unroll
        for (int k = 0; k < 24; k += 4) {
          const double2 p0 = s2[k / 2], p1 = s2[k / 2 + 1];
          a0 = fma(prow[k], p0.x, a0);          // rows are zero padded beyond d
          a1 = fma(prow[k + 1], p0.y, a1);
          a2 = fma(prow[k + 2], p1.x, a2);
          a3 = fma(prow[k + 3], p1.y, a3);
        }
        s = (a0 + a1) + (a2 + a3);
        __syncwarp();
      }
    }
#pragma unroll
    for (int u = 0; u < kPf; ++u) cur[u] = nxt[u];
  }
}

// sequential sweep over groups (one warp per signal): gstate holds e_g on entry,
// the group START states on exit; PG = P^kGroup.
__global__ void __launch_bounds__(32) sos_carry_sweep(int d, int64_t n_groups, const double* __restrict__ PG,
                                                      double* __restrict__ gstate, const double* __restrict__ init) {
  const int lane = threadIdx.x;
  const int64_t sig = blockIdx.x;
  double prow[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) prow[k] = (lane < d && k < d) ? PG[lane * d + k] : 0.0;
  double s = (init && lane < d) ? init[sig * d + lane] : 0.0;   // sosfilt(..., zi=...): state before the first sample
  double* base = gstate + sig * n_groups * d;
  constexpr int kPf = 8;
  double cur[kPf], nxt[kPf];
#pragma unroll
  for (int u = 0; u < kPf; ++u) cur[u] = (lane < d && u < n_groups) ? base[u * d + lane] : 0.0;
  for (int64_t gb = 0; gb < n_groups; gb += kPf) {
#pragma unroll
    for (int u = 0; u < kPf; ++u) nxt[u] = (lane < d && gb + kPf + u < n_groups) ? base[(gb + kPf + u) * d + lane] : 0.0;
#pragma unroll
    for (int u = 0; u < kPf; ++u) {
      if (gb + u < n_groups) {
        if (lane < d) base[(gb + u) * d + lane] = s;        // start state of group gb+u (its e was read already)
        s = carry_step(prow, s, cur[u], d);
      }
    }
#pragma unroll
    for (int u = 0; u < kPf; ++u) cur[u] = nxt[u];
  }
}

// ------------------------------------------------------ polyphase FIR (upfirdn)
// y[k] = sum_i x[i] h[(k + k0) down - i up]: the zero-stuffed input is never formed and only every
// down-th output is computed -- `n_taps / up` multiply-adds per output (scipy.signal.upfirdn /
// resample_poly).  A block owns kFirTile consecutive outputs of one signal: the input span they touch
// (with the optional detrend / taper prologue applied) and the taps sit in shared memory as float64.
constexpr int kFirTile = 512;
template <class TI, class TO>
__global__ void __launch_bounds__(256) upfirdn_kernel(const TI* __restrict__ x, int64_t xs, int64_t n_in,
                                                      TO* __restrict__ y, int64_t ys, int64_t n_out,
                                                      const double* __restrict__ taps, int n_taps, int up, int down,
                                                      int64_t k0, int span, PointwiseDev pw,
                                                      const double* __restrict__ stats) {
  extern __shared__ double fir_sm[];                     // taps [n_taps], then the input span
  double* h = fir_sm;
  double* xw = fir_sm + n_taps;
  const int sig = blockIdx.y;
  const int64_t ka = (int64_t)blockIdx.x * kFirTile;
  int64_t kb = ka + kFirTile;
  if (kb > n_out) kb = n_out;
  // first input any output of the tile can touch: ceil(((ka + k0) down - (n_taps - 1)) / up), floor division safe
  const int64_t qa = (ka + k0) * (int64_t)down - (n_taps - 1);
  const int64_t i_lo = qa >= 0 ? (qa + up - 1) / up : -((-qa) / up);
  double c0, c1;
  trend_coefficients(pw, stats ? stats + sig * 4 : nullptr, n_in, c0, c1);
  const TI* xi = x + (int64_t)sig * xs;
  for (int t = threadIdx.x; t < n_taps; t += 256) h[t] = taps[t];
  for (int t = threadIdx.x; t < span; t += 256) {
    const int64_t i = i_lo + t;
    xw[t] = (i >= 0 && i < n_in) ? pointwise_value(pw, (double)xi[i], i, n_in, c0, c1) : 0.0;
  }
  __syncthreads();
  TO* yo = y + (int64_t)sig * ys;
  for (int64_t k = ka + threadIdx.x; k < kb; k += 256) {
    const int64_t q = (k + k0) * (int64_t)down;          // position in the zero-stuffed input
    const int64_t i_hi = q / up;
    int m = (int)(q - i_hi * up);                        // tap that meets x[i_hi]
    const double* xp = xw + (i_hi - i_lo);
    double a0 = 0.0, a1 = 0.0;
    int t = 0;
    for (; m + up < n_taps; m += 2 * up, t += 2) {       // xw is zero outside the signal: no bounds in the loop
      a0 = fma(h[m], xp[-t], a0);
      a1 = fma(h[m + up], xp[-t - 1], a1);
    }
    if (m < n_taps) a0 = fma(h[m], xp[-t], a0);
    yo[k] = (TO)(a0 + a1);
  }
}

// ------------------------------------------------------------------ dispatch
struct SosArgs {
  const void* x; int64_t xs; void* y; int64_t ys; int64_t n; int n_signals;
  int chunk; int64_t n_chunks; PointwiseDev pw; const double* stats; int reverse, decim;
  double* zstate; double* sstart; const double (*coef)[5];
  double* yzs; int64_t n_out; const double* zin;     // decimating fast path (see sos_phase3_zin)
};

// FIR decimation (up == 1: resample_poly(x, 1, down), the north-star decimator), register tiled.
// y[k] = sum_m h[m] x[(k + k0) down - m].  A thread owns kFirR CONSECUTIVE outputs: the input x[c - p - down q]
// (c = (k + k0) down its first output's position, p = m mod down) meets output r with tap (q + r) of polyphase
// component p, so the thread walks each component once with a sliding window of kFirR taps in registers -- one
// shared load of x and one broadcast load of a tap per kFirR multiply-adds (the generic kernel above: two loads per
// multiply-add, 97 % L1 wavefronts, 860 us for 5 x 24 h @ 250 Hz -> 50 Hz).  The components sit in shared memory
// zero-padded by kFirR - 1 on both sides (no bounds in the loop); the input span is laid out with one pad word
// per kFirR * down samples so that the lanes' stride is odd (no bank conflicts).
constexpr int kFirR = 8;
#ifndef XCB_FIR_THREADS
#define XCB_FIR_THREADS 128
#endif
constexpr int kFirThreads = XCB_FIR_THREADS;
constexpr int kFirTileR = kFirR * kFirThreads;
template <class TI, class TO>
__global__ void __launch_bounds__(kFirThreads) fir_decimate_kernel(const TI* __restrict__ x, int64_t xs, int64_t n_in,
                                                                   TO* __restrict__ y, int64_t ys, int64_t n_out,
                                                                   const double* __restrict__ taps, int n_taps, int down,
                                                                   int64_t k0, int span, int qp, PointwiseDev pw,
                                                                   const double* __restrict__ stats) {
  extern __shared__ double fir_sm[];                     // components [down][qp], then the padded input span
  constexpr int R = kFirR;
  double* hp_all = fir_sm;
  double* xw = fir_sm + (size_t)down * qp;
  const int W = R * down;
  const int sig = blockIdx.y;
  const int64_t ka = (int64_t)blockIdx.x * kFirTileR;
  const int64_t i_lo = (ka + k0) * (int64_t)down - (n_taps - 1);      // input under the last tap of the tile's first output
  double c0, c1;
  trend_coefficients(pw, stats ? stats + sig * 4 : nullptr, n_in, c0, c1);
  const TI* xi = x + (int64_t)sig * xs;
  for (int t = threadIdx.x; t < down * qp; t += kFirThreads) {
    const int p = t / qp, tt = t - p * qp - (R - 1);
    const int m = p + down * tt;
    hp_all[t] = (tt >= 0 && m < n_taps) ? taps[m] : 0.0;
  }
  // The span is fetched kFill values per thread at a time, all loads issued before the first use, and a tile that
  // lies inside the trace and off the taper ramps (all but a handful) skips the per-sample bounds / ramp tests;
  // t / W by multiplication (t W < 2^32).  (First version: one load in flight per thread and ~100 instructions
  // per sample -- the fill, not the filter, took three quarters of the kernel: ncu source page.)
  constexpr int kFill = 8;
  const uint32_t w_magic = (uint32_t)((0x100000000ull + (uint32_t)W - 1) / (uint32_t)W);
  const bool interior = i_lo >= 0 && i_lo + span <= n_in && i_lo >= pw.taper_left &&
                        i_lo + span <= n_in - pw.taper_right;
  const bool trend = pw.detrend != 0;
  for (int t0 = threadIdx.x; t0 < span; t0 += kFirThreads * kFill) {
    TI v[kFill];
#pragma unroll
    for (int u = 0; u < kFill; ++u) {
      const int t = t0 + u * kFirThreads;
      const int64_t i = i_lo + t;
      v[u] = (t < span && (interior || (i >= 0 && i < n_in))) ? __ldg(xi + i) : TI(0);
    }
#pragma unroll
    for (int u = 0; u < kFill; ++u) {
      const int t = t0 + u * kFirThreads;
      const int64_t i = i_lo + t;
      if (t < span) {
        double val;
        if (interior) {
          val = (double)v[u];
          if (trend) val -= c0 + c1 * (double)i;
        } else {
          val = (i >= 0 && i < n_in) ? pointwise_value(pw, (double)v[u], i, n_in, c0, c1) : 0.0;
        }
        xw[t + (int)__umulhi((uint32_t)t, w_magic)] = val;
      }
    }
  }
  __syncthreads();
  const double* xb = xw + threadIdx.x * (W + 1);          // position t W (+ its pads) of this thread's first output
  double acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;
  for (int p = 0; p < down && p < n_taps; ++p) {
    const double* hp = hp_all + p * qp;
    const int steps = (n_taps - p + down - 1) / down + R - 1;
    double tw[R];
#pragma unroll
    for (int r = 0; r < R - 1; ++r) tw[r] = hp[r];
    int e = (n_taps - 1) - p + down * (R - 1);             // input position of step 0 relative to t W (same for all threads)
    int ediv = e / W, erem = e - ediv * W;
    int s = 0;
    for (; s + R <= steps; s += R) {
#pragma unroll
      for (int u = 0; u < R; ++u) {                        // the window rotates through the registers: no moves
        tw[(R - 1 + u) % R] = hp[s + u + R - 1];
        const double xv = xb[e + ediv];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fma(tw[(r + u) % R], xv, acc[r]);
        e -= down; erem -= down;
        if (erem < 0) { erem += W; --ediv; }
      }
    }
    for (; s < steps; ++s) {
      tw[R - 1] = hp[s + R - 1];
      const double xv = xb[e + ediv];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fma(tw[r], xv, acc[r]);
#pragma unroll
      for (int r = 0; r < R - 1; ++r) tw[r] = tw[r + 1];
      e -= down; erem -= down;
      if (erem < 0) { erem += W; --ediv; }
    }
  }
  TO* yo = y + (int64_t)sig * ys;
  const int64_t k = ka + (int64_t)threadIdx.x * R;
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (k + r < n_out) yo[k + r] = (TO)acc[r];
}

template <int S, class TI> cudaError_t run_phase1(const SosArgs& a, cudaStream_t st) {
  SosCoef<S> f;
  for (int s = 0; s < S; ++s) for (int k = 0; k < 5; ++k) f.c[s][k] = a.coef[s][k];
  const int64_t total = (int64_t)a.n_signals * ((a.n_chunks + kIlp - 1) / kIlp);
  sos_phase1<S, TI><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(static_cast<const TI*>(a.x), a.n, a.xs,
                                                                    a.n_signals, a.chunk, a.n_chunks, f, a.pw,
                                                                    a.stats, a.reverse, a.zstate, a.yzs, a.decim, a.n_out);
  return cudaGetLastError();
}

template <int S, class TI, class TO> cudaError_t run_phase3(const SosArgs& a, cudaStream_t st) {
  SosCoef<S> f;
  for (int s = 0; s < S; ++s) for (int k = 0; k < 5; ++k) f.c[s][k] = a.coef[s][k];
  const int64_t total = (int64_t)a.n_signals * ((a.n_chunks + kIlp - 1) / kIlp);
  sos_phase3<S, TI, TO><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
      static_cast<const TI*>(a.x), a.n, a.xs, static_cast<TO*>(a.y), a.ys, a.n_signals, a.chunk, a.n_chunks, f, a.pw,
      a.stats, a.reverse, a.decim, a.sstart);
  return cudaGetLastError();
}

template <int S> cudaError_t run_phase(int phase, int dti, int dto, const SosArgs& a, cudaStream_t st) {
  if (phase == 1) return dti == XCB_F32 ? run_phase1<S, float>(a, st) : run_phase1<S, double>(a, st);
  if (dti == XCB_F32) return dto == XCB_F32 ? run_phase3<S, float, float>(a, st) : run_phase3<S, float, double>(a, st);
  return dto == XCB_F32 ? run_phase3<S, double, float>(a, st) : run_phase3<S, double, double>(a, st);
}

cudaError_t run_phase_s(int S, int phase, int dti, int dto, const SosArgs& a, cudaStream_t st) {
  switch (S) {
    case 1: return run_phase<1>(phase, dti, dto, a, st);
    case 2: return run_phase<2>(phase, dti, dto, a, st);
    case 3: return run_phase<3>(phase, dti, dto, a, st);
    case 4: return run_phase<4>(phase, dti, dto, a, st);
    case 5: return run_phase<5>(phase, dti, dto, a, st);
    case 6: return run_phase<6>(phase, dti, dto, a, st);
    case 7: return run_phase<7>(phase, dti, dto, a, st);
    case 8: return run_phase<8>(phase, dti, dto, a, st);
    case 9: return run_phase<9>(phase, dti, dto, a, st);
    case 10: return run_phase<10>(phase, dti, dto, a, st);
    case 11: return run_phase<11>(phase, dti, dto, a, st);
    case 12: return run_phase<12>(phase, dti, dto, a, st);
  }
  return cudaErrorInvalidValue;
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct SosLayout { int64_t n_chunks, n_groups, n_out; int d; size_t off_P, off_PG, off_z, off_s, off_g, off_y, total; };

SosLayout sos_layout(int64_t n, int n_signals, const xcb_sos* f) {
  SosLayout L;
  L.d = 2 * f->n_sections;
  L.n_chunks = (n + f->chunk - 1) / f->chunk;
  L.n_groups = (L.n_chunks + kGroup - 1) / kGroup;
  size_t o = 0;
  L.off_P = o;  o += align256(sizeof(double) * L.d * L.d);
  L.off_PG = o; o += align256(sizeof(double) * L.d * L.d);
  L.off_z = o;  o += align256(sizeof(double) * (size_t)n_signals * L.n_chunks * L.d);
  L.off_s = o;  o += align256(sizeof(double) * (size_t)n_signals * L.n_chunks * L.d);
  L.off_g = o;  o += align256(sizeof(double) * (size_t)n_signals * L.n_groups * L.d);
  L.n_out = (n + f->decimate - 1) / f->decimate;
  L.off_y = o;
  if (f->zin_response) o += align256(sizeof(double) * (size_t)n_signals * L.n_out);
  L.total = o;
  return L;
}

PointwiseDev to_dev(const xcb_pointwise* p) {
  PointwiseDev d;
  memset(&d, 0, sizeof(d));
  if (p) {
    d.detrend = p->detrend;
    d.taper_left = p->taper_left;
    d.taper_right = p->taper_right;
    d.ramp_left = p->ramp_left;
    d.ramp_right = p->ramp_right;
  }
  return d;
}

bool dtype_ok(int d) { return d == XCB_F32 || d == XCB_F64; }

}  // namespace

extern "C" {

XCB_API int xcb_stats_workspace_bytes(int64_t n, int32_t n_signals, size_t* bytes) {
  if (!bytes || n < 1 || n_signals < 1) return XCB_ERR_ARG;
  *bytes = sizeof(double) * 2 * 1024 * (size_t)n_signals;
  return XCB_OK;
}

XCB_API int xcb_stats(int32_t dtype, const void* x, int64_t n, int64_t stride, int32_t n_signals, double* stats,
                      void* workspace, void* stream) {
  if (!x || !stats || !workspace || n < 1 || n_signals < 1 || stride < n || !dtype_ok(dtype)) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int blocks = (int)((n + 16383) / 16384);           // enough CTAs to fill the GPU for a few traces of a day at 50 Hz
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  double* partial = static_cast<double*>(workspace);
  dim3 grid((unsigned)blocks, (unsigned)n_signals);
  if (dtype == XCB_F32) {
    stats_partial<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), n, stride, blocks, partial);
    stats_final<float><<<n_signals, 32, 0, st>>>(static_cast<const float*>(x), n, stride, blocks, partial, stats);
  } else {
    stats_partial<double><<<grid, 256, 0, st>>>(static_cast<const double*>(x), n, stride, blocks, partial);
    stats_final<double><<<n_signals, 32, 0, st>>>(static_cast<const double*>(x), n, stride, blocks, partial, stats);
  }
  return cudaGetLastError() == cudaSuccess ? XCB_OK : XCB_ERR_CUDA;
}

XCB_API int xcb_pointwise_apply(int32_t dtype_in, const void* x, int64_t x_stride, int32_t dtype_out, void* y,
                                int64_t y_stride, int64_t n, int32_t n_signals, const xcb_pointwise* pw,
                                const double* stats, void* stream) {
  if (!x || !y || !pw || n < 1 || n_signals < 1 || !dtype_ok(dtype_in) || !dtype_ok(dtype_out)) return XCB_ERR_ARG;
  if (pw->detrend && !stats) return XCB_ERR_ARG;
  if (pw->taper_left + pw->taper_right > n) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PointwiseDev d = to_dev(pw);
  int bx = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (bx > 65535) bx = 65535;
  dim3 grid((unsigned)bx, (unsigned)n_signals);
  if (dtype_in == XCB_F32 && dtype_out == XCB_F32)
    pointwise_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, x_stride, (float*)y, y_stride, n, d, stats);
  else if (dtype_in == XCB_F32)
    pointwise_kernel<float, double><<<grid, 256, 0, st>>>((const float*)x, x_stride, (double*)y, y_stride, n, d, stats);
  else if (dtype_out == XCB_F32)
    pointwise_kernel<double, float><<<grid, 256, 0, st>>>((const double*)x, x_stride, (float*)y, y_stride, n, d, stats);
  else
    pointwise_kernel<double, double><<<grid, 256, 0, st>>>((const double*)x, x_stride, (double*)y, y_stride, n, d,
                                                          stats);
  return cudaGetLastError() == cudaSuccess ? XCB_OK : XCB_ERR_CUDA;
}

XCB_API int xcb_sosfilt_workspace_bytes(int64_t n, int32_t n_signals, const xcb_sos* f, size_t* bytes) {
  if (!f || !bytes || n < 1 || n_signals < 1 || f->n_sections < 1 || f->n_sections > kMaxSections || f->chunk < 1)
    return XCB_ERR_ARG;
  *bytes = sos_layout(n, n_signals, f).total;
  return XCB_OK;
}

XCB_API int xcb_sosfilt(int32_t dtype_in, const void* x, int64_t x_stride, int32_t dtype_out, void* y,
                        int64_t y_stride, int64_t n, int32_t n_signals, const xcb_sos* f, const xcb_pointwise* pre,
                        const double* stats, void* workspace, void* stream) {
  if (!x || !y || !f || !workspace || n < 1 || n_signals < 1 || !dtype_ok(dtype_in) || !dtype_ok(dtype_out))
    return XCB_ERR_ARG;
  if (f->n_sections < 1 || f->n_sections > kMaxSections || f->chunk < 1 || f->decimate < 1 || f->carry_window < 0)
    return XCB_ERR_ARG;
  if (!f->carry || !f->carry_group) return XCB_ERR_ARG;
  if (f->reverse && f->decimate != 1) return XCB_ERR_UNSUPPORTED;
  if (n >= 2147483647LL) return XCB_ERR_UNSUPPORTED;             // the scan indexes samples with 32 bits
  if (pre && pre->detrend && !stats) return XCB_ERR_ARG;
  if (pre && pre->taper_left + pre->taper_right > n) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const SosLayout L = sos_layout(n, n_signals, f);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  double* dP = reinterpret_cast<double*>(ws + L.off_P);
  double* dPG = reinterpret_cast<double*>(ws + L.off_PG);
  double* dz = reinterpret_cast<double*>(ws + L.off_z);
  double* ds = reinterpret_cast<double*>(ws + L.off_s);
  double* dg = reinterpret_cast<double*>(ws + L.off_g);
  // the carry matrices are device resident (uploaded once by the caller): no host copy, no implicit sync
  dP = const_cast<double*>(f->carry);
  dPG = const_cast<double*>(f->carry_group);
  SosArgs a;
  a.x = x; a.xs = x_stride; a.y = y; a.ys = y_stride; a.n = n; a.n_signals = n_signals;
  a.chunk = f->chunk; a.n_chunks = L.n_chunks; a.pw = to_dev(pre); a.stats = stats;
  a.reverse = f->reverse; a.decim = f->decimate; a.zstate = dz; a.sstart = ds; a.coef = f->coef;
  if (f->zin_response && (f->decimate < 2 || f->chunk % f->decimate != 0 || f->reverse)) return XCB_ERR_ARG;
  // the fast path keeps the response table in (default-sized) shared memory; else rerun the cascade
  const bool zin = f->zin_response != nullptr;
  a.yzs = zin ? reinterpret_cast<double*>(ws + L.off_y) : nullptr;
  a.n_out = L.n_out;
  a.zin = f->zin_response;
  if (run_phase_s(f->n_sections, 1, dtype_in, dtype_out, a, st) != cudaSuccess) return XCB_ERR_CUDA;
  if (f->carry_window >= 1 && L.n_chunks > 2 * kRun) {
    const int64_t n_runs = (L.n_chunks + kRun - 1) / kRun;
    sos_carry_window<<<(unsigned)((int64_t)n_signals * n_runs), 32, 0, st>>>(dz, L.d, L.n_chunks, n_runs,
                                                                          f->carry_window, dP, f->init_state, ds);
  } else {
    const unsigned ng = (unsigned)((int64_t)n_signals * L.n_groups);
    sos_carry_group<<<ng, 32, 0, st>>>(dz, L.d, L.n_chunks, L.n_groups, dP, 0, dg, ds);
    sos_carry_sweep<<<(unsigned)n_signals, 32, 0, st>>>(L.d, L.n_groups, dPG, dg, f->init_state);
    sos_carry_group<<<ng, 32, 0, st>>>(dz, L.d, L.n_chunks, L.n_groups, dP, 2, dg, ds);
  }
  if (cudaGetLastError() != cudaSuccess) return XCB_ERR_CUDA;
  if (zin) {
    const int rows = f->chunk / f->decimate;
    const int64_t blocks_per_sig = (L.n_chunks + kZinChunks - 1) / kZinChunks;
    const size_t smem = sizeof(double) * (size_t)kZinChunks * 24;
    const unsigned grid = (unsigned)((int64_t)n_signals * blocks_per_sig);
    // one thread per kept local sample: no idle lanes beyond the last partial warp (96 rows: 96 threads, not 128)
    const unsigned zin_threads = (unsigned)(rows >= 128 ? 128 : (rows + 31) / 32 * 32);
    if (dtype_out == XCB_F32)
      sos_phase3_zin<float><<<grid, zin_threads, smem, st>>>(a.yzs, static_cast<float*>(y), y_stride, L.n_out, rows,
                                                             L.n_chunks, L.d, a.zin, ds);
    else
      sos_phase3_zin<double><<<grid, zin_threads, smem, st>>>(a.yzs, static_cast<double*>(y), y_stride, L.n_out, rows,
                                                              L.n_chunks, L.d, a.zin, ds);
    return cudaGetLastError() == cudaSuccess ? XCB_OK : XCB_ERR_CUDA;
  }
  if (run_phase_s(f->n_sections, 3, dtype_in, dtype_out, a, st) != cudaSuccess) return XCB_ERR_CUDA;
  return XCB_OK;
}

XCB_API int xcb_upfirdn(int32_t dtype_in, const void* x, int64_t x_stride, int64_t n_in, int32_t dtype_out, void* y,
                        int64_t y_stride, int64_t n_out, int32_t n_signals, const double* taps, int32_t n_taps,
                        int32_t up, int32_t down, int64_t k0, const xcb_pointwise* pre, const double* stats,
                        void* stream) {
  if (!x || !y || !taps || n_in < 1 || n_out < 1 || n_signals < 1 || n_taps < 1 || up < 1 || down < 1 || k0 < 0 ||
      !dtype_ok(dtype_in) || !dtype_ok(dtype_out))
    return XCB_ERR_ARG;
  if (pre && pre->detrend && !stats) return XCB_ERR_ARG;
  if (pre && pre->taper_left + pre->taper_right > n_in) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (up == 1) {
    // decimation: register-tiled kernel (kFirR consecutive outputs per thread)
    const int span2 = (kFirTileR - 1) * down + n_taps;
    const int qp = (n_taps + down - 1) / down + 2 * (kFirR - 1);
    const size_t smem2 = sizeof(double) * ((size_t)down * qp + (size_t)span2 + span2 / (kFirR * down) + 1);
    if ((int64_t)(kFirTileR - 1) * down + n_taps < (1 << 30) && smem2 <= 160 * 1024) {
      dim3 grid2((unsigned)((n_out + kFirTileR - 1) / kFirTileR), (unsigned)n_signals);
      const PointwiseDev pw2 = to_dev(pre);
#define XCB_FIRD(TI, TO)                                                                                             \
  do {                                                                                                               \
    if (smem2 > 48 * 1024 &&                                                                                         \
        cudaFuncSetAttribute(fir_decimate_kernel<TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != \
            cudaSuccess)                                                                                             \
      return XCB_ERR_CUDA;                                                                                           \
    fir_decimate_kernel<TI, TO><<<grid2, kFirThreads, smem2, st>>>(static_cast<const TI*>(x), x_stride, n_in,        \
                                                                   static_cast<TO*>(y), y_stride, n_out, taps,      \
                                                                   n_taps, down, k0, span2, qp, pw2, stats);        \
  } while (0)
      if (dtype_in == XCB_F32 && dtype_out == XCB_F32) XCB_FIRD(float, float);
      else if (dtype_in == XCB_F32) XCB_FIRD(float, double);
      else if (dtype_out == XCB_F32) XCB_FIRD(double, float);
      else XCB_FIRD(double, double);
#undef XCB_FIRD
      return cudaGetLastError() == cudaSuccess ? XCB_OK : XCB_ERR_CUDA;
    }
  }
  // inputs touched by one tile of outputs: (kFirTile - 1) * down / up + n_taps / up + 2
  const int span = (int)(((int64_t)(kFirTile - 1) * down + n_taps - 1) / up + 2);
  const size_t smem = sizeof(double) * ((size_t)n_taps + (size_t)span);
  if (smem > 200 * 1024) return XCB_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((n_out + kFirTile - 1) / kFirTile), (unsigned)n_signals);
  const PointwiseDev pw = to_dev(pre);
#define XCB_FIR(TI, TO)                                                                                              \
  do {                                                                                                               \
    if (smem > 48 * 1024 &&                                                                                          \
        cudaFuncSetAttribute(upfirdn_kernel<TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=      \
            cudaSuccess)                                                                                             \
      return XCB_ERR_CUDA;                                                                                           \
    upfirdn_kernel<TI, TO><<<grid, 256, smem, st>>>(static_cast<const TI*>(x), x_stride, n_in, static_cast<TO*>(y),  \
                                                    y_stride, n_out, taps, n_taps, up, down, k0, span, pw, stats);   \
  } while (0)
  if (dtype_in == XCB_F32 && dtype_out == XCB_F32) XCB_FIR(float, float);
  else if (dtype_in == XCB_F32) XCB_FIR(float, double);
  else if (dtype_out == XCB_F32) XCB_FIR(double, float);
  else XCB_FIR(double, double);
#undef XCB_FIR
  return cudaGetLastError() == cudaSuccess ? XCB_OK : XCB_ERR_CUDA;
}

}  // extern "C"
