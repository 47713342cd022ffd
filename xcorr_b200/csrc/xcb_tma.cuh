// Persistent column-pass kernel with TMA-staged tiles (sm_100a).
//
// One CTA stays resident per SM slot and walks tiles blockIdx.x, +gridDim.x, ...
// Each tile ([N rows] x [128 B] strided in HBM) is fetched by ONE elected thread
// with cp.async.bulk.tensor.2d into a ring of shared-memory slots, completion
// signalled on an mbarrier (complete_tx::bytes).  While the CTA computes tile i
// the copies of tiles i+1.. are in flight, so HBM latency is hidden without
// spending registers or LSU issue slots on loads, and out-of-range columns are
// zero-filled by the TMA unit.  The two radix stages run in place in the slot:
//
//   wait(full[s]) -> stage-1 read -> bar -> radix-R1, write -> bar
//   -> stage-2 read -> fence.proxy.async + bar -> re-arm slot s (TMA for tile
//   i+NSLOT) -> radix-R2, twiddle, store to HBM straight from registers.
//
// Arithmetic (twiddle chains, butterflies, lag epilogue) is the same code as the
// plain kernels in xcb_passes.h, which the CPU emulator tests cover.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "xcb_passes.h"

namespace xcb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// same, for waits that can be long (a consumer waiting for its producer): back off between polls so
// the spinning warps do not take issue slots from the warps that have work
#ifndef XCB_PROD_SLEEP_CAP
#define XCB_PROD_SLEEP_CAP 0
#endif
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
#if XCB_PROD_SLEEP_CAP == 0
  while (!mbar_try_wait(bar, parity)) {}
#else
  uint32_t ns = 32;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if (ns < XCB_PROD_SLEEP_CAP) ns *= 2;
  }
#endif
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Warp-wide np.argmax order on (value, index) candidates (amax_better) without branches: the value becomes an
// unsigned key that orders like the floats (NaN on top, -0 = +0, "no candidate" = 0), REDUX.MAX finds the best key
// and REDUX.MIN the lowest index among its holders.  (The shuffle tree with amax_better took a third of the
// arg-max pass of config 5: five rounds of ~25 instructions and six branches each.)
__device__ __forceinline__ void amax_warp_reduce(unsigned live, float& v, uint32_t& o) {
  const float c = v + 0.0f;                                        // -0 -> +0
  const uint32_t u = __float_as_uint(c);
  uint32_t key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  if (c != c) key = 0xffffffffu;
  if (o == kAmaxNone) key = 0u;
  const uint32_t m = __reduce_max_sync(live, key);
  o = __reduce_min_sync(live, key == m ? o : kAmaxNone);
  const uint32_t back = (m & 0x80000000u) ? (m & 0x7fffffffu) : ~m;
  v = m == 0xffffffffu ? __uint_as_float(0x7fc00000u) : __uint_as_float(back);
}
__device__ __forceinline__ void amax_warp_reduce(unsigned live, double& v, uint32_t& o) {
  const double c = v + 0.0;
  const unsigned long long u = (unsigned long long)__double_as_longlong(c);
  unsigned long long key = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
  if (c != c) key = ~0ull;
  if (o == kAmaxNone) key = 0ull;
  const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
  const uint32_t mh = __reduce_max_sync(live, hi);
  const uint32_t ml = __reduce_max_sync(live, hi == mh ? lo : 0u);
  o = __reduce_min_sync(live, (hi == mh && lo == ml) ? o : kAmaxNone);
  const unsigned long long m = ((unsigned long long)mh << 32) | ml;
  const unsigned long long back = (m >> 63) ? (m & 0x7fffffffffffffffull) : ~m;
  v = m == ~0ull ? __longlong_as_double(0x7ff8000000000000ll) : __longlong_as_double((long long)back);
}

#ifndef XCB_HOIST_TW
#define XCB_HOIST_TW 1
#endif
#ifndef XCB_COL_MINB_SMALL
#define XCB_COL_MINB_SMALL 8
#endif
#ifndef XCB_COL_MINB
#define XCB_COL_MINB 3      // resident CTAs per SM the column pass is compiled for, where shared memory allows
#endif
template <class T, int R1, int R2> struct ColTma {
  using G = Cfg<T, R1, R2>;
  static constexpr int N = G::N, C = G::C, TJ = G::TJ, THREADS = G::THREADS;
  static constexpr int SLOT_BYTES = N * C * 2 * (int)sizeof(T);
  static constexpr int BOX_ROWS = N <= 256 ? N : N / 2;
  static constexpr int NBOX = N / BOX_ROWS;
  static constexpr int SMEM_SM = 227 * 1024;                                  // per SM, 1 KB reserved per CTA
  static constexpr int fits(int slots, int ctas) { return (slots * SLOT_BYTES + 64 + THREADS * 8 + 1024) * ctas <= SMEM_SM; }
  // small sub-FFTs have small CTAs (N = 64: 128 threads): compile them for enough CTAs per SM to keep ~24-32
  // warps resident (ncu, config 5: 84 / 103 registers left 5 / 4 CTAs = 16-20 warps per SM)
  static constexpr int WANT = sizeof(T) == 4 ? (THREADS <= 128 ? XCB_COL_MINB_SMALL : THREADS <= 192 ? 4 : 3) : 3;
  static constexpr int MINB = (XCB_COL_MINB >= 3 && fits(3, WANT)) ? WANT : (XCB_COL_MINB >= 3 && fits(2, 3)) ? 3 : 2;
  static constexpr int NSLOT = fits(3, MINB) ? 3 : 2;
  static constexpr int RED_OFF = NSLOT * SLOT_BYTES + 64;
  static constexpr int SMEM = RED_OFF + THREADS * 8;
  static_assert(N % BOX_ROWS == 0 && BOX_ROWS <= 256, "TMA box");
};

// ROLE: 1 forward first pass (real input, norms), 2 forward middle, 3 inverse middle, 4 inverse last (lag
// epilogue).  One kernel per role: the middle passes do not carry the epilogue's code and registers.
enum ColRole : int { COL_ANY = 0, COL_FWD_REAL = 1, COL_FWD = 2, COL_INV = 3, COL_INV_EPI = 4, COL_INV_AMAX = 5 };

template <class T, int R1, int R2, int ROLE>
__global__ void __launch_bounds__(ColTma<T, R1, R2>::THREADS, ColTma<T, R1, R2>::MINB)
xcb_col_pass_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_tail,
                 const int64_t split_col, const __grid_constant__ PassParams<T> p) {
  // (ROLE 0 = COL_ANY reads the modes at run time)
  const int kLoad = ROLE == COL_ANY ? p.load_mode : ROLE == COL_FWD_REAL ? CL_REAL : (ROLE == COL_FWD ? CL_CPLX : CL_CPLX_INV);
  const int kStore = ROLE == COL_ANY ? p.store_mode : ROLE <= COL_FWD ? CS_TW
                     : (ROLE == COL_INV ? CS_CONJ : (ROLE == COL_INV_AMAX ? CS_AMAX : CS_EPI));
  const bool kSumsq = ROLE == COL_ANY || ROLE == COL_FWD_REAL;
  using K = ColTma<T, R1, R2>;
  constexpr int N = K::N, C = K::C;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + K::NSLOT * K::SLOT_BYTES);

  const int tid = threadIdx.x;
  const int c = tid % C, j = tid / C;
  const int64_t ntiles = (int64_t)p.n_batch * p.tiles_per_batch;
  const int64_t stride_tiles = gridDim.x;
  double* red = reinterpret_cast<double*>(smem_raw + K::RED_OFF);

  auto issue = [&](int64_t tile, int s) {
    const int64_t batch = tile / p.tiles_per_batch;
    const int64_t rem = tile - batch * p.tiles_per_batch;
    const int64_t o = rem / p.col_tiles;
    const int64_t cb = rem - o * p.col_tiles;
    const int32_t x0 = (int32_t)(cb * C * 2);                     // in units of T
    const int32_t y0 = (int32_t)(o * N);                          // row within the signal
    unsigned char* dst = smem_raw + s * K::SLOT_BYTES;
    mbar_expect_tx(&full[s], (uint32_t)K::SLOT_BYTES);
    // real input ending inside a row: column tiles left of the end see one more valid row
    const CUtensorMap* map = (cb * C < split_col) ? &tmap : &tmap_tail;
#pragma unroll
    for (int b = 0; b < K::NBOX; ++b)
      tma_load_3d(dst + b * (K::SLOT_BYTES / K::NBOX), map, &full[s], x0, y0 + b * K::BOX_ROWS, (int32_t)batch);
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < K::NSLOT; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < K::NSLOT; ++s) {
      int64_t t0 = (int64_t)blockIdx.x + s * stride_tiles;
      if (t0 < ntiles) issue(t0, s);
    }
  }

  const cx<T> w_inner = ld_ro(p.tw_inner + (j < R1 ? j : 0));   // W_N^j, tile independent
  uint32_t it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += stride_tiles, ++it) {
    const int s = (int)(it % K::NSLOT);
    const uint32_t parity = (it / K::NSLOT) & 1u;
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw + s * K::SLOT_BYTES);

    const int64_t batch = tile / p.tiles_per_batch;
    const int64_t rem = tile - batch * p.tiles_per_batch;
    const int64_t o = rem / p.col_tiles;
    const int64_t col0 = (rem - o * p.col_tiles) * C;
    const int64_t base = o * (int64_t)N * p.inner + col0;
    const bool col_ok = (col0 + c) < p.inner;
    const uint32_t colscale = (uint32_t)(col0 + c) * p.tw_scale;

    cx<T> v[K::TJ];
    // the last inverse pass looks its input twiddles up BEFORE the wait for the tile (two L2-resident table reads
    // per thread whose latency otherwise sits between the first barrier and the first butterfly).  Measured per
    // role: last pass 2.80 -> 2.73 ms, but the run-time-moded N = 150 middle pass 3.68 -> 3.80 ms (four more live
    // registers at its 80-register cap), so only the epilogue roles do it.
    constexpr bool kHoistTw = XCB_HOIST_TW && (ROLE == COL_INV_EPI || ROLE == COL_INV_AMAX);
    cx<T> tw_w, tw_step;
    tw_w.x = tw_step.x = T(1); tw_w.y = tw_step.y = T(0);
    if (kHoistTw && j < R2) {
      tw_w = tw_big(p, colscale * (uint32_t)j);
      tw_step = tw_big(p, colscale * (uint32_t)R2);
    }
    mbar_wait(&full[s], parity);
    if (j < R2) {
#pragma unroll
      for (int r = 0; r < R1; ++r) v[r] = sm[(j + r * R2) * C + c];
    }
    if (kSumsq && p.sumsq_partial) {          // real-input first pass: sum of squares of this tile (rows past n are zero)
      T ssq = T(0);
      if (j < R2) {
#pragma unroll
        for (int r = 0; r < R1; ++r) ssq += v[r].x * v[r].x + v[r].y * v[r].y;
      }
      red[tid] = (double)ssq;
    }
    __syncthreads();
    if (kSumsq && p.sumsq_partial) block_sumsq<K::THREADS>(p.sumsq_partial + tile, red, tid);
    if (j < R2) {
      if (kHoistTw) twiddle_chain_from<R1, true>(tw_w, tw_step, v);
      else if (kLoad == CL_CPLX_INV) twiddle_chain<R1, true>(p, colscale * (uint32_t)j, colscale * (uint32_t)R2, v);
      fft_reg<R1>(v);
#pragma unroll
      for (int r = 0; r < R1; ++r) sm[(j * R1 + r) * C + c] = v[r];
    }
    __syncthreads();
    if (j < R1) {
#pragma unroll
      for (int r = 0; r < R2; ++r) v[r] = sm[(j + r * R1) * C + c];
    }
    fence_proxy_async_smem();   // our generic-proxy accesses to this slot precede its next TMA fill
    __syncthreads();
    if (tid == 0) {
      const int64_t next = tile + K::NSLOT * stride_tiles;
      if (next < ntiles) issue(next, s);
    }
    T amax_v = T(0);
    uint32_t amax_o = kAmaxNone;
    if (j < R1 && col_ok) {
      twiddle_powers<R2>(w_inner, v);
      fft_reg<R2>(v);
      if (ROLE == COL_INV_AMAX) {
        // fused arg-max: nothing is stored.  Branch-free first pass over the thread's 2 R2 lags: largest value
        // and its slot (a compare and two selects per lag); ties and NaNs only raise a flag.  The index / 2-D mask /
        // weight arithmetic then runs ONCE, for the winning slot; the exact np.argmax walk over all lags
        // (amax_consider) is the fallback for flagged threads, a masked winner, or weighted lags.  (First version:
        // exact walk for every lag plus a block reduction per tile -- 0.9 TB/s on the 8 KB tiles of config 5.)
        const uint32_t m0 = (uint32_t)(base + c), inner = (uint32_t)p.inner;
        T best = T(0);
        int slot = -1;
        bool exact = p.unbiased != 0;
        // every row of the tile can hold a kept lag (all lags wanted: the timelapse kernel) and the scale is
        // positive: compare the raw values (same order) and scale the winner only
        const bool all_rows = p.keep_lo[0] <= 0 && p.keep_lo[1] <= p.keep_hi[0] && p.keep_hi[1] >= N && p.escale > T(0);
        if (all_rows) {
          best = v[0].x; slot = 0;
#pragma unroll
          for (int r = 0; r < R2; ++r) {
            const T re = v[r].x, im = -v[r].y;
            if (r > 0) { if (re > best) { best = re; slot = 2 * r; } else if (!(re < best)) exact = true; }
            if (im > best) { best = im; slot = 2 * r + 1; } else if (!(im < best)) exact = true;
          }
          best *= p.escale;
        } else {
#pragma unroll
          for (int r = 0; r < R2; ++r) {
            const int k = j + r * R1;
            if ((k >= p.keep_lo[0] && k < p.keep_hi[0]) || (k >= p.keep_lo[1] && k < p.keep_hi[1])) {
              const T re = v[r].x * p.escale, im = -v[r].y * p.escale;      // the values np.argmax would see
              if (slot < 0 || re > best) { best = re; slot = 2 * r; } else if (!(re < best)) exact = true;
              if (im > best) { best = im; slot = 2 * r + 1; } else if (!(im < best)) exact = true;
            }
          }
        }
        if (slot >= 0 && !exact) {
          const uint32_t l = 2u * (m0 + (uint32_t)(j + (slot >> 1) * R1) * inner) + (uint32_t)(slot & 1);
          if (!amax_lag(p, l, best, amax_o, amax_v, true)) { exact = true; amax_o = kAmaxNone; }
        }
        if (exact) {
#pragma unroll
          for (int r = 0; r < R2; ++r) {
            const int k = j + r * R1;
            if ((k >= p.keep_lo[0] && k < p.keep_hi[0]) || (k >= p.keep_lo[1] && k < p.keep_hi[1]))
              amax_consider(p, m0 + (uint32_t)k * inner, v[r].x, -v[r].y, amax_v, amax_o);
          }
        }
      } else if (kStore == CS_TW) {
        cx<T>* dst = reinterpret_cast<cx<T>*>(p.out) + batch * p.out_stride + base + c;
        twiddle_chain<R2, false>(p, colscale * (uint32_t)j, colscale * (uint32_t)R1, v);
#pragma unroll
        for (int r = 0; r < R2; ++r) dst[(int64_t)(j + r * R1) * p.inner] = v[r];
      } else if (kStore == CS_CONJ) {
        cx<T>* dst = reinterpret_cast<cx<T>*>(p.out) + batch * p.out_stride + base + c;
#pragma unroll
        for (int r = 0; r < R2; ++r) dst[(int64_t)(j + r * R1) * p.inner] = cconj(v[r]);
      } else {
        const int64_t out_off = p.jobs[batch].out_off;
        T* outp = p.ccf + out_off;
        const uint32_t m0 = (uint32_t)(base + c), inner = (uint32_t)p.inner;
        if (epilogue_fast_ok(p) && (out_off & 1) == 0) {
#pragma unroll
          for (int r = 0; r < R2; ++r) {
            const int k = j + r * R1;
            if (k < p.keep_hi[0]) epilogue_store_fast(p, outp, m0 + (uint32_t)k * inner, v[r].x, -v[r].y);
          }
        } else {
#pragma unroll
          for (int r = 0; r < R2; ++r) {
            const int k = j + r * R1;
            if ((k >= p.keep_lo[0] && k < p.keep_hi[0]) || (k >= p.keep_lo[1] && k < p.keep_hi[1]))
              epilogue_store(p, outp, m0 + (uint32_t)k * inner, v[r].x, -v[r].y);
          }
        }
      }
    }
    if (ROLE == COL_INV_AMAX) {
      // one candidate per warp and tile goes to global memory; xcb_argmax_final_rows reduces the tiles x warps of a
      // job (no block barrier here: the tiles of a CTA belong to different jobs anyway)
      constexpr int NW = (K::THREADS + 31) / 32;
      const int lane = tid & 31;
      const int nlane = K::THREADS - (tid & ~31) < 32 ? K::THREADS - (tid & ~31) : 32;   // the last warp may be partial
      const unsigned live = nlane == 32 ? 0xffffffffu : ((1u << nlane) - 1u);
      amax_warp_reduce(live, amax_v, amax_o);
      if (lane == 0) {
        static_cast<T*>(p.amax_val)[tile * NW + (tid >> 5)] = amax_v;
        p.amax_idx[tile * NW + (tid >> 5)] = amax_o == kAmaxNone ? (int64_t)-1 : (int64_t)amax_o;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Pair product with staged, shared spectra rows (16 x R2 row axes)
// ---------------------------------------------------------------------------
// Every pair (i, j) reads the rows of X_i and X_j; the pairs of a window share
// receivers, so a CTA takes ONE mirrored row pair and a GROUP of up to WARPS jobs
// whose distinct spectra (PairGroup, built by the planner) are fetched once with
// cp.async.bulk into shared memory -- one 1-D bulk copy per row, completion on an
// mbarrier -- and then read by all the group's warps (one job per warp, RowWarp
// arithmetic).  Terms (blocks / stacked windows) alternate between two staging
// buffers; the CTA is persistent and the copies of the next item's first terms are
// in flight while the current item is transformed and stored, so no warp waits on
// an L2 / HBM load and the L2 -> SM traffic drops by the sharing factor.
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#ifndef XCB_PAIR_PREFETCH
#define XCB_PAIR_PREFETCH 0     // row pairs ahead (0 = off: 24 / 48 pairs ahead measured 3.80 -> 4.20 ms)
#endif
#ifndef XCB_NBUF8
#define XCB_NBUF8 3
#endif
#ifndef XCB_ACC_BATCH
#define XCB_ACC_BATCH 4     // elements staged per batch when a further term is added (registers: 4 per element)
#endif
template <class T, int R1, int R2> struct PairGrouped {
  using W = RowWarp<T, R1, R2>;
  static constexpr int N = W::N, C = W::C, WARPS = W::THREADS / 32;      // consumer warps = jobs per group
  static constexpr int PP = W::PP;                                       // mirrored row pairs per warp and item
  static constexpr int THREADS = W::THREADS + 32;                        // + one producer warp
  static constexpr int MAX_ROWS = stage_rows<T>();                       // distinct spectra per group
  static constexpr int NBUF = W::LN == 8 ? XCB_NBUF8 : (sizeof(T) == 4 ? 3 : 2);   // staging ring
  static constexpr int ROW_BYTES = N * 2 * (int)sizeof(T);
  static constexpr int STAGE_BYTES = MAX_ROWS * PP * 2 * ROW_BYTES;      // [row][pair][side]
  // resident CTAs the kernel is compiled for: four for the float 8 x 8 kernel (128 consumer threads, 64 registers);
  // two otherwise.  (`W::THREADS <= 128` also matched the DOUBLE 16 x R2 kernels -- 8 columns of 16 lanes -- and capped
  // them at 96 registers: 132 spill instructions, config 3 in float64 30.3 -> 36.6 ms until this was noticed.)
  static constexpr int MINB = (W::LN == 8 && sizeof(T) == 4) ? 4 : 2;
  static constexpr int EXCH_OFF = NBUF * STAGE_BYTES;
  static constexpr int BAR_OFF = EXCH_OFF + C * W::COLBUF * 2 * (int)sizeof(T);
  static constexpr int MAIL_OFF = BAR_OFF + 64;
  // per-CTA tables (filled once): the half-step twiddles W_2N^e of the row axis and, when the launch has few
  // enough groups, one record per (group, consumer warp) -- so that an item starts from shared memory
  // instead of a chain of dependent global loads (group -> job -> norms)
  static constexpr int TW2_OFF = MAIL_OFF + NBUF * 256;
  static constexpr int REC_OFF = TW2_OFF + N * 2 * (int)sizeof(T);
  static constexpr int SMEM_MAX = (227 * 1024 - MINB * 1024) / MINB;                              // MINB CTAs per SM
  static constexpr int MAX_REC_GROUPS = (SMEM_MAX - REC_OFF) / (WARPS * 16) > 0 ? (SMEM_MAX - REC_OFF) / (WARPS * 16) : 0;
  static constexpr int SMEM = REC_OFF + MAX_REC_GROUPS * WARPS * 16;
  static constexpr bool kOk = (ROW_BYTES % 16) == 0 && REC_OFF <= SMEM_MAX && 2 * PP * MAX_ROWS <= 32 && PP <= 2;
};

// the record of consumer warp `w` of group `g` (PairRec, xcb_passes.h)
template <class T> __device__ __forceinline__ PairRec make_pair_rec(const PassParams<T>& p, const PairGroup& g, int w) {
  PairRec r;
  r.batch = -1; r.row_i = r.row_j = r.pad0 = r.pad1 = 0; r.fs = 0.5;
  if (w < g.n_jobs) {
    r.batch = g.job[w];
    r.row_i = g.ji[w];
    r.row_j = (uint8_t)(g.n_i + g.jj[w]);
    const Job jb = p.jobs[r.batch];
    if (p.normalize && p.stack_sig_stride == 0) r.fs = 0.5 * p.rnorm[jb.sig_i] * p.rnorm[jb.sig_j];
  }
  return r;
}
// fills p.recs_global (WARPS records per group) before a launch whose groups do not fit the kernel's shared table
template <class T> __global__ void xcb_fill_pair_recs(const __grid_constant__ PassParams<T> p, int warps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.n_groups * warps) p.recs_global[i] = make_pair_rec(p, p.groups[i / warps], i % warps);
}

// what the producer tells the consumers about the item whose first term sits in a buffer
template <class T> struct alignas(256) PairMail {
  int32_t item;        // -1: no more work
  int32_t ra[2], rb[2];        // per row pair of the item (-1: past the last pair)
  cx<T> wrow[2][2];            // [pair][side]
  alignas(16) PairRec rec[kGroupJobs];   // the item's records, bulk-copied from p.recs_global (launches with many groups)
};
static_assert(sizeof(PairMail<double>) <= 256, "mail slots are 256 bytes apart");

// (Two CTAs per SM: ptxas budgets 96 registers for the float kernel -- the register file is allocated for an even
// number of warps, so the 9 warps count as 10; __maxnreg__(112) leaves ONE CTA per SM: 3.9 -> 5.7 ms.  Forcing the
// staged-row loads of the accumulate phase 2 / 4 / 8 elements ahead of their multiply-adds with volatile shared loads
// (ncu shows ptxas reusing one register pair per element: LDS, FFMA2, FFMA2, LDS ...) spills at 96 registers and
// measured 4.38 / 4.09 / 4.31 ms against 3.91 ms for the compiler's own schedule.)
template <class T, int R1, int R2>
__global__ void __launch_bounds__(PairGrouped<T, R1, R2>::THREADS, PairGrouped<T, R1, R2>::MINB)
xcb_pair_product_grouped(const __grid_constant__ PassParams<T> p) {
  using K = PairGrouped<T, R1, R2>;
  using W = RowWarp<T, R1, R2>;
  using RP = RowPass<T, R1, R2>;
  constexpr int N = K::N, NBUF = K::NBUF, PP = K::PP, LN = W::LN;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + K::BAR_OFF);   // [NBUF] filled by the bulk copies
  uint64_t* empty = full + NBUF;                                         // [NBUF] released by the consumer warps
  cx<T>* exch = reinterpret_cast<cx<T>*>(smem_raw + K::EXCH_OFF);
  PairMail<T>* mail = reinterpret_cast<PairMail<T>*>(smem_raw + K::MAIL_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // item = (block of PP row pairs) * n_groups + group
  const int64_t n_items = ((p.n_pairs_rows + PP - 1) / PP) * (int64_t)p.n_groups;
  const int n_terms = p.n_stack;

  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) { mbar_init(&full[b], 1); mbar_init(&empty[b], K::WARPS); }
    fence_barrier_init();
  }
  cx<T>* tw2 = reinterpret_cast<cx<T>*>(smem_raw + K::TW2_OFF);
  PairRec* recs = reinterpret_cast<PairRec*>(smem_raw + K::REC_OFF);
  const bool use_recs = p.n_groups <= K::MAX_REC_GROUPS;
  for (int i = tid; i < N; i += K::THREADS) tw2[i] = ld_ro(p.tw2_inner + i);
  if (use_recs) {
    for (int i = tid; i < p.n_groups * K::WARPS; i += K::THREADS) {
      recs[i] = make_pair_rec(p, p.groups[i / K::WARPS], i % K::WARPS);
    }
  }
  __syncthreads();

  if (warp == K::WARPS) {
    // ---- producer warp: takes items in order from a global counter (so that the
    // CTAs of the grid stay on neighbouring row pairs and the spectra rows they share
    // come from L2); lane l fetches staged row l/2, side l%2.  The bookkeeping is a
    // chain of dependent global accesses (counter -> group record -> twiddle tables),
    // so it is software pipelined: while the copies of item k are issued, the record
    // of item k+1 is in flight and the counter of item k+2 is being claimed. ----
    struct Raw {                                        // loads issued, values not yet used
      int64_t item, ra, rb, spec;                       // ra, rb: the rows of THIS lane's pair
      int32_t n_i, n_rows, own;
      cx<T> hi, lo;
    };
    const int r = lane / (2 * PP), pp = (lane >> 1) % PP, sd = lane & 1;   // staged row, row pair, side of this lane
    auto claim = [&]() -> unsigned long long { return lane == 0 ? atomicAdd(p.work_counter, 1ULL) : 0ULL; };
    auto start = [&](unsigned long long tok) {
      Raw w;
      w.item = (int64_t)__shfl_sync(0xffffffffu, tok, 0);
      w.ra = w.rb = w.spec = 0; w.n_i = w.n_rows = w.own = 0;
      w.hi.x = w.lo.x = T(1); w.hi.y = w.lo.y = T(0);
      if (w.item < n_items) {
        const int64_t q = w.item / p.n_groups;
        const PairGroup& g = p.groups[w.item - q * p.n_groups];
        if (!RP::pair_rows(p, q * PP + pp, w.ra, w.rb)) w.ra = w.rb = -1;      // past the last row pair
        w.n_i = g.n_i;
        w.n_rows = g.n_rows;
        w.own = g.pad;
        w.spec = g.rows[r < 2 * kGroupJobs ? r : 0];
        if (lane < 2 * PP && w.ra >= 0) {
          const uint32_t kl = RP::klow_of_row(p, sd ? w.rb : w.ra);
          w.hi = ld_ro(p.tw2_hi + (kl >> p.tw_shift));
          w.lo = ld_ro(p.tw2_lo + (kl & p.tw_mask));
        }
      }
      return w;
    };
    uint32_t step = 0;
    Raw cur = start(claim());
    unsigned long long tok = claim();
    for (;;) {
      if (cur.item >= n_items) {
        const int buf = (int)(step % NBUF);
        mbar_wait_backoff(&empty[buf], ((step / NBUF) & 1u) ^ 1u);
        if (lane == 0) { mail[buf].item = -1; mbar_arrive(&full[buf]); }
        break;
      }
      const Raw nxt = start(tok);                       // its loads are in flight during this item's copies
      tok = claim();
      // a self-mirrored row pair stages one side only, a pair past the end nothing
      const bool mine = r < cur.n_rows && r < K::MAX_ROWS && cur.ra >= 0 && (sd == 0 || cur.ra != cur.rb);
      const uint32_t n_copies = __popc(__ballot_sync(0xffffffffu, mine));
      const int64_t src = cur.spec + (sd ? cur.rb : cur.ra) * N;
      const cx<T> wr = cmul(cur.hi, cur.lo);
#if XCB_PAIR_PREFETCH > 0
      // EXPERIMENT (off): the rows this group is the first to hold are requested into L2 for the row pair
      // XCB_PAIR_PREFETCH pairs ahead (the ~300 resident CTAs spread over ~17 row pairs), so that the bulk copies
      // of the staging ring hit L2.  Each row is requested once (owner mask from the planner) with one
      // cp.async.bulk.prefetch.L2 -- and the kernel still got 10 % SLOWER, like the per-line prefetch of round 2.
      if (((cur.own >> r) & 1) && r < K::MAX_ROWS) {
        int64_t pa, pb;
        if (RP::pair_rows(p, (cur.item / p.n_groups + XCB_PAIR_PREFETCH) * PP + pp, pa, pb) && (sd == 0 || pa != pb)) {
          const int64_t psrc = cur.spec + (sd ? pb : pa) * N;
          for (int w = 0; w < n_terms; ++w) {
            const int64_t toff = p.use_blk_off ? (r < cur.n_i ? p.blk_off_i[w] : p.blk_off_j[w]) : w * p.stack_spec_stride;
            bulk_prefetch_l2(p.spec + psrc + toff, K::ROW_BYTES);
          }
        }
      }
#endif
      for (int w = 0; w < n_terms; ++w, ++step) {
        const int buf = (int)(step % NBUF);
        mbar_wait_backoff(&empty[buf], ((step / NBUF) & 1u) ^ 1u);
        if (w == 0 && lane < 2 * PP) {
          mail[buf].wrow[pp][sd] = wr;
          if (sd == 0) { mail[buf].ra[pp] = (int32_t)cur.ra; mail[buf].rb[pp] = (int32_t)cur.rb; }
          if (lane == 0) mail[buf].item = (int32_t)cur.item;
        }
        __syncwarp();
        const bool mail_recs = w == 0 && !use_recs && p.recs_global != nullptr;
        if (lane == 0) {
          mbar_expect_tx(&full[buf], n_copies * (uint32_t)K::ROW_BYTES + (mail_recs ? K::WARPS * 16u : 0u));
          if (mail_recs)
            bulk_load(&mail[buf].rec[0], p.recs_global + (cur.item % p.n_groups) * K::WARPS, K::WARPS * 16u, &full[buf]);
        }
        if (mine) {
          const int64_t toff = p.use_blk_off ? (r < cur.n_i ? p.blk_off_i[w] : p.blk_off_j[w]) : w * p.stack_spec_stride;
          bulk_load(smem_raw + buf * K::STAGE_BYTES + ((r * PP + pp) * 2 + sd) * K::ROW_BYTES, p.spec + src + toff,
                    K::ROW_BYTES, &full[buf]);
        }
      }
      cur = nxt;
    }
    return;
  }

  // ---- consumer warps: one job of the group each ----
  const int side = (tid / LN) & 1, pp = ((tid & 31) / (2 * LN));       // side and row pair of this lane within its warp
  typename W::Ctx ctx;
  ctx.tid = tid;
  uint32_t step = 0;
  const bool per_term = p.normalize && p.stack_sig_stride != 0;
  const cx<T> w_inner = ld_ro(p.tw_inner + (tid % LN));             // W_N^j of this lane, item independent
  for (;;) {
    int buf = (int)(step % NBUF);
    mbar_wait(&full[buf], (step / NBUF) & 1u);         // (no back-off: a sleeping warp wakes up to 0.5 us late)
    const int32_t item = mail[buf].item;                // (the launch guarantees fewer than 2^31 items)
    if (item < 0) break;
    const int32_t ra = mail[buf].ra[pp], rb = mail[buf].rb[pp];
    const int gi = (int)((uint32_t)item % (uint32_t)p.n_groups);
    const bool self = ra == rb;
    // the shuffle path needs every row pair of the warp to be an ordinary one (neither row 0 nor self-mirrored)
    bool ordinary = ra > 0 && ra != rb;
    if (PP == 2) {
      const int32_t oa = mail[buf].ra[pp ^ 1], ob = mail[buf].rb[pp ^ 1];
      ordinary = ordinary && oa > 0 && oa != ob;
    }
    typename W::Slot s;
    bool has_job;
    int row_i, row_j;
    T fs_rec = T(0.5);
    const bool mailed = !use_recs && p.recs_global != nullptr;
    if (use_recs || mailed) {
      const PairRec rec = use_recs ? recs[gi * K::WARPS + warp] : mail[buf].rec[warp];
      has_job = rec.batch >= 0;
      s.batch = has_job ? rec.batch : 0;
      row_i = rec.row_i; row_j = rec.row_j;
      fs_rec = (T)rec.fs;
    } else {
      const PairGroup& g = p.groups[gi];
      has_job = warp < g.n_jobs;
      s.batch = has_job ? g.job[warp] : 0;
      row_i = has_job ? g.ji[warp] : 0; row_j = has_job ? g.n_i + g.jj[warp] : 0;
    }
    s.row = side ? rb : ra;
    s.row0 = ra == 0;
    s.j = tid % LN;
    s.active = has_job && ra >= 0 && !(side && self);
    s.wrow = mail[buf].wrow[pp][side];
    s.w_inner = w_inner;
    s.tw2 = tw2;
    s.fs = ((use_recs || mailed) && !per_term) ? fs_rec : T(-1);
    const int col = tid / LN;
    s.buf = exch + col * W::COLBUF;
    s.mbuf = exch + (self ? col : (col ^ 1)) * W::COLBUF;
    ctx.bid = item;
    const bool lane_on = s.active && s.j < R2;
    Job jb;
    jb.spec_i = jb.spec_j = 0; jb.out_off = 0; jb.sig_i = jb.sig_j = 0;
    if (!(use_recs || mailed) || per_term || s.row0) jb = p.jobs[s.batch];      // (row 0 reads the spectra's slot 0 directly)

    cx<T>* v = ctx.v;
    if (per_term || !lane_on) {
#pragma unroll
      for (int r = 0; r < R1; ++r) v[r].x = v[r].y = T(0);
    }
    int w0 = 0;
    if (!per_term && n_terms == 2 && NBUF >= 3) {
      // both terms in one go (straight-line code: the loads of a whole batch are in flight together)
      const int buf1 = (int)((step + 1) % NBUF);
      mbar_wait(&full[buf1], ((step + 1) / NBUF) & 1u);
      if (lane_on) {
        const cx<T>* st0 = reinterpret_cast<const cx<T>*>(smem_raw + buf * K::STAGE_BYTES) + s.j;
        const cx<T>* st1 = reinterpret_cast<const cx<T>*>(smem_raw + buf1 * K::STAGE_BYTES) + s.j;
        const int oa = ((row_i * PP + pp) * 2 + side) * N, ob = ((row_j * PP + pp) * 2 + side) * N;
        W::template accumulate_two<true, XCB_ACC_BATCH>(v, st0 + oa, st0 + ob, st1 + oa, st1 + ob);
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(&empty[buf]); mbar_arrive(&empty[buf1]); }
      step += 2;
      w0 = 2;
    }
    for (int w = w0; w < n_terms; ++w, ++step) {
      if (w > 0) {
        buf = (int)(step % NBUF);
        mbar_wait(&full[buf], (step / NBUF) & 1u);
      }
      if (lane_on) {
        const cx<T>* st = reinterpret_cast<const cx<T>*>(smem_raw + buf * K::STAGE_BYTES) + s.j;
        const cx<T>* sa = st + ((row_i * PP + pp) * 2 + side) * N;
        const cx<T>* sb = st + ((row_j * PP + pp) * 2 + side) * N;
        if (per_term) {
          const T sw = T(p.rnorm[jb.sig_i + w * p.stack_sig_stride] * p.rnorm[jb.sig_j + w * p.stack_sig_stride]);
          W::template accumulate<true>(v, sa, sb, sw, true);
        } else if (w == 0) {
          W::template accumulate_first<true>(v, sa, sb);
        } else {
          W::template accumulate_next<true, XCB_ACC_BATCH>(v, sa, sb);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[buf]);       // this warp is done with the buffer (and its mail)
    }
    if (sizeof(T) == 4 && has_job && ordinary) {
      // mirrored elements by warp shuffle (warp-uniform branch; lanes j >= R2 hold zeros and only take part)
      W::phase_b_shfl(p, ctx, s);
    } else {
      if (lane_on) {
        if (s.row0 && s.j == 0) v[0] = W::packed_dc(p, jb, per_term);
#pragma unroll
        for (int r = 0; r < R1; ++r) s.buf[s.j + r * R2] = v[r];           // publish C (linear index)
      }
      __syncwarp();
      W::template phase_b<2>(p, ctx, s);
    }
    __syncwarp();
    W::template phase_c<2>(p, ctx, s);
    __syncwarp();
    W::template phase_d<2>(p, ctx, s, smem_raw);
    __syncwarp();                    // the exchange buffers are reused by the next item
  }
}

// Host: 3-D tensor map [batch][rows][row_elems] of T over the pass input, box =
// [1][box_rows][128 bytes]; rows >= `rows` and columns >= row_elems read as
// zero.  Implemented in xcb_api.cu (needs the driver entry point).
int encode_col_tile_map(CUtensorMap* map, int dtype, const void* base, int64_t row_elems, int64_t rows,
                        int64_t batches, int64_t row_pitch_bytes, int64_t batch_pitch_bytes, int box_rows);

}  // namespace xcb
