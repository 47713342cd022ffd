// __global__ wrappers around the pass bodies of xcb_passes.h.
// Compiled once per (dtype, group): -DXCB_DT={0,1} -DXCB_GROUP={0..5}; each
// object registers its kernels in the launch table (see Makefile).
#include <cuda_runtime.h>
#include "xcb_launch.h"
#include "xcb_passes.h"
#include "xcb_tma.cuh"
#include <stdlib.h>

#ifndef XCB_DT
#error "XCB_DT must be 0 (float) or 1 (double)"
#endif
#if XCB_DT == 0
typedef float real_t;
#else
typedef double real_t;
#endif

#if XCB_GROUP == 0
#define XCB_GROUP_MENU(X) X(64, 8, 8) X(80, 10, 8) X(96, 12, 8)
#elif XCB_GROUP == 1
#define XCB_GROUP_MENU(X) X(100, 10, 10) X(120, 12, 10) X(128, 16, 8)
#elif XCB_GROUP == 2
#define XCB_GROUP_MENU(X) X(144, 12, 12) X(150, 15, 10) X(160, 16, 10)
#elif XCB_GROUP == 3
#define XCB_GROUP_MENU(X) X(180, 15, 12) X(192, 16, 12) X(200, 20, 10)
#elif XCB_GROUP == 4
#define XCB_GROUP_MENU(X) X(225, 15, 15) X(240, 16, 15) X(256, 16, 16)
#elif XCB_GROUP == 5
#define XCB_GROUP_MENU(X) X(300, 20, 15) X(320, 20, 16) X(400, 20, 20)
#else
#error "XCB_GROUP must be 0..5"
#endif

namespace {

using namespace xcb;

template <class T, int R1, int R2>
__global__ void __launch_bounds__(Cfg<T, R1, R2>::THREADS)
xcb_col_pass(const __grid_constant__ PassParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem[];
  using P = ColPass<T, R1, R2>;
  typename P::Ctx ctx;
  ctx.tid = threadIdx.x;
  ctx.bid = blockIdx.x;
  P::phase_a(p, ctx, smem);
  __syncthreads();
  P::phase_b(p, ctx, smem);
}

// PRO: the pair product (first inverse pass); else the forward last pass / plain complex passes.  Compiled
// separately so that neither role carries the other's registers.
template <class T, int R1, int R2, bool PRO>
__global__ void __launch_bounds__(Cfg<T, R1, R2>::THREADS)
xcb_row_pass(const __grid_constant__ PassParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem[];
  if constexpr (RowWarpOk<R1, R2>::value) {
    // one warp per mirrored row pair, warp-synchronous (RowWarp in xcb_passes.h)
    using P = RowWarp<T, R1, R2>;
    constexpr int MODE = PRO ? 2 : 1;
    typename P::Ctx ctx;
    ctx.tid = threadIdx.x;
    ctx.bid = blockIdx.x;
    const typename P::Slot slot = P::decode(p, ctx, smem);
    P::template phase_a<MODE>(p, ctx, slot, smem);
    __syncwarp();
    P::template phase_b<MODE>(p, ctx, slot);
    __syncwarp();
    P::template phase_c<MODE>(p, ctx, slot);
    __syncwarp();
    if (!PRO && p.sumsq_partial && p.load_mode == RL_REAL) __syncthreads();   // single-pass plans: norms of the tile
    P::template phase_d<MODE>(p, ctx, slot, smem);
  } else {
    using P = RowPass<T, R1, R2>;
    typename P::Ctx ctx;
    ctx.tid = threadIdx.x;
    ctx.bid = blockIdx.x;
    constexpr int MODE = PRO ? 2 : 1;
    P::phase_info(p, ctx, smem);
    if (!PRO) __syncthreads();   // the pair product derives its rows itself
    P::template phase_load<MODE>(p, ctx, smem);
    __syncthreads();
    P::phase_s1_read(p, ctx, smem);
    __syncthreads();
    P::phase_s1_write(p, ctx, smem);
    __syncthreads();
    P::phase_s2(p, ctx, smem);
    __syncthreads();
    P::template phase_store<MODE>(p, ctx, smem);
  }
}

// CTAs of `kernel` that stay resident on the CURRENT device; cached per (device, slot) by the callers
// (slots of one template instantiation; a box may hold different GPUs).
constexpr int kMaxDev = 64;
static int resident_ctas(const void* kernel, int threads, int smem, int* cache) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  int* slot = (dev >= 0 && dev < kMaxDev) ? cache + dev : nullptr;
  if (slot && *slot) return *slot;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const int r = per_sm * (sms > 0 ? sms : 148);
  if (slot) *slot = r;
  return r;
}

// TMA path: the pass input viewed as [batch][rows][128-byte column tiles] with
// 16-byte aligned pitches.  Real input (first forward pass): the zero padding is
// the tensor map's out-of-bounds fill and costs no traffic.  When the samples end
// inside a row (n_real = rows_full * row + rem) two maps are used: column tiles
// left of the end read rows_full + 1 rows, the others rows_full; this needs the
// end to fall on a tile boundary (rem a multiple of 128 bytes).
template <class T> struct TmaGeom {
  bool ok;
  const void* base;
  int64_t row_elems, rows, rows_tail, split_col, row_pitch, batch_pitch;
};

template <class T> TmaGeom<T> tma_geom(const PassParams<T>& p) {
  TmaGeom<T> g;
  g.ok = false;
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("XCB_TMA"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled) return g;
  g.base = p.in;
  if (p.load_mode == CL_REAL) {
    if (p.real_begin < 0) return g;                       // segment starts before the signal: plain kernel
    g.base = static_cast<const T*>(p.in) + p.real_begin;
  }
  if (reinterpret_cast<uintptr_t>(g.base) % 16 != 0) return g;
  g.row_elems = 2 * p.inner;
  g.row_pitch = g.row_elems * (int64_t)sizeof(T);
  if (g.row_pitch % 16 != 0) return g;
  g.split_col = 0;
  if (p.load_mode == CL_REAL) {
    const int64_t tile_reals = 128 / (int64_t)sizeof(T);
    const int64_t rem = p.n_real % g.row_elems;
    if (rem % tile_reals != 0) return g;
    g.rows_tail = p.n_real / g.row_elems;
    g.rows = g.rows_tail + (rem ? 1 : 0);
    g.split_col = rem / 2;
    g.batch_pitch = p.in_stride * (int64_t)sizeof(T);
  } else {
    g.rows = g.rows_tail = p.M / p.inner;
    g.batch_pitch = p.in_stride * (int64_t)sizeof(cx<T>);
  }
  if (g.batch_pitch % 16 != 0 || g.rows_tail < 1) return g;
  g.ok = true;
  return g;
}

template <class T, int R1, int R2> int launch_col(const void* params, int64_t nblocks, cudaStream_t s) {
  using G = Cfg<T, R1, R2>;
  const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
  const TmaGeom<T> tg = tma_geom(p);
  if (tg.ok) {
    using K = ColTma<T, R1, R2>;
    CUtensorMap map, map_tail;
    int rc = encode_col_tile_map(&map, sizeof(T) == 4 ? 0 : 1, tg.base, tg.row_elems, tg.rows, p.n_batch, tg.row_pitch,
                                 tg.batch_pitch, K::BOX_ROWS);
    if (rc == 0) {
      if (tg.rows_tail != tg.rows)
        rc = encode_col_tile_map(&map_tail, sizeof(T) == 4 ? 0 : 1, tg.base, tg.row_elems, tg.rows_tail, p.n_batch,
                                 tg.row_pitch, tg.batch_pitch, K::BOX_ROWS);
      else
        map_tail = map;
    }
    if (rc == 0) {
      // the first forward pass (real input, norms) and the last inverse pass (lag epilogue) have their own
      // kernels; the middle passes share the run-time-moded one (measured: specialising them is slower)
      const int role = p.load_mode == CL_REAL ? COL_FWD_REAL : p.store_mode == CS_EPI ? COL_INV_EPI
                       : p.store_mode == CS_AMAX ? COL_INV_AMAX : COL_ANY;
      static int resident[6][kMaxDev];            // persistent grid: as many CTAs as stay resident (per device)
      const void* fn = role == COL_FWD_REAL ? (const void*)xcb_col_pass_tma<T, R1, R2, COL_FWD_REAL>
                       : role == COL_INV_EPI ? (const void*)xcb_col_pass_tma<T, R1, R2, COL_INV_EPI>
                       : role == COL_INV_AMAX ? (const void*)xcb_col_pass_tma<T, R1, R2, COL_INV_AMAX>
                                             : (const void*)xcb_col_pass_tma<T, R1, R2, COL_ANY>;
      const int res = resident_ctas(fn, K::THREADS, K::SMEM, resident[role]);
      const int64_t grid = nblocks < res ? nblocks : res;
      if (role == COL_FWD_REAL)
        xcb_col_pass_tma<T, R1, R2, COL_FWD_REAL><<<(unsigned)grid, K::THREADS, K::SMEM, s>>>(map, map_tail, tg.split_col, p);
      else if (role == COL_INV_EPI)
        xcb_col_pass_tma<T, R1, R2, COL_INV_EPI><<<(unsigned)grid, K::THREADS, K::SMEM, s>>>(map, map_tail, tg.split_col, p);
      else if (role == COL_INV_AMAX)
        xcb_col_pass_tma<T, R1, R2, COL_INV_AMAX><<<(unsigned)grid, K::THREADS, K::SMEM, s>>>(map, map_tail, tg.split_col, p);
      else
        xcb_col_pass_tma<T, R1, R2, COL_ANY><<<(unsigned)grid, K::THREADS, K::SMEM, s>>>(map, map_tail, tg.split_col, p);
      return (int)cudaGetLastError();
    }
  }
  if (p.store_mode == CS_AMAX) return (int)cudaErrorNotSupported;   // the fused arg-max lives in the TMA kernel only
  xcb_col_pass<T, R1, R2><<<(unsigned)nblocks, G::THREADS, G::COL_SMEM, s>>>(p);
  return (int)cudaGetLastError();
}

template <class T, int R1, int R2> int launch_row(const void* params, int64_t nblocks, cudaStream_t s) {
  using G = Cfg<T, R1, R2>;
  const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
  if constexpr (RowWarpOk<R1, R2>::value) {
    using K = PairGrouped<T, R1, R2>;
    static int staged = -1;
    if (staged < 0) { const char* e = getenv("XCB_STAGED"); staged = (e && e[0] == '0') ? 0 : 1; }
    // worth it when the jobs share spectra (>= 3 jobs per group on average); a lone pair keeps
    // the plain warp kernel, which spreads its row pairs over all warps of a CTA
    if (K::kOk && staged && p.load_mode == RL_PROLOGUE && p.n_groups > 0 && p.work_counter &&
        p.n_batch >= (R1 == 16 ? 3 : 2) * p.n_groups && reinterpret_cast<uintptr_t>(p.spec) % 16 == 0 &&
        p.n_pairs_rows * (int64_t)p.n_groups < 2147483647LL) {
      static int resident_dev[kMaxDev];   // persistent grid (per device)
      const int resident = resident_ctas((const void*)xcb_pair_product_grouped<T, R1, R2>, K::THREADS, K::SMEM, resident_dev);
      const int64_t items = ((p.n_pairs_rows + K::PP - 1) / K::PP) * (int64_t)p.n_groups;
      const int64_t grid = items < resident ? items : resident;
      cudaError_t ce = cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), s);
      if (ce != cudaSuccess) return (int)ce;
      if (p.n_groups > K::MAX_REC_GROUPS && p.recs_global) {
        const int nrec = p.n_groups * K::WARPS;
        xcb_fill_pair_recs<T><<<(unsigned)((nrec + 255) / 256), 256, 0, s>>>(p, K::WARPS);
      }
      xcb_pair_product_grouped<T, R1, R2><<<(unsigned)grid, K::THREADS, K::SMEM, s>>>(p);
      return (int)cudaGetLastError();
    }
  }
  if (p.load_mode == RL_PROLOGUE) xcb_row_pass<T, R1, R2, true><<<(unsigned)nblocks, G::THREADS, G::ROW_SMEM, s>>>(p);
  else xcb_row_pass<T, R1, R2, false><<<(unsigned)nblocks, G::THREADS, G::ROW_SMEM, s>>>(p);
  return (int)cudaGetLastError();
}

template <class T, int R1, int R2> int prepare() {
  using G = Cfg<T, R1, R2>;
  cudaError_t e = cudaFuncSetAttribute(xcb_col_pass<T, R1, R2>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::COL_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(xcb_row_pass<T, R1, R2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::ROW_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(xcb_row_pass<T, R1, R2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::ROW_SMEM);
  if (e != cudaSuccess) return (int)e;
  constexpr int kTmaSmem = ColTma<T, R1, R2>::SMEM;
  e = cudaFuncSetAttribute(xcb_col_pass_tma<T, R1, R2, COL_FWD_REAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(xcb_col_pass_tma<T, R1, R2, COL_ANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(xcb_col_pass_tma<T, R1, R2, COL_INV_EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(xcb_col_pass_tma<T, R1, R2, COL_INV_AMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmem);
  if (e != cudaSuccess) return (int)e;
  if constexpr (RowWarpOk<R1, R2>::value) {
    e = cudaFuncSetAttribute(xcb_pair_product_grouped<T, R1, R2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             PairGrouped<T, R1, R2>::SMEM);
    if (e != cudaSuccess) return (int)e;
  }
  return (int)e;
}

template <class T, int R1, int R2> void add(LaunchTable* t) {
  using G = Cfg<T, R1, R2>;
  PassLauncher l;
  l.N = G::N;
  l.col = &launch_col<T, R1, R2>;
  l.row = &launch_row<T, R1, R2>;
  l.prepare = &prepare<T, R1, R2>;
  l.col_smem = G::COL_SMEM;
  l.row_smem = G::ROW_SMEM;
  l.threads = G::THREADS;
  if (t->count < kMaxLaunchers) t->e[t->count++] = l;
}

}  // namespace

#define XCB_CAT_(a, b, c, d) a##b##c##d
#define XCB_CAT(a, b, c, d) XCB_CAT_(a, b, c, d)

extern "C" void XCB_CAT(xcb_register_dt, XCB_DT, _g, XCB_GROUP)(xcb::LaunchTable* t) {
#define X(n, a, b) add<real_t, a, b>(t);
  XCB_GROUP_MENU(X)
#undef X
}
