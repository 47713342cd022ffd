// C ABI of libxcorr_b200 (see include/xcorr_b200.h): plans, forward
// transforms, pair correlation, argmax, and the host-buffer cc() entry.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/xcorr_b200.h"
#include "xcb_launch.h"
#include "xcb_plan.h"
#include "xcb_tma.cuh"

using namespace xcb;

// ---------------------------------------------------------------------------
// launch table
// ---------------------------------------------------------------------------
namespace xcb {

static LaunchTable g_tables[2];
static std::once_flag g_tables_once;

static void init_tables() {
  g_tables[0].count = g_tables[1].count = 0;
  xcb_register_dt0_g0(&g_tables[0]); xcb_register_dt0_g1(&g_tables[0]); xcb_register_dt0_g2(&g_tables[0]);
  xcb_register_dt0_g3(&g_tables[0]); xcb_register_dt0_g4(&g_tables[0]); xcb_register_dt0_g5(&g_tables[0]);
  xcb_register_dt1_g0(&g_tables[1]); xcb_register_dt1_g1(&g_tables[1]); xcb_register_dt1_g2(&g_tables[1]);
  xcb_register_dt1_g3(&g_tables[1]); xcb_register_dt1_g4(&g_tables[1]); xcb_register_dt1_g5(&g_tables[1]);
}

LaunchTable& launch_table(int dtype) {
  std::call_once(g_tables_once, init_tables);
  return g_tables[dtype ? 1 : 0];
}

const PassLauncher* find_launcher(int dtype, int N) {
  LaunchTable& t = launch_table(dtype);
  for (int i = 0; i < t.count; ++i) if (t.e[i].N == N) return &t.e[i];
  return nullptr;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is
// not linked; the build container has no driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_col_tile_map(CUtensorMap* map, int dtype, const void* base, int64_t row_elems, int64_t rows,
                        int64_t batches, int64_t row_pitch_bytes, int64_t batch_pitch_bytes, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  if (!fn) return -1;
  const cuuint64_t esz = dtype == 0 ? 4 : 8;
  cuuint64_t gdim[3] = {(cuuint64_t)row_elems, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t gstride[2] = {(cuuint64_t)row_pitch_bytes, (cuuint64_t)batch_pitch_bytes};
  cuuint32_t box[3] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                  const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

}  // namespace xcb

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_cuda_err[256] = "";

static int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", where, cudaGetErrorString(e));
  return XCB_ERR_CUDA;
}
#define XCB_CUDA(call)                                         \
  do {                                                         \
    cudaError_t e_ = (call);                                   \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);        \
  } while (0)

// Runs the enclosed launches on the plan's device whatever the caller's current device is
// (a plan created for cuda:1 may be used while cuda:0 is current).
namespace {
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); switched = err == cudaSuccess; }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
}  // namespace

extern "C" const char* xcb_strerror(int status) {
  switch (status) {
    case XCB_OK: return "ok";
    case XCB_ERR_ARG: return "invalid argument";
    case XCB_ERR_SIZE: return "no supported transform length for this n";
    case XCB_ERR_CUDA: return "CUDA error (see xcb_last_cuda_error)";
    case XCB_ERR_ALLOC: return "allocation failed";
    case XCB_ERR_UNSUPPORTED: return "unsupported";
  }
  return "unknown status";
}
extern "C" const char* xcb_last_cuda_error(void) { return g_cuda_err; }
extern "C" int xcb_version(void) { return XCB_VERSION; }

extern "C" int xcb_choose_transform(int64_t n, int32_t dims[4], int32_t* nd) {
  if (n < 1 || !dims || !nd) return XCB_ERR_ARG;
  int d[kMaxDims] = {0, 0, 0, 0}, k = 0;
  if (!plan_choose_dims(n, d, &k)) return XCB_ERR_SIZE;
  for (int i = 0; i < 4; ++i) dims[i] = d[i];
  *nd = k;
  return XCB_OK;
}

// ---------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------
// sumsq[s] = sum of the per-tile partial sums, fixed order -> bit-reproducible
__global__ void __launch_bounds__(256) xcb_sumsq_finish(const double* __restrict__ partial, int64_t tiles,
                                                        int n_blocks, int64_t block_stride,
                                                        double* __restrict__ sumsq) {
  __shared__ double red[256];
  double a = 0.0;
  for (int b = 0; b < n_blocks; ++b) {       // block mode: the blocks partition the signal
    const double* p = partial + b * block_stride + (int64_t)blockIdx.x * tiles;
    for (int64_t i = threadIdx.x; i < tiles; i += 256) a += p[i];
  }
  red[threadIdx.x] = a;
  __syncthreads();
  block_sumsq<256>(sumsq + blockIdx.x, red, threadIdx.x);
}

// rnorm[s] = 1 / ||x_s||  (zero norm -> inf, so the product becomes NaN like util/cc.py:77)
__global__ void xcb_rnorm(const double* __restrict__ sumsq, int64_t n, double* __restrict__ rnorm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rnorm[i] = 1.0 / sqrt(sumsq[i]);
}

// First index of the row maximum (np.argmax): two stages, ties -> lowest index.
template <class T>
__global__ void __launch_bounds__(256) xcb_argmax_partial(const T* __restrict__ rows, int64_t row_len, int chunks,
                                                          T* __restrict__ pval, int64_t* __restrict__ pidx) {
  const int64_t row = blockIdx.y;
  const int chunk = blockIdx.x;
  const int64_t per = (row_len + chunks - 1) / chunks;
  const int64_t lo = (int64_t)chunk * per;
  int64_t hi = lo + per;
  if (hi > row_len) hi = row_len;
  const T* r = rows + row * row_len;
  bool have = false, isnan_ = false;
  T best = T(0);
  int64_t bi = 0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
    T v = r[i];
    if (v != v) { if (!isnan_) { isnan_ = true; best = v; bi = i; have = true; } continue; }  // np.argmax: first NaN wins
    if (!isnan_ && (!have || v > best)) { best = v; bi = i; have = true; }
  }
  __shared__ T sv[256];
  __shared__ int64_t si[256];
  __shared__ int sh[256];   // 0 none, 1 value, 2 nan
  sv[threadIdx.x] = best; si[threadIdx.x] = bi; sh[threadIdx.x] = have ? (isnan_ ? 2 : 1) : 0;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      int a = threadIdx.x, b = threadIdx.x + o;
      bool take = false;
      if (sh[b] > sh[a]) take = true;
      else if (sh[b] == sh[a] && sh[b] == 2) take = si[b] < si[a];
      else if (sh[b] == sh[a] && sh[b] == 1) take = (sv[b] > sv[a]) || (sv[b] == sv[a] && si[b] < si[a]);
      if (take) { sv[a] = sv[b]; si[a] = si[b]; sh[a] = sh[b]; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    pval[row * chunks + chunk] = sv[0];
    pidx[row * chunks + chunk] = sh[0] ? si[0] : (int64_t)-1;
  }
}

// order of np.argmax candidates: a NaN beats any number, then larger value, then lower index
template <class T> __device__ __forceinline__ bool argmax_better(bool hb, T vb, int64_t ib, bool ha, T va, int64_t ia) {
  if (!hb) return false;
  if (!ha) return true;
  const bool nb = vb != vb, na = va != va;
  if (nb != na) return nb;
  if (nb) return ib < ia;
  return vb > va || (vb == va && ib < ia);
}

// one warp per row over the per-chunk partial results
template <class T>
__global__ void __launch_bounds__(32) xcb_argmax_final_rows(const T* __restrict__ pval, const int64_t* __restrict__ pidx,
                                                            int chunks, int64_t n_rows, int64_t* __restrict__ index,
                                                            T* __restrict__ value) {
  const int64_t row = blockIdx.x;
  if (row >= n_rows) return;
  bool have = false;
  T best = T(0);
  int64_t bi = 0;
  for (int c = threadIdx.x; c < chunks; c += 32) {
    const int64_t i = pidx[row * chunks + c];
    const T v = pval[row * chunks + c];
    if (argmax_better<T>(i >= 0, v, i, have, best, bi)) { have = true; best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool ho = __shfl_down_sync(0xffffffffu, (int)have, o) != 0;
    const T vo = __shfl_down_sync(0xffffffffu, best, o);
    const int64_t io = __shfl_down_sync(0xffffffffu, bi, o);
    if (argmax_better<T>(ho, vo, io, have, best, bi)) { have = true; best = vo; bi = io; }
  }
  if (threadIdx.x == 0) {
    index[row] = have ? bi : 0;
    if (value) value[row] = best;
  }
}

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
struct xcb_plan {
  xcb_desc desc;
  std::vector<int32_t> pairs;
  std::vector<Job> jobs;
  PlanGeom g;
  int dtype = 0;
  int device = 0;
  int64_t K = 0, i0 = 0, i1 = 0;
  int64_t n_signals = 0, n_jobs = 0, jobs_in_flight = 0;
  int64_t tiles0 = 0;               // CTAs per signal of the first forward pass
  int64_t win_stride = 1;           // signals between consecutive windows of a receiver
  size_t work_bytes = 0, partial_bytes = 0, rnorm_bytes = 0, recs_bytes = 0;
  void* d_hi = nullptr;
  void* d_lo = nullptr;
  void* d_inner[kMaxDims] = {nullptr, nullptr, nullptr, nullptr};
  void* d_hi2 = nullptr;
  void* d_lo2 = nullptr;
  void* d_inner2 = nullptr;
  Job* d_jobs = nullptr;
  // staged pair product: job groups per launch chunk (chunk c covers jobs [c*jobs_in_flight, ...))
  PairGroup* d_groups = nullptr;
  std::vector<int64_t> group_off;     // first group of chunk c; group_off[n_chunks] = total
  const PassLauncher* launch[kMaxDims] = {nullptr, nullptr, nullptr, nullptr};
  // measurement aid (xcb_profile_*): events around every pass launch
  struct Timed { cudaEvent_t a, b; int kind; };
  mutable bool profiling = false;
  mutable std::vector<Timed> timed;
};

namespace {
struct LaunchTimer {
  const xcb_plan* pl; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr; int kind;
  LaunchTimer(const xcb_plan* p, int k, cudaStream_t s) : pl(p), st(s), kind(k) {
    if (pl->profiling && cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) cudaEventRecord(a, st);
  }
  ~LaunchTimer() {
    if (pl->profiling && a && b) { cudaEventRecord(b, st); pl->timed.push_back({a, b, kind}); }
  }
};
}  // namespace

template <class T> static int upload_table(void** dst, const std::vector<cx<T>>& v) {
  XCB_CUDA(cudaMalloc(dst, v.size() * sizeof(cx<T>)));
  XCB_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(cx<T>), cudaMemcpyHostToDevice));
  return XCB_OK;
}

template <class T> static int upload_tables(xcb_plan* pl) {
  HostTables<T> ht;
  build_tables(pl->g, ht);
  int rc;
  if ((rc = upload_table<T>(&pl->d_hi, ht.hi)) || (rc = upload_table<T>(&pl->d_lo, ht.lo))) return rc;
  for (int a = 0; a < pl->g.nd; ++a)
    if ((rc = upload_table<T>(&pl->d_inner[a], ht.inner[a]))) return rc;
  if ((rc = upload_table<T>(&pl->d_hi2, ht.hi2)) || (rc = upload_table<T>(&pl->d_lo2, ht.lo2)) ||
      (rc = upload_table<T>(&pl->d_inner2, ht.inner2)))
    return rc;
  return XCB_OK;
}

template <class T> static DevTables<T> dev_tables(const xcb_plan* pl) {
  DevTables<T> tb;
  tb.hi = static_cast<const cx<T>*>(pl->d_hi);
  tb.lo = static_cast<const cx<T>*>(pl->d_lo);
  for (int a = 0; a < pl->g.nd; ++a) tb.inner[a] = static_cast<const cx<T>*>(pl->d_inner[a]);
  tb.hi2 = static_cast<const cx<T>*>(pl->d_hi2);
  tb.lo2 = static_cast<const cx<T>*>(pl->d_lo2);
  tb.inner2 = static_cast<const cx<T>*>(pl->d_inner2);
  return tb;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int xcb_plan_destroy(xcb_plan* pl) {
  if (!pl) return XCB_OK;
  cudaFree(pl->d_hi);
  cudaFree(pl->d_lo);
  for (int a = 0; a < kMaxDims; ++a) cudaFree(pl->d_inner[a]);
  cudaFree(pl->d_hi2);
  cudaFree(pl->d_lo2);
  cudaFree(pl->d_inner2);
  cudaFree(pl->d_jobs);
  cudaFree(pl->d_groups);
  delete pl;
  return XCB_OK;
}

extern "C" int xcb_plan_create(xcb_plan** out, const xcb_desc* d) {
  if (!out || !d) return XCB_ERR_ARG;
  *out = nullptr;
  if (d->n < 1 || (d->dtype != XCB_F32 && d->dtype != XCB_F64)) return XCB_ERR_ARG;
  if (d->n_receivers < 1 || d->n_windows < 1 || d->n_pairs < 0 || (d->n_pairs > 0 && !d->pairs)) return XCB_ERR_ARG;
  if (d->nd < 0 || d->nd > kMaxDims) return XCB_ERR_ARG;
  xcb_plan* pl = new (std::nothrow) xcb_plan();
  if (!pl) return XCB_ERR_ALLOC;
  pl->desc = *d;
  pl->dtype = d->dtype;
  pl->pairs.assign(d->pairs, d->pairs + 2 * (size_t)d->n_pairs);
  pl->desc.pairs = nullptr;
  for (int q = 0; q < 2 * d->n_pairs; ++q)
    if (pl->pairs[q] < 0 || pl->pairs[q] >= d->n_receivers) { delete pl; return XCB_ERR_ARG; }

  pl->i0 = d->lag_index_min;
  pl->i1 = d->lag_index_max;
  if (pl->i0 == 0 && pl->i1 == 0) pl->i1 = 2 * d->n - 1;
  if (pl->i0 < 0 || pl->i1 > 2 * d->n - 1 || pl->i0 >= pl->i1) { delete pl; return XCB_ERR_ARG; }
  pl->K = pl->i1 - pl->i0;

  // transform: classic single transform of the window, or -- for cropped
  // non-negative lags -- blocks with a shorter transform (xcb_plan.h, PlanGeom)
  int dims[kMaxDims] = {0, 0, 0, 0}, nd = d->nd, n_blocks = d->n_blocks;
  int64_t block_len = d->block_len;
  const bool blocks_possible = !d->stack;
  if (n_blocks > 1 && (!blocks_possible || nd == 0 || block_len < 1)) { delete pl; return XCB_ERR_ARG; }
  if (n_blocks == 0) {
    n_blocks = 1;
    const char* env = getenv("XCB_BLOCKS");
    if (nd == 0 && blocks_possible && !(env && env[0] == '0')) {
      int bd[kMaxDims], bnd = 0, nb = 0;
      int64_t bl = 0;
      if (plan_choose_blocks(d->n, pl->K, d->n_receivers, d->n_pairs, bd, &bnd, &nb, &bl)) {
        n_blocks = nb; block_len = bl; nd = bnd;
        for (int a = 0; a < bnd; ++a) dims[a] = bd[a];
      }
    }
  }
  if (n_blocks <= 1) {
    if (nd == 0) {
      if (!plan_choose_dims(d->n, dims, &nd)) { delete pl; return XCB_ERR_SIZE; }
      if (d->n_pairs > 0) plan_prefer_row16(d->n, dims, &nd, 1.10);
    } else {
      for (int a = 0; a < nd; ++a) dims[a] = d->dims[a];
    }
    if (!plan_geom_init(pl->g, d->n, dims, nd)) { delete pl; return XCB_ERR_SIZE; }
  } else {
    if (d->nd > 0) for (int a = 0; a < nd; ++a) dims[a] = d->dims[a];
    if (!plan_geom_init(pl->g, d->n, dims, nd, n_blocks, block_len, pl->i0 - (d->n - 1), pl->K)) {
      delete pl;
      return XCB_ERR_SIZE;
    }
  }
  pl->n_signals = (int64_t)d->n_receivers * d->n_windows;
  // stacking sums the cross-spectra over windows inside the pair-product pass: one job per pair
  pl->n_jobs = d->stack ? (int64_t)d->n_pairs : (int64_t)d->n_pairs * d->n_windows;

  { cudaError_t e = cudaGetDevice(&pl->device); if (e != cudaSuccess) { delete pl; return cuda_fail(e, "cudaGetDevice"); } }
  for (int a = 0; a < nd; ++a) {
    pl->launch[a] = find_launcher(pl->dtype, dims[a]);
    if (!pl->launch[a]) { delete pl; return XCB_ERR_SIZE; }
    int e = pl->launch[a]->prepare();
    if (e) { delete pl; return cuda_fail((cudaError_t)e, "cudaFuncSetAttribute"); }
  }
  int rc = pl->dtype == XCB_F32 ? upload_tables<float>(pl) : upload_tables<double>(pl);
  if (rc) { xcb_plan_destroy(pl); return rc; }

  // job table: job = w * n_pairs + p  (window-major, so a stacked launch never
  // holds two jobs that write the same output row)
  const int64_t rs = (d->recv_stride == 0 && d->win_stride == 0) ? d->n_windows : d->recv_stride;
  const int64_t ws = (d->recv_stride == 0 && d->win_stride == 0) ? 1 : d->win_stride;
  pl->win_stride = ws;
  if (pl->n_jobs > 0) {
    std::vector<Job>& jobs = pl->jobs;
    jobs.resize((size_t)pl->n_jobs);
    const int64_t nw = d->stack ? 1 : d->n_windows;
    for (int64_t w = 0; w < nw; ++w)
      for (int64_t q = 0; q < d->n_pairs; ++q) {
        int64_t si = pl->pairs[2 * q] * rs + w * ws, sj = pl->pairs[2 * q + 1] * rs + w * ws;
        jobs[(size_t)(w * d->n_pairs + q)] =
            make_job(pl->g, si, sj, d->stack ? q * pl->K : (q * d->n_windows + w) * pl->K);
      }
    cudaError_t e = cudaMalloc(&pl->d_jobs, jobs.size() * sizeof(Job));
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_jobs, jobs.data(), jobs.size() * sizeof(Job), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { xcb_plan_destroy(pl); return cuda_fail(e, "job table"); }
  }

  // workspace: [sumsq partials][inverse scratch]
  const int C = pl->dtype == XCB_F32 ? tile_cols<float>() : tile_cols<double>();
  pl->tiles0 = pass0_tiles(pl->g, C);
  pl->partial_bytes = align_up((size_t)pl->g.n_blocks * pl->n_signals * pl->tiles0 * sizeof(double), 256);
  pl->rnorm_bytes = align_up((size_t)pl->n_signals * sizeof(double), 256) + 256;   // + the pair product's work counter
  const size_t per_job = (size_t)pl->g.M * (pl->dtype == XCB_F32 ? 8 : 16);
  size_t cap = d->max_work_bytes > 0 ? (size_t)d->max_work_bytes : ((size_t)16 << 30);
  int64_t fl = (int64_t)(cap / per_job);
  if (fl < 1) fl = 1;
  if (fl > pl->n_jobs) fl = pl->n_jobs > 0 ? pl->n_jobs : 1;
  pl->jobs_in_flight = fl;
  pl->work_bytes = align_up((size_t)fl * per_job, 256);

  // job groups of the staged pair product (row axes 16 x R2 only; see launch_row)
  const MenuEntry* rowe = menu_find(dims[nd - 1]);
  if (pl->n_jobs > 0 && (rowe->R1 == 16 || (rowe->R1 == 8 && rowe->R2 == 8))) {
    // one job per consumer warp: R1 lanes per row, C rows per CTA (xcb_tma.cuh, PairGrouped::WARPS)
    const int warps = rowe->R1 * C / 32, max_rows = pl->dtype == XCB_F32 ? stage_rows<float>() : stage_rows<double>();
    std::vector<PairGroup> all, chunk;
    pl->group_off.push_back(0);
    for (int64_t j0 = 0; j0 < pl->n_jobs; j0 += fl) {
      const int nj = (int)std::min<int64_t>(fl, pl->n_jobs - j0);
      build_pair_groups(pl->jobs.data() + j0, nj, warps, max_rows, chunk);
      all.insert(all.end(), chunk.begin(), chunk.end());
      pl->group_off.push_back((int64_t)all.size());
    }
    // per-call records of the largest chunk (xcb_fill_pair_recs), appended to the caller's workspace
    size_t most = 0;
    for (size_t c = 0; c + 1 < pl->group_off.size(); ++c) most = std::max<size_t>(most, (size_t)(pl->group_off[c + 1] - pl->group_off[c]));
    pl->recs_bytes = align_up(most * kGroupJobs * sizeof(PairRec), 256);
    cudaError_t e = cudaMalloc(&pl->d_groups, all.size() * sizeof(PairGroup));
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_groups, all.data(), all.size() * sizeof(PairGroup), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { xcb_plan_destroy(pl); return cuda_fail(e, "group table"); }
  }
  *out = pl;
  return XCB_OK;
}

extern "C" int xcb_plan_get_info(const xcb_plan* pl, xcb_plan_info* info) {
  if (!pl || !info) return XCB_ERR_ARG;
  memset(info, 0, sizeof(*info));
  info->n = pl->g.n;
  info->fft_len = pl->g.L;
  info->spectrum_elems = pl->g.M * spectra_per_signal(pl->g);
  info->n_blocks = pl->g.n_blocks;
  info->block_len = pl->g.block_len;
  info->n_lags = pl->K;
  info->n_signals = pl->n_signals;
  info->n_jobs = pl->n_jobs;
  info->jobs_in_flight = pl->jobs_in_flight;
  info->workspace_bytes = pl->partial_bytes + pl->rnorm_bytes + pl->work_bytes + pl->recs_bytes;
  info->nd = pl->g.nd;
  for (int a = 0; a < kMaxDims; ++a) info->dims[a] = pl->g.dims[a];
  info->dtype = pl->dtype;
  // staged / grouped pair product: job groups of the first launch chunk (0 = the plain row kernel runs)
  info->n_groups = 0;
  if (pl->d_groups && pl->group_off.size() > 1) {
    const int64_t ng = pl->group_off[1] - pl->group_off[0];
    const int64_t nj = std::min<int64_t>(pl->jobs_in_flight, pl->n_jobs);
    const MenuEntry* rowe = menu_find(pl->g.dims[pl->g.nd - 1]);
    if (nj >= (rowe && rowe->R1 == 16 ? 3 : 2) * ng) info->n_groups = (int32_t)ng;       // (launch_row's criterion)
  }
  return XCB_OK;
}

// ---------------------------------------------------------------------------
// forward / correlate
// ---------------------------------------------------------------------------
// runs one pass on the device (the `run` callback of run_forward / run_inverse)
template <class T> struct DeviceRunner {
  const xcb_plan* pl;
  cudaStream_t st;
  int64_t chunk = -1; // inverse passes: launch chunk (selects the job groups)
  unsigned long long* counter = nullptr;   // work-item counter of the staged pair product (in the caller's workspace)
  PairRec* recs = nullptr;                  // per-call group records (caller's workspace, after the inverse scratch)
  int err = 0;        // cudaError_t of a failed launch, or -1 for an oversized grid
  bool operator()(const PassParams<T>& p0, bool row, int kind) {
    const int a = kind % pl->g.nd;
    PassParams<T> p = p0;
    if (row && p.load_mode == RL_PROLOGUE && pl->d_groups && chunk >= 0) {
      p.groups = pl->d_groups + pl->group_off[(size_t)chunk];
      p.n_groups = (int32_t)(pl->group_off[(size_t)chunk + 1] - pl->group_off[(size_t)chunk]);
      p.work_counter = counter;
      p.recs_global = recs;
    }
    const int64_t nblocks = (int64_t)p.n_batch * p.tiles_per_batch;
    if (nblocks >= 2147483647LL) { err = -1; return false; }
    int e;
    {
      LaunchTimer tm(pl, kind, st);
      e = row ? pl->launch[a]->row(&p, nblocks, st) : pl->launch[a]->col(&p, nblocks, st);
    }
    if (e) { err = e; return false; }
    return true;
  }
};

template <class T>
static int forward_t(const xcb_plan* pl, const void* x, int64_t x_stride, void* spec, double* sumsq, void* ws,
                     cudaStream_t st) {
  DevTables<T> tb = dev_tables<T>(pl);
  const PlanGeom& g = pl->g;
  double* partial = static_cast<double*>(ws);
  DeviceRunner<T> run{pl, st};
  if (!run_forward<T>(g, tb, static_cast<const T*>(x), x_stride, (int)pl->n_signals, static_cast<cx<T>*>(spec),
                      partial, run))
    return run.err < 0 ? XCB_ERR_UNSUPPORTED : cuda_fail((cudaError_t)run.err, "forward pass launch");
  xcb_sumsq_finish<<<(unsigned)pl->n_signals, 256, 0, st>>>(partial, pl->tiles0, g.n_blocks,
                                                            pl->n_signals * pl->tiles0, sumsq);
  XCB_CUDA(cudaGetLastError());
  return XCB_OK;
}

extern "C" int xcb_forward(const xcb_plan* pl, const void* x, int64_t x_stride, void* spec, double* sumsq,
                           void* workspace, void* stream) {
  if (!pl || !x || !spec || !sumsq || !workspace || x_stride < pl->g.n) return XCB_ERR_ARG;
  DeviceGuard guard(pl->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return pl->dtype == XCB_F32 ? forward_t<float>(pl, x, x_stride, spec, sumsq, workspace, st)
                              : forward_t<double>(pl, x, x_stride, spec, sumsq, workspace, st);
}

template <class T>
static int correlate_t(const xcb_plan* pl, const void* spec, const double* sumsq, void* ccf, void* ws,
                       cudaStream_t st) {
  DevTables<T> tb = dev_tables<T>(pl);
  const PlanGeom& g = pl->g;
  double* rnorm = reinterpret_cast<double*>(static_cast<unsigned char*>(ws) + pl->partial_bytes);
  cx<T>* work = reinterpret_cast<cx<T>*>(static_cast<unsigned char*>(ws) + pl->partial_bytes + pl->rnorm_bytes);
  xcb_rnorm<<<(unsigned)((pl->n_signals + 255) / 256), 256, 0, st>>>(sumsq, pl->n_signals, rnorm);
  XCB_CUDA(cudaGetLastError());
  CorrFlags f;
  f.i0 = pl->i0;
  f.i1 = pl->i1;
  f.normalize = pl->desc.normalize;
  f.unbiased = pl->desc.unbiased;
  f.accumulate = pl->desc.stack;
  if (pl->desc.stack) {
    f.n_stack = pl->desc.n_windows;
    f.stack_sig_stride = (int)pl->win_stride;
    f.stack_spec_stride = pl->win_stride * g.M;
  }
  DeviceRunner<T> run{pl, st};
  run.counter = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(ws) + pl->partial_bytes +
                                                      pl->rnorm_bytes - 256);
  if (pl->recs_bytes) run.recs = reinterpret_cast<PairRec*>(static_cast<unsigned char*>(ws) + pl->partial_bytes + pl->rnorm_bytes + pl->work_bytes);
  for (int64_t j0 = 0; j0 < pl->n_jobs; j0 += pl->jobs_in_flight) {
    int nj = (int)std::min<int64_t>(pl->jobs_in_flight, pl->n_jobs - j0);
    run.chunk = j0 / pl->jobs_in_flight;
    if (!run_inverse<T>(g, tb, static_cast<const cx<T>*>(spec), rnorm, pl->d_jobs + j0, nj, work,
                        static_cast<T*>(ccf), f, run))
      return run.err < 0 ? XCB_ERR_UNSUPPORTED : cuda_fail((cudaError_t)run.err, "inverse pass launch");
  }
  return XCB_OK;
}

// Fused variant for callers that keep only the maximum of every correlation function
// (scripts/timelapse.py:255-268): the last inverse pass reduces its tile instead of storing it.
template <class T>
static int correlate_argmax_t(const xcb_plan* pl, const void* spec, const double* sumsq, const xcb_lag_mask* mask,
                              int64_t* index, void* value, void* ws, cudaStream_t st) {
  DevTables<T> tb = dev_tables<T>(pl);
  const PlanGeom& g = pl->g;
  if (g.nd < 2 || pl->desc.stack) return XCB_ERR_UNSUPPORTED;
  double* rnorm = reinterpret_cast<double*>(static_cast<unsigned char*>(ws) + pl->partial_bytes);
  cx<T>* work = reinterpret_cast<cx<T>*>(static_cast<unsigned char*>(ws) + pl->partial_bytes + pl->rnorm_bytes);
  xcb_rnorm<<<(unsigned)((pl->n_signals + 255) / 256), 256, 0, st>>>(sumsq, pl->n_signals, rnorm);
  XCB_CUDA(cudaGetLastError());
  const int C = tile_cols<T>();
  // candidates per job of the last inverse pass: tiles x warps of its CTA (xcb_tma.cuh, COL_INV_AMAX)
  const int64_t tiles = (g.M / g.S[0]) * ((g.S[1] + C - 1) / C) * ((pl->launch[0]->threads + 31) / 32);
  const int64_t fl = std::min<int64_t>(pl->jobs_in_flight, pl->n_jobs);
  T* pval = nullptr;
  int64_t* pidx = nullptr;
  XCB_CUDA(cudaMallocAsync(&pval, sizeof(T) * fl * tiles, st));
  XCB_CUDA(cudaMallocAsync(&pidx, sizeof(int64_t) * fl * tiles, st));
  CorrFlags f;
  f.i0 = pl->i0;
  f.i1 = pl->i1;
  f.normalize = pl->desc.normalize;
  f.unbiased = pl->desc.unbiased;
  f.amax_val = pval;
  f.amax_idx = pidx;
  if (mask) { f.mask_w = (int32_t)mask->row_width; f.mask_t = (int32_t)mask->half_width; f.mask_f = (int32_t)mask->n_rows; }
  DeviceRunner<T> run{pl, st};
  run.counter = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(ws) + pl->partial_bytes +
                                                      pl->rnorm_bytes - 256);
  if (pl->recs_bytes) run.recs = reinterpret_cast<PairRec*>(static_cast<unsigned char*>(ws) + pl->partial_bytes + pl->rnorm_bytes + pl->work_bytes);
  int rc = XCB_OK;
  for (int64_t j0 = 0; j0 < pl->n_jobs && rc == XCB_OK; j0 += pl->jobs_in_flight) {
    const int nj = (int)std::min<int64_t>(pl->jobs_in_flight, pl->n_jobs - j0);
    run.chunk = j0 / pl->jobs_in_flight;
    if (!run_inverse<T>(g, tb, static_cast<const cx<T>*>(spec), rnorm, pl->d_jobs + j0, nj, work, (T*)nullptr, f, run)) {
      rc = run.err < 0 || run.err == (int)cudaErrorNotSupported ? XCB_ERR_UNSUPPORTED
                                                                : cuda_fail((cudaError_t)run.err, "inverse pass launch");
      break;
    }
    xcb_argmax_final_rows<T><<<(unsigned)nj, 32, 0, st>>>(pval, pidx, (int)tiles, nj, index + j0,
                                                          value ? static_cast<T*>(value) + j0 : nullptr);
    if (cudaGetLastError() != cudaSuccess) rc = XCB_ERR_CUDA;
  }
  cudaFreeAsync(pval, st);
  cudaFreeAsync(pidx, st);
  return rc;
}

extern "C" int xcb_correlate_argmax(const xcb_plan* pl, const void* spec, const double* sumsq, const xcb_lag_mask* mask,
                                    int64_t* index, void* value, void* workspace, void* stream) {
  if (!pl || !spec || !sumsq || !index || !workspace) return XCB_ERR_ARG;
  if (mask && (mask->row_width < 1 || mask->half_width < 1 || mask->n_rows < 1 ||
               2 * mask->half_width - 1 > mask->row_width || mask->row_width * mask->n_rows > pl->g.n ||
               mask->row_width * (2 * mask->n_rows) >= 2147483647LL))
    return XCB_ERR_ARG;
  DeviceGuard guard(pl->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return pl->dtype == XCB_F32 ? correlate_argmax_t<float>(pl, spec, sumsq, mask, index, value, workspace, st)
                              : correlate_argmax_t<double>(pl, spec, sumsq, mask, index, value, workspace, st);
}

extern "C" int xcb_correlate(const xcb_plan* pl, const void* spec, const double* sumsq, void* ccf, void* workspace,
                             void* stream) {
  if (!pl || !spec || !sumsq || !ccf || !workspace) return XCB_ERR_ARG;
  DeviceGuard guard(pl->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return pl->dtype == XCB_F32 ? correlate_t<float>(pl, spec, sumsq, ccf, workspace, st)
                              : correlate_t<double>(pl, spec, sumsq, ccf, workspace, st);
}

// ---------------------------------------------------------------------------
// per-kernel timing (measurement aid)
// ---------------------------------------------------------------------------
extern "C" int xcb_profile_enable(xcb_plan* pl, int32_t enable) {
  if (!pl) return XCB_ERR_ARG;
  pl->profiling = enable != 0;
  return XCB_OK;
}

extern "C" int xcb_profile_read(xcb_plan* pl, int32_t* n_kinds, double* total_ms, int64_t* launches) {
  if (!pl || !n_kinds || !total_ms || !launches) return XCB_ERR_ARG;
  for (int k = 0; k < 8; ++k) { total_ms[k] = 0.0; launches[k] = 0; }
  *n_kinds = 2 * pl->g.nd;
  cudaError_t err = cudaSuccess;
  for (auto& t : pl->timed) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(t.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, t.a, t.b);
    if (e == cudaSuccess && t.kind >= 0 && t.kind < 8) { total_ms[t.kind] += ms; launches[t.kind] += 1; }
    if (e != cudaSuccess) err = e;
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  pl->timed.clear();
  if (err != cudaSuccess) return cuda_fail(err, "xcb_profile_read");
  return XCB_OK;
}

// ---------------------------------------------------------------------------
// argmax
// ---------------------------------------------------------------------------
template <class T>
static int argmax_t(const void* rows, int64_t n_rows, int64_t row_len, int64_t* index, void* value, cudaStream_t st) {
  int chunks = (int)std::min<int64_t>(1024, (row_len + 16383) / 16384);
  if (chunks < 1) chunks = 1;
  T* pval = nullptr;
  int64_t* pidx = nullptr;
  XCB_CUDA(cudaMallocAsync(&pval, sizeof(T) * n_rows * chunks, st));
  XCB_CUDA(cudaMallocAsync(&pidx, sizeof(int64_t) * n_rows * chunks, st));
  for (int64_t r0 = 0; r0 < n_rows; r0 += 65535) {
    int64_t nr = std::min<int64_t>(65535, n_rows - r0);
    dim3 grid((unsigned)chunks, (unsigned)nr);
    xcb_argmax_partial<T><<<grid, 256, 0, st>>>(static_cast<const T*>(rows) + r0 * row_len, row_len, chunks,
                                               pval + r0 * chunks, pidx + r0 * chunks);
  }
  XCB_CUDA(cudaGetLastError());
  xcb_argmax_final_rows<T><<<(unsigned)n_rows, 32, 0, st>>>(pval, pidx, chunks, n_rows, index, static_cast<T*>(value));
  XCB_CUDA(cudaGetLastError());
  XCB_CUDA(cudaFreeAsync(pval, st));
  XCB_CUDA(cudaFreeAsync(pidx, st));
  return XCB_OK;
}

extern "C" int xcb_argmax(int32_t dtype, const void* rows, int64_t n_rows, int64_t row_len, int64_t* index,
                          void* value, void* stream) {
  if (!rows || !index || n_rows < 1 || row_len < 1) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XCB_F32) return argmax_t<float>(rows, n_rows, row_len, index, value, st);
  if (dtype == XCB_F64) return argmax_t<double>(rows, n_rows, row_len, index, value, st);
  return XCB_ERR_ARG;
}

// ---------------------------------------------------------------------------
// Plain complex transforms on the same pass kernels, and Bluestein's chirp-z
// for arbitrary lengths (obspy Trace.resample, signal/timeshift.py need the
// DFT of exactly npts points).
// ---------------------------------------------------------------------------
struct xcb_fft {
  xcb_plan* pl = nullptr;        // geometry, twiddle tables, launchers (no pairs)
  Job* d_jobs = nullptr;         // output offsets of the inverse epilogue
  int32_t max_batch = 0;
};

__global__ void xcb_fill_jobs(Job* jobs, int32_t n, int64_t out_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { Job j; j.spec_i = j.spec_j = 0; j.sig_i = j.sig_j = 0; j.out_off = (int64_t)i * out_stride; jobs[i] = j; }
}

extern "C" int xcb_fft_good_length(int64_t n_min, int64_t* length) {
  if (n_min < 1 || !length) return XCB_ERR_ARG;
  int d[kMaxDims] = {0, 0, 0, 0}, k = 0;
  if (!plan_choose_dims(n_min, d, &k)) return XCB_ERR_SIZE;
  int64_t m = 1;
  for (int a = 0; a < k; ++a) m *= d[a];
  *length = m;
  return XCB_OK;
}

extern "C" int xcb_fft_destroy(xcb_fft* f) {
  if (!f) return XCB_OK;
  xcb_plan_destroy(f->pl);
  cudaFree(f->d_jobs);
  delete f;
  return XCB_OK;
}

extern "C" int xcb_fft_create(xcb_fft** out, int32_t dtype, int64_t length, int32_t max_batch) {
  if (!out || length < 1 || max_batch < 1) return XCB_ERR_ARG;
  *out = nullptr;
  int d[kMaxDims] = {0, 0, 0, 0}, k = 0;
  if (!plan_choose_dims(length, d, &k)) return XCB_ERR_SIZE;
  int64_t m = 1;
  for (int a = 0; a < k; ++a) m *= d[a];
  if (m != length) return XCB_ERR_SIZE;            // use xcb_fft_good_length
  xcb_desc desc;
  memset(&desc, 0, sizeof(desc));
  desc.n = length;
  desc.dtype = dtype;
  desc.nd = k;
  for (int a = 0; a < k; ++a) desc.dims[a] = d[a];
  desc.n_receivers = 1;
  desc.n_windows = 1;
  desc.n_blocks = 1;
  xcb_fft* f = new (std::nothrow) xcb_fft();
  if (!f) return XCB_ERR_ALLOC;
  int rc = xcb_plan_create(&f->pl, &desc);
  if (rc) { delete f; return rc; }
  f->max_batch = max_batch;
  cudaError_t e = cudaMalloc(&f->d_jobs, sizeof(Job) * (size_t)max_batch);
  if (e != cudaSuccess) { xcb_fft_destroy(f); return cuda_fail(e, "cudaMalloc(fft jobs)"); }
  *out = f;
  return XCB_OK;
}

template <class T>
static int fft_forward_t(const xcb_fft* f, const void* x, int64_t x_stride, int64_t n_valid, int32_t batch, void* spec,
                         cudaStream_t st) {
  const xcb_plan* pl = f->pl;
  const PlanGeom& g = pl->g;
  DevTables<T> tb = dev_tables<T>(pl);
  DeviceRunner<T> run{pl, st};
  for (int a = 0; a < g.nd; ++a) {
    // interleaved complex input == packed real input of 2 n_valid samples
    PassParams<T> p = make_forward_pass<T>(g, tb, a, static_cast<const T*>(x), 2 * x_stride, static_cast<cx<T>*>(spec),
                                           g.M, batch, nullptr, 2 * n_valid, 0);
    if (a == g.nd - 1) p.store_mode = RS_CPLX;      // no real <-> packed conversion
    if (!run(p, a == g.nd - 1, a))
      return run.err < 0 ? XCB_ERR_UNSUPPORTED : cuda_fail((cudaError_t)run.err, "fft forward launch");
  }
  return XCB_OK;
}

extern "C" int xcb_fft_forward(const xcb_fft* f, const void* x, int64_t x_stride, int64_t n_valid, int32_t batch,
                               void* spec, void* stream) {
  if (!f || !x || !spec || batch < 1 || n_valid < 0 || n_valid > f->pl->g.M || x_stride < n_valid) return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return f->pl->dtype == XCB_F32 ? fft_forward_t<float>(f, x, x_stride, n_valid, batch, spec, st)
                                 : fft_forward_t<double>(f, x, x_stride, n_valid, batch, spec, st);
}

template <class T>
static int fft_inverse_t(const xcb_fft* f, void* spec, int32_t batch, void* y, int64_t y_stride, int64_t n_out,
                         cudaStream_t st) {
  const xcb_plan* pl = f->pl;
  const PlanGeom& g = pl->g;
  DevTables<T> tb = dev_tables<T>(pl);
  xcb_fill_jobs<<<(batch + 127) / 128, 128, 0, st>>>(f->d_jobs, batch, 2 * y_stride);
  XCB_CUDA(cudaGetLastError());
  CorrFlags fl;
  fl.i0 = 0;
  fl.i1 = 2 * n_out;
  fl.normalize = 0;
  DeviceRunner<T> run{pl, st};
  for (int a = g.nd - 1; a >= 0; --a) {
    PassParams<T> p = make_inverse_pass<T>(g, tb, a, nullptr, nullptr, f->d_jobs, batch, static_cast<cx<T>*>(spec),
                                           static_cast<T*>(y), fl);
    // output element m = (re, im) of point m: the "direct" lag epilogue with 2 n_out kept reals
    p.epi_direct = 1;
    p.keep_lo[0] = p.keep_lo[1] = p.keep_hi[1] = 0;
    p.keep_hi[0] = g.nd > 1 ? (int32_t)((n_out - 1) / g.S[1] + 1) : 0;
    if (a == g.nd - 1) {               // the spectrum is given (already conjugated), not a pair product
      p.load_mode = RL_CPLX;
      p.tile_major = 0;
    }
    if (!run(p, a == g.nd - 1, g.nd + a))
      return run.err < 0 ? XCB_ERR_UNSUPPORTED : cuda_fail((cudaError_t)run.err, "fft inverse launch");
  }
  return XCB_OK;
}

extern "C" int xcb_fft_inverse_conj(const xcb_fft* f, void* spec_conj, int32_t batch, void* y, int64_t y_stride,
                                    int64_t n_out, void* stream) {
  if (!f || !spec_conj || !y || batch < 1 || batch > f->max_batch || n_out < 1 || n_out > f->pl->g.M ||
      y_stride < n_out)
    return XCB_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return f->pl->dtype == XCB_F32 ? fft_inverse_t<float>(f, spec_conj, batch, y, y_stride, n_out, st)
                                 : fft_inverse_t<double>(f, spec_conj, batch, y, y_stride, n_out, st);
}

// chirp w[k] = exp(sign * i pi k^2 / K): the phase is reduced exactly in integers (k^2 mod 2K)
__device__ __forceinline__ void chirp(int64_t k, int64_t K, int sign, double& c, double& s) {
  const unsigned long long kk = (unsigned long long)(k % (2 * K));       // < 2^32 for every supported K
  const unsigned long long r = (kk * kk) % (unsigned long long)(2 * K);
  sincospi((double)r / (double)K, &s, &c);
  s *= (double)sign;
}

// a[b][k] = x[b][k] * chirp(k), k < n.  MODE 0: x real; 1: x complex; 2: x is the half spectrum
// (K/2 + 1 entries) of a real signal, extended to all K frequencies by conjugate symmetry (imaginary
// parts of the zero and Nyquist bins ignored, like numpy's irfft).
template <class T, int MODE>
__global__ void xcb_chirp_in(const void* __restrict__ x, int64_t x_stride, cx<T>* __restrict__ a, int64_t a_stride,
                             int64_t n, int64_t K, int sign) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t b = blockIdx.y;
  double c, s;
  chirp(k, K, sign, c, s);
  double re, im = 0.0;
  if (MODE == 0) {
    re = static_cast<const T*>(x)[b * x_stride + k];
  } else if (MODE == 1) {
    const cx<T> v = static_cast<const cx<T>*>(x)[b * x_stride + k];
    re = v.x; im = v.y;
  } else {
    const int64_t kh = k <= K / 2 ? k : K - k;
    const cx<T> v = static_cast<const cx<T>*>(x)[b * x_stride + kh];
    re = v.x;
    im = k <= K / 2 ? (double)v.y : -(double)v.y;
    if (k == 0 || 2 * k == K) im = 0.0;
  }
  cx<T> o; o.x = (T)(re * c - im * s); o.y = (T)(re * s + im * c);
  a[b * a_stride + k] = o;
}

// filter b[m] = chirp(-sign)(m) for |m| < K, circular over M points
template <class T> __global__ void xcb_chirp_filter(cx<T>* __restrict__ bf, int64_t M, int64_t K, int sign) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  cx<T> o; o.x = o.y = T(0);
  int64_t d = -1;
  if (m < K) d = m; else if (M - m < K) d = M - m;
  if (d >= 0) { double c, s; chirp(d, K, -sign, c, s); o.x = (T)c; o.y = (T)s; }
  bf[m] = o;
}

// out = conj(a * b)  (both digit-transposed spectra of the same plan; b is shared by the batch)
template <class T>
__global__ void xcb_mul_conj(cx<T>* __restrict__ a, const cx<T>* __restrict__ b, int64_t M) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  cx<T>* row = a + (int64_t)blockIdx.y * M;
  const cx<T> v = cmul(row[m], b[m]);
  row[m] = cconj(v);
}

// y[b][k] = scale * c[b][k] * chirp(k) (optionally conjugated / real part only)
template <class T, bool REAL_OUT>
__global__ void xcb_chirp_out(const cx<T>* __restrict__ c_, int64_t c_stride, void* __restrict__ y, int64_t y_stride,
                              int64_t K, int64_t n_out, int sign, int conj_out, double scale) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_out) return;
  double c, s;
  chirp(k, K, sign, c, s);
  const cx<T> v = c_[(int64_t)blockIdx.y * c_stride + k];
  const double re = ((double)v.x * c - (double)v.y * s) * scale, im = ((double)v.x * s + (double)v.y * c) * scale;
  if (REAL_OUT) static_cast<T*>(y)[(int64_t)blockIdx.y * y_stride + k] = (T)re;
  else { cx<T> o; o.x = (T)re; o.y = (T)(conj_out ? -im : im); static_cast<cx<T>*>(y)[(int64_t)blockIdx.y * y_stride + k] = o; }
}

template <class T>
static int bluestein_t(const xcb_fft* f, int64_t K, int32_t sign, const void* x, int32_t x_complex, int64_t x_stride,
                       int64_t n_in, int32_t batch, void* y, int32_t y_real, int64_t y_stride, int64_t n_out,
                       double scale, void* filter_spec, int32_t filter_ready, void* work, cudaStream_t st) {
  const int64_t M = f->pl->g.M;
  cx<T>* a = static_cast<cx<T>*>(work);                 // [batch][K] chirped input, later [batch][n_out] result
  cx<T>* spec = a + (int64_t)batch * K;                 // [batch][M]
  cx<T>* bf = static_cast<cx<T>*>(filter_spec);         // [M] spectrum of the chirp filter
  const int TB = 256;
  if (!filter_ready) {
    // the filter's time series is built in `spec` (scratch) and transformed into bf
    xcb_chirp_filter<T><<<(unsigned)((M + TB - 1) / TB), TB, 0, st>>>(spec, M, K, sign);
    XCB_CUDA(cudaGetLastError());
    int rc = xcb_fft_forward(f, spec, M, M, 1, bf, st);
    if (rc) return rc;
  }
  // x is zero beyond n_in (n_in <= K); the DFT sign convention: sign = -1 forward, +1 backward
  XCB_CUDA(cudaMemsetAsync(a, 0, sizeof(cx<T>) * (size_t)batch * K, st));
  dim3 gi((unsigned)((n_in + TB - 1) / TB), (unsigned)batch);
  if (x_complex == 2) xcb_chirp_in<T, 2><<<gi, TB, 0, st>>>(x, x_stride, a, K, n_in, K, sign);
  else if (x_complex == 1) xcb_chirp_in<T, 1><<<gi, TB, 0, st>>>(x, x_stride, a, K, n_in, K, sign);
  else xcb_chirp_in<T, 0><<<gi, TB, 0, st>>>(x, x_stride, a, K, n_in, K, sign);
  XCB_CUDA(cudaGetLastError());
  // NB: the chirp of sample k uses length K even when n_in < K (zero-padded DFT of K points)
  int rc = xcb_fft_forward(f, a, K, K, batch, spec, st);
  if (rc) return rc;
  dim3 gm((unsigned)((M + TB - 1) / TB), (unsigned)batch);
  xcb_mul_conj<T><<<gm, TB, 0, st>>>(spec, bf, M);
  XCB_CUDA(cudaGetLastError());
  rc = xcb_fft_inverse_conj(f, spec, batch, a, K, n_out, st);
  if (rc) return rc;
  dim3 go((unsigned)((n_out + TB - 1) / TB), (unsigned)batch);
  if (y_real) xcb_chirp_out<T, true><<<go, TB, 0, st>>>(a, K, y, y_stride, K, n_out, sign, 0, scale);
  else xcb_chirp_out<T, false><<<go, TB, 0, st>>>(a, K, y, y_stride, K, n_out, sign, 0, scale);
  XCB_CUDA(cudaGetLastError());
  return XCB_OK;
}

// DFT of exactly K points (any K) of `batch` signals through Bluestein's algorithm on a plan of
// length M >= 2K - 1:  y[k] = scale * sum_n x[n] exp(sign * 2 pi i n k / K), k < n_out <= K.
//   x     device, real dtype[batch][x_stride] (x_complex 0), complex (1), or the half spectrum of a real signal
//         (2: K/2 + 1 complex entries per signal, n_in = K); n_in <= K samples (zero beyond)
//   y     device, complex[batch][y_stride], or real part only (y_real)
//   filter_spec  device complex[M]: spectrum of the chirp filter; computed when !filter_ready (cache it per (K, sign))
//   work  device, complex[batch * (K + M)]
extern "C" int xcb_dft_bluestein(const xcb_fft* f, int64_t K, int32_t sign, const void* x, int32_t x_complex,
                                 int64_t x_stride, int64_t n_in, int32_t batch, void* y, int32_t y_real,
                                 int64_t y_stride, int64_t n_out, double scale, void* filter_spec,
                                 int32_t filter_ready, void* work, void* stream) {
  if (!f || !x || !y || !filter_spec || !work || K < 1 || batch < 1 || batch > f->max_batch) return XCB_ERR_ARG;
  if (x_complex < 0 || x_complex > 2 || (x_complex == 2 && n_in != K)) return XCB_ERR_ARG;
  if ((sign != 1 && sign != -1) || n_in < 1 || n_in > K || n_out < 1 || n_out > K) return XCB_ERR_ARG;
  if (f->pl->g.M < 2 * K - 1) return XCB_ERR_SIZE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return f->pl->dtype == XCB_F32
             ? bluestein_t<float>(f, K, sign, x, x_complex, x_stride, n_in, batch, y, y_real, y_stride, n_out, scale,
                                  filter_spec, filter_ready, work, st)
             : bluestein_t<double>(f, K, sign, x, x_complex, x_stride, n_in, batch, y, y_real, y_stride, n_out, scale,
                                   filter_spec, filter_ready, work, st);
}

// ---------------------------------------------------------------------------
// cc() with host buffers
// ---------------------------------------------------------------------------
// The reference calls cc() once per (pair, window) with the same n, dtype and flags
// (core/process.py:155-162), so the plan (twiddle tables, job table), the device
// buffers and a private stream are cached per (device, dtype, n, crop, flags): a call
// is two host->device copies, the kernels and one device->host copy.  The host
// buffers are the caller's pageable numpy arrays; the driver stages such copies through
// its own pinned buffers in pipelined chunks, which is what a hand-made pinned staging
// copy would do with an extra pass over the data.
namespace {
struct CcHostEntry {
  int device, dtype, normalize, unbiased;
  int64_t n, i0, i1;
  xcb_plan* pl = nullptr;
  unsigned char* buf = nullptr;
  cudaStream_t st = nullptr;
  size_t x_bytes = 0, spec_bytes = 0, ccf_bytes = 0;
  int64_t stride = 0, n_lags = 0;
  uint64_t last_use = 0;
};
constexpr int kCcHostSlots = 4;
std::mutex g_cc_mu;                  // one cc() at a time per process: the cached buffers are shared
CcHostEntry g_cc_cache[kCcHostSlots];
uint64_t g_cc_clock = 0;

void cc_entry_release(CcHostEntry& e) {
  if (e.pl) {
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != e.device) cudaSetDevice(e.device);
    if (e.st) cudaStreamDestroy(e.st);
    cudaFree(e.buf);
    xcb_plan_destroy(e.pl);
    if (prev != e.device && prev >= 0) cudaSetDevice(prev);
  }
  e = CcHostEntry();
}
}  // namespace

extern "C" int xcb_cc_host_clear_cache(void) {
  std::lock_guard<std::mutex> lock(g_cc_mu);
  for (auto& e : g_cc_cache) cc_entry_release(e);
  return XCB_OK;
}

extern "C" int xcb_cc_host(int32_t dtype, int64_t n, const void* x, const void* y, int32_t normalize,
                           int32_t unbiased, int64_t i0, int64_t i1, void* out) {
  if (!x || !y || !out || n < 1) return XCB_ERR_ARG;
  if (dtype != XCB_F32 && dtype != XCB_F64) return XCB_ERR_ARG;
  const size_t b = dtype == XCB_F32 ? 4 : 8;
  int device = 0;
  XCB_CUDA(cudaGetDevice(&device));
  std::lock_guard<std::mutex> lock(g_cc_mu);
  CcHostEntry* e = nullptr;
  for (auto& c : g_cc_cache)
    if (c.pl && c.device == device && c.dtype == dtype && c.n == n && c.i0 == i0 && c.i1 == i1 &&
        c.normalize == (normalize != 0) && c.unbiased == (unbiased != 0)) { e = &c; break; }
  if (!e) {
    CcHostEntry* victim = &g_cc_cache[0];                 // an empty slot, else the least recently used
    for (auto& c : g_cc_cache) {
      if (!c.pl) { victim = &c; break; }
      if (c.last_use < victim->last_use) victim = &c;
    }
    cc_entry_release(*victim);
    e = victim;
    int32_t pr[2] = {0, 1};
    xcb_desc d;
    memset(&d, 0, sizeof(d));
    d.n = n;
    d.dtype = dtype;
    d.n_receivers = 2;
    d.n_windows = 1;
    d.n_pairs = 1;
    d.pairs = pr;
    d.normalize = normalize;
    d.unbiased = unbiased;
    d.lag_index_min = i0;
    d.lag_index_max = i1;
    int rc = xcb_plan_create(&e->pl, &d);
    if (rc) { *e = CcHostEntry(); return rc; }
    xcb_plan_info info;
    xcb_plan_get_info(e->pl, &info);
    e->device = device; e->dtype = dtype; e->n = n; e->i0 = i0; e->i1 = i1;
    e->normalize = normalize != 0; e->unbiased = unbiased != 0;
    e->stride = (n + 1) / 2 * 2;             // even stride keeps both signals vector-aligned
    e->n_lags = info.n_lags;
    e->x_bytes = align_up((size_t)2 * e->stride * b, 256);
    e->spec_bytes = align_up((size_t)2 * info.spectrum_elems * 2 * b, 256);
    e->ccf_bytes = align_up((size_t)info.n_lags * b, 256);
    cudaError_t ce = cudaMalloc(&e->buf, e->x_bytes + e->spec_bytes + 256 + e->ccf_bytes + info.workspace_bytes);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { cc_entry_release(*e); return cuda_fail(ce, "cudaMalloc(cc_host)"); }
  }
  e->last_use = ++g_cc_clock;
  unsigned char* dx = e->buf;
  unsigned char* dspec = dx + e->x_bytes;
  double* dss = reinterpret_cast<double*>(dspec + e->spec_bytes);
  unsigned char* dccf = dspec + e->spec_bytes + 256;
  unsigned char* dws = dccf + e->ccf_bytes;
  int rc = XCB_OK;
  cudaError_t ce = cudaSuccess;
  do {
    if ((ce = cudaMemcpyAsync(dx, x, n * b, cudaMemcpyHostToDevice, e->st)) != cudaSuccess) break;
    if ((ce = cudaMemcpyAsync(dx + e->stride * b, y, n * b, cudaMemcpyHostToDevice, e->st)) != cudaSuccess) break;
    if ((rc = xcb_forward(e->pl, dx, e->stride, dspec, dss, dws, e->st)) != XCB_OK) break;
    if ((rc = xcb_correlate(e->pl, dspec, dss, dccf, dws, e->st)) != XCB_OK) break;
    if ((ce = cudaMemcpyAsync(out, dccf, (size_t)e->n_lags * b, cudaMemcpyDeviceToHost, e->st)) != cudaSuccess) break;
    ce = cudaStreamSynchronize(e->st);
  } while (0);
  if (rc) return rc;
  if (ce != cudaSuccess) return cuda_fail(ce, "xcb_cc_host");
  return XCB_OK;
}
