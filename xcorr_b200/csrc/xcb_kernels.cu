// __global__ wrappers around the pass bodies of xcb_passes.h.
// Compiled once per (dtype, group): -DXCB_DT={0,1} -DXCB_GROUP={0..5}; each
// object registers its kernels in the launch table (see Makefile).
#include <cuda_runtime.h>
#include "xcb_launch.h"
#include "xcb_passes.h"

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

template <class T, int R1, int R2>
__global__ void __launch_bounds__(Cfg<T, R1, R2>::THREADS)
xcb_row_pass(const __grid_constant__ PassParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem[];
  using P = RowPass<T, R1, R2>;
  typename P::Ctx ctx;
  ctx.tid = threadIdx.x;
  ctx.bid = blockIdx.x;
  P::phase_info(p, ctx, smem);
  __syncthreads();
  P::phase_load(p, ctx, smem);
  __syncthreads();
  P::phase_s1_read(p, ctx, smem);
  __syncthreads();
  P::phase_s1_write(p, ctx, smem);
  __syncthreads();
  P::phase_s2(p, ctx, smem);
  __syncthreads();
  P::phase_store(p, ctx, smem);
}

template <class T, int R1, int R2> int launch_col(const void* params, int64_t nblocks, cudaStream_t s) {
  using G = Cfg<T, R1, R2>;
  const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
  xcb_col_pass<T, R1, R2><<<(unsigned)nblocks, G::THREADS, G::COL_SMEM, s>>>(p);
  return (int)cudaGetLastError();
}

template <class T, int R1, int R2> int launch_row(const void* params, int64_t nblocks, cudaStream_t s) {
  using G = Cfg<T, R1, R2>;
  const PassParams<T>& p = *static_cast<const PassParams<T>*>(params);
  xcb_row_pass<T, R1, R2><<<(unsigned)nblocks, G::THREADS, G::ROW_SMEM, s>>>(p);
  return (int)cudaGetLastError();
}

template <class T, int R1, int R2> int prepare() {
  using G = Cfg<T, R1, R2>;
  cudaError_t e = cudaFuncSetAttribute(xcb_col_pass<T, R1, R2>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::COL_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(xcb_row_pass<T, R1, R2>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::ROW_SMEM);
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
