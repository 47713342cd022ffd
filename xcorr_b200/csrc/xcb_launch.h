// Launch table connecting the planner to the compiled pass kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace xcb {

struct PassLauncher {
  int N;
  // params: const PassParams<T>* (host copy); returns cudaError_t as int
  int (*col)(const void* params, int64_t nblocks, cudaStream_t stream);
  int (*row)(const void* params, int64_t nblocks, cudaStream_t stream);
  int (*prepare)();   // opt in to > 48 KB dynamic shared memory; call once per device
  int col_smem, row_smem, threads;
};

constexpr int kMaxLaunchers = 32;
struct LaunchTable {
  PassLauncher e[kMaxLaunchers];
  int count;
};

LaunchTable& launch_table(int dtype);                    // 0 = f32, 1 = f64
const PassLauncher* find_launcher(int dtype, int N);

}  // namespace xcb

// one per (dtype, group) object file
#define XCB_DECL_REG(dt, g) extern "C" void xcb_register_dt##dt##_g##g(xcb::LaunchTable*);
XCB_DECL_REG(0, 0) XCB_DECL_REG(0, 1) XCB_DECL_REG(0, 2) XCB_DECL_REG(0, 3) XCB_DECL_REG(0, 4) XCB_DECL_REG(0, 5)
XCB_DECL_REG(1, 0) XCB_DECL_REG(1, 1) XCB_DECL_REG(1, 2) XCB_DECL_REG(1, 3) XCB_DECL_REG(1, 4) XCB_DECL_REG(1, 5)
