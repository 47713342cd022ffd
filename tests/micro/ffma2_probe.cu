// Microbenchmark: issue throughput of scalar FFMA vs packed FFMA2 / FADD2 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(float* out, int iters, float a, float b) {
  float x[16];
  float2 p[8];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = make_float2(x[2 * i], x[2 * i + 1]);
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);          // 16 FFMA = 16 flop-pairs
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(p[i], aa, bb);   // 8 FFMA2 = 16 flop-pairs
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = x[i] + a;                  // 16 FADD
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = __fadd2_rn(p[i], aa);       // 8 FADD2
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += x[i];
  for (int i = 0; i < 8; ++i) s += p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(float* d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(d, 10, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  const int iters = 20000;
  const double lanes = 148.0 * 8 * 256 * 16.0 * iters;   // scalar-op lanes per run
  float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters);
  printf("FFMA  %.2f ms  %.1f Tlane-op/s\nFFMA2 %.2f ms  %.1f Tlane-op/s\nFADD  %.2f ms  %.1f\nFADD2 %.2f ms  %.1f\n",
         t0, lanes / t0 * 1e-9, t1, lanes / t1 * 1e-9, t2, lanes / t2 * 1e-9, t3, lanes / t3 * 1e-9);
  return 0;
}
