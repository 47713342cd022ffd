// Microbenchmark: HBM throughput of streaming kernels as a function of their read : write mix on B200.
// (The pair product writes 2.5 bytes for every byte it reads; copy-peak numbers are 1 : 1.)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rw_mix_probe rw_mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
// every thread block streams: reads NR float4 arrays, writes NW float4 arrays of `n` elements each
template <int NR, int NW> __global__ void k(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
    for (int r = 0; r < NR; ++r) { const float4 v = in[r * n + i]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
#pragma unroll
    for (int w = 0; w < NW; ++w) { float4 o = acc; o.x += w; out[w * n + i] = o; }
  }
}
template <int NR, int NW> void run(const float4* in, float4* out, size_t n, const char* name) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * 8;
  k<NR, NW><<<grid, 512>>>(in, out, n);
  cudaEventRecord(a);
  for (int it = 0; it < 5; ++it) k<NR, NW><<<grid, 512>>>(in, out, n);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  const double gb = (double)(NR + NW) * n * 16 / 1e9;
  printf("%-14s read %d : write %d   %.3f ms  total %.0f GB/s  (write %.0f GB/s, read %.0f GB/s)\n", name, NR, NW, ms,
         gb / ms * 1e3, (double)NW * n * 16 / 1e6 / ms, (double)NR * n * 16 / 1e6 / ms);
}
int main() {
  const size_t n = (size_t)1 << 26;            // 1 GiB per array
  float4 *in, *out;
  cudaMalloc(&in, 4 * n * 16); cudaMalloc(&out, 5 * n * 16);
  cudaMemset(in, 0, 4 * n * 16);
  run<1, 1>(in, out, n, "copy");
  run<2, 1>(in, out, n, "read-heavy");
  run<4, 1>(in, out, n, "read-heavy");
  run<1, 2>(in, out, n, "write-heavy");
  run<2, 5>(in, out, n, "pair product");
  run<1, 4>(in, out, n, "write-heavy");
  run<0, 4>(in, out, n, "fill");
  run<4, 0>(in, out, n, "read only");       // (result discarded: the compiler keeps the loads? checked below)
  return 0;
}
