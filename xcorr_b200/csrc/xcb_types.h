// Basic types shared by the device kernels and the host-side thread emulator
// used by the CPU tests (tests/host_emu).  No CUDA headers needed here.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define XCB_HD __host__ __device__ __forceinline__
#else
#define XCB_HD inline
#endif

namespace xcb {

// Complex number with natural vector alignment: cx<float> moves as one 64-bit
// access, cx<double> as one 128-bit access.
template <class T> struct alignas(2 * sizeof(T)) cx {
  T x, y;
};

template <class T> XCB_HD cx<T> cmul(const cx<T>& a, const cx<T>& b) {
  cx<T> r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
// a * conj(b)
template <class T> XCB_HD cx<T> cmulc(const cx<T>& a, const cx<T>& b) {
  cx<T> r;
  r.x = a.x * b.x + a.y * b.y;
  r.y = a.y * b.x - a.x * b.y;
  return r;
}
template <class T> XCB_HD cx<T> cadd(const cx<T>& a, const cx<T>& b) {
  cx<T> r; r.x = a.x + b.x; r.y = a.y + b.y; return r;
}
template <class T> XCB_HD cx<T> csub(const cx<T>& a, const cx<T>& b) {
  cx<T> r; r.x = a.x - b.x; r.y = a.y - b.y; return r;
}
template <class T> XCB_HD cx<T> cconj(const cx<T>& a) {
  cx<T> r; r.x = a.x; r.y = -a.y; return r;
}

}  // namespace xcb
