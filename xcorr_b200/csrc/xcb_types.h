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

// ---------------------------------------------------------------------------
// Complex primitives with an optional packed-f32x2 realisation (PK).
//
// float32 on sm_100a: a complex number IS a packed f32x2 operand (x, y).  FADD2 / FMUL2 / FFMA2 take a
// source with its halves swapped and individually negated for free (SASS `.F32x2.LO_HI.NP`) and broadcast a
// scalar to both halves, so a complex add is ONE instruction, a product by +-i costs nothing and a full
// complex multiply is FMUL2 + FFMA2 -- half the issue slots of the scalar code.  It pays where a kernel is
// bound by instruction issue (the pair product: 4.57 -> 4.2 ms) and costs where it is not (the column
// passes, compiled for 80 registers, spill under the pair-alignment constraints: 3.64 -> 4.10 ms), so the
// kernels opt in per call site with PK = true; PK = false (and double, and the host emulator) is the plain
// scalar form.
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define XCB_PK_DEVICE 1
__device__ __forceinline__ float2 xcb_f2(const cx<float>& a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ cx<float> xcb_c2(float2 v) { cx<float> r; r.x = v.x; r.y = v.y; return r; }
#else
#define XCB_PK_DEVICE 0
#endif
template <bool PK, class T> struct use_pk { static constexpr bool value = PK && XCB_PK_DEVICE && sizeof(T) == 4; };

// -i a, +i a, -a: pure relabelling (on the packed path these fold into operand swizzles)
template <class T> XCB_HD cx<T> cmi(const cx<T>& a) { cx<T> r; r.x = a.y; r.y = -a.x; return r; }
template <class T> XCB_HD cx<T> cpi(const cx<T>& a) { cx<T> r; r.x = -a.y; r.y = a.x; return r; }
template <class T> XCB_HD cx<T> cneg(const cx<T>& a) { cx<T> r; r.x = -a.x; r.y = -a.y; return r; }

template <bool PK = false, class T> XCB_HD cx<T> xadd(const cx<T>& a, const cx<T>& b) {
#if XCB_PK_DEVICE
  if constexpr (use_pk<PK, T>::value) return xcb_c2(__fadd2_rn(xcb_f2(a), xcb_f2(b)));
  else
#endif
  { cx<T> r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
}
template <bool PK = false, class T> XCB_HD cx<T> xsub(const cx<T>& a, const cx<T>& b) {
#if XCB_PK_DEVICE
  if constexpr (use_pk<PK, T>::value) return xcb_c2(__fadd2_rn(xcb_f2(a), make_float2(-b.x, -b.y)));
  else
#endif
  { cx<T> r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
}
// s a  and  s a + b  with a real factor s
template <bool PK = false, class T> XCB_HD cx<T> xscale(const cx<T>& a, T s) {
#if XCB_PK_DEVICE
  if constexpr (use_pk<PK, T>::value) return xcb_c2(__fmul2_rn(xcb_f2(a), make_float2(s, s)));
  else
#endif
  { cx<T> r; r.x = a.x * s; r.y = a.y * s; return r; }
}
template <bool PK = false, class T> XCB_HD cx<T> xfma(T s, const cx<T>& a, const cx<T>& b) {
#if XCB_PK_DEVICE
  if constexpr (use_pk<PK, T>::value) return xcb_c2(__ffma2_rn(xcb_f2(a), make_float2(s, s), xcb_f2(b)));
  else
#endif
  { cx<T> r; r.x = s * a.x + b.x; r.y = s * a.y + b.y; return r; }
}
// a b = a b.x + (i a) b.y ;  a conj(b) = a b.x + (-i a) b.y
template <bool PK = false, class T> XCB_HD cx<T> xmul(const cx<T>& a, const cx<T>& b) {
#if XCB_PK_DEVICE
  if constexpr (use_pk<PK, T>::value) {
    const float2 t = __fmul2_rn(xcb_f2(a), make_float2(b.x, b.x));
    return xcb_c2(__ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), t));
  } else
#endif
  return cmul(a, b);
}
template <bool PK = false, class T> XCB_HD cx<T> xmulc(const cx<T>& a, const cx<T>& b) {
#if XCB_PK_DEVICE
  if constexpr (use_pk<PK, T>::value) {
    const float2 t = __fmul2_rn(xcb_f2(a), make_float2(b.x, b.x));
    return xcb_c2(__ffma2_rn(make_float2(a.y, -a.x), make_float2(b.y, b.y), t));
  } else
#endif
  return cmulc(a, b);
}
// acc + a conj(b)
template <bool PK = false, class T> XCB_HD cx<T> xmulc_acc(const cx<T>& a, const cx<T>& b, const cx<T>& acc) {
#if XCB_PK_DEVICE
  if constexpr (use_pk<PK, T>::value) {
    const float2 t = __ffma2_rn(xcb_f2(a), make_float2(b.x, b.x), xcb_f2(acc));
    return xcb_c2(__ffma2_rn(make_float2(a.y, -a.x), make_float2(b.y, b.y), t));
  } else
#endif
  { cx<T> r; r.x = acc.x + (a.x * b.x + a.y * b.y); r.y = acc.y + (a.y * b.x - a.x * b.y); return r; }
}

}  // namespace xcb
