// Host-side planning for the multi-pass packed-real FFT correlation: choice of
// the transform length and its factorisation into passes, twiddle tables, the
// mirrored-row regions of the pair product, and PassParams construction.
// Pure C++ (no CUDA) so the CPU tests can exercise it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <vector>
#include "xcb_passes.h"

namespace xcb {

struct MenuEntry { int N, R1, R2; };

// Sub-FFT lengths with compiled kernels: N = R1*R2, radices in {8,10,12,15,16,20}.
#define XCB_MENU(X) \
  X(64, 8, 8) X(80, 10, 8) X(96, 12, 8) X(100, 10, 10) X(120, 12, 10) X(128, 16, 8) \
  X(144, 12, 12) X(150, 15, 10) X(160, 16, 10) X(180, 15, 12) X(192, 16, 12) X(200, 20, 10) \
  X(225, 15, 15) X(240, 16, 15) X(256, 16, 16) X(300, 20, 15) X(320, 20, 16) X(400, 20, 20)

static const MenuEntry kMenu[] = {
#define X(n, a, b) {n, a, b},
    XCB_MENU(X)
#undef X
};
static const int kMenuSize = (int)(sizeof(kMenu) / sizeof(kMenu[0]));

inline const MenuEntry* menu_find(int N) {
  for (int i = 0; i < kMenuSize; ++i) if (kMenu[i].N == N) return &kMenu[i];
  return nullptr;
}

struct PlanGeom {
  int64_t n = 0, M = 0, L = 0;
  int nd = 0;
  int dims[kMaxDims] = {0, 0, 0, 0};
  int64_t S[kMaxDims + 1] = {1, 1, 1, 1, 1};   // suffix products, S[nd] = 1
  int tw_shift = 0;
  int n_regions = 0;
  int64_t region_lo[kMaxDims] = {0, 0, 0, 0}, region_hi[kMaxDims] = {0, 0, 0, 0};
  int64_t pair_start[kMaxDims + 1] = {0, 0, 0, 0, 0};
  int64_t n_pairs_rows = 0;
  // Block mode (cropped lags [lag0, lag0 + K), lag0 of either sign): the second signal
  // of a pair is cut into n_blocks blocks of block_len samples, the first into
  // overlapping segments of L samples starting at b*block_len + lag0 (zeros
  // where that falls outside the signal); then
  // r[lag0 + l] = sum_b (segment_b (*) block_b)[l] for l < K as long as
  // L >= block_len + K, so the transforms have length L instead of 2n.
  int n_blocks = 1;
  int64_t block_len = 0, lag0 = 0;
  // lag0 == 0: the last segment holds exactly the samples of the last block (the
  // signal ends inside it), so it is not transformed; its spectrum is the block's
  bool alias_last = false;
};

constexpr int kMaxBlocks = 8;

inline int pow2_exponent(int v) { int e = 0; while (v % 2 == 0) { v /= 2; ++e; } return e; }

// Smallest-cost factorisation M = prod(dims) >= n_min over the menu (<= 4 passes).
// cost ~ HBM traffic: M * (passes + 0.5).  Returns false if nothing fits.
inline bool plan_choose_dims(int64_t n_min, int dims[kMaxDims], int* nd) {
  double best = 1e300;
  int bd[kMaxDims] = {0, 0, 0, 0}, bn = 0;
  int idx[kMaxDims];
  for (int p = 1; p <= kMaxDims; ++p) {
    // non-decreasing index tuples
    for (int i = 0; i < p; ++i) idx[i] = 0;
    while (true) {
      double m = 1;
      int mx = 0;
      bool good_last = false;     // a length <= 256 and multiple of 16 can be the contiguous (row) axis
      bool has256 = false;
      for (int i = 0; i < p; ++i) {
        const int N = kMenu[idx[i]].N;
        m *= N;
        if (N > mx) mx = N;
        if (N <= 256 && N % 16 == 0) good_last = true;
        if (N == 256) has256 = true;
      }
      if (m >= (double)n_min && m < 2147483648.0) {
        // primary: HBM traffic; ties: row axis that fits two CTAs per SM, then balanced lengths.
        // A 256-point row axis (16 x 16, every thread busy in both radix stages, shift/mask
        // indexing) runs the row passes -- pair product, forward last -- ~20 % faster
        // (profiles/tune_r01_dims.log), worth ~6 % more points.
        double cost = m * (p + 0.5 - (has256 && p > 1 ? 0.2 : 0.0)) * (1.0 + (good_last || p == 1 ? 0.0 : 1e-3) + 1e-6 * mx);
        if (cost < best) { best = cost; bn = p; for (int i = 0; i < p; ++i) bd[i] = kMenu[idx[i]].N; }
      }
      int k = p - 1;
      while (k >= 0 && idx[k] == kMenuSize - 1) --k;
      if (k < 0) break;
      ++idx[k];
      for (int i = k + 1; i < p; ++i) idx[i] = idx[k];
    }
  }
  if (bn == 0) return false;
  // Axis order: last = the length <= 256 with the most factors of two (row
  // pass: staging for two CTAs per SM; its length is also the inner stride of
  // the pass before it -> 128-byte aligned column tiles); else the most
  // factors of two; the rest in decreasing order.
  std::sort(bd, bd + bn, [](int a, int b) { return a > b; });
  int last = -1;
  for (int i = 0; i < bn; ++i)
    if (bd[i] <= 256 && (last < 0 || pow2_exponent(bd[i]) > pow2_exponent(bd[last]))) last = i;
  if (last < 0 || pow2_exponent(bd[last]) < 4) {
    int alt = 0;
    for (int i = 1; i < bn; ++i) if (pow2_exponent(bd[i]) > pow2_exponent(bd[alt])) alt = i;
    if (last < 0 || pow2_exponent(bd[alt]) >= 4) last = alt;
  }
  int lv = bd[last];
  for (int i = last; i + 1 < bn; ++i) bd[i] = bd[i + 1];
  bd[bn - 1] = lv;
  for (int i = 0; i < bn; ++i) dims[i] = bd[i];
  *nd = bn;
  return true;
}

// Correlation plans whose chosen row axis is not 16 x R2: the cheapest factorisation WITH such an axis, if it
// costs at most `slack` times the points.  The 16 x R2 row axes run the warp-synchronous row pass and the
// staged / grouped pair product (config 4, 256 stacked windows: 100 x 96 x 96 -> 64 x 64 x 240 has 6.7 % more
// points and takes 10.8 instead of 12.2 ms; larger row axes measured better: 240 > 192 > 160 > 128).
inline bool plan_prefer_row16(int64_t n_min, int dims[kMaxDims], int* nd, double slack) {
  double m0 = 1;
  for (int a = 0; a < *nd; ++a) m0 *= dims[a];
  const MenuEntry* last = menu_find(dims[*nd - 1]);
  if (*nd < 2 || (last && last->R1 == 16)) return false;
  double best = 1e300;
  int bd[kMaxDims] = {0, 0, 0, 0}, bn = 0, idx[kMaxDims];
  for (int p = 2; p <= *nd; ++p) {                      // never more passes than the plan it replaces
    for (int i = 0; i < p; ++i) idx[i] = 0;
    while (true) {
      double m = 1;
      int row = 0;
      for (int i = 0; i < p; ++i) {
        const MenuEntry& e = kMenu[idx[i]];
        m *= e.N;
        if (e.R1 == 16 && e.N <= 256 && e.N > row) row = e.N;
      }
      if (row && m >= (double)n_min && m <= slack * m0) {
        const double cost = m * (p + 0.5) * (1.0 - 1e-4 * row);
        if (cost < best) { best = cost; bn = p; for (int i = 0; i < p; ++i) bd[i] = kMenu[idx[i]].N; }
      }
      int k = p - 1;
      while (k >= 0 && idx[k] == kMenuSize - 1) --k;
      if (k < 0) break;
      ++idx[k];
      for (int i = k + 1; i < p; ++i) idx[i] = idx[k];
    }
  }
  if (bn == 0) return false;
  // row axis last (the largest 16 x R2 length), the others in increasing order (the first pass reads the real
  // input: a short first axis measured faster there)
  int row = -1;
  for (int i = 0; i < bn; ++i) {
    const MenuEntry* e = menu_find(bd[i]);
    if (e->R1 == 16 && bd[i] <= 256 && (row < 0 || bd[i] > bd[row])) row = i;
  }
  const int rv = bd[row];
  for (int i = row; i + 1 < bn; ++i) bd[i] = bd[i + 1];
  std::sort(bd, bd + bn - 1);
  bd[bn - 1] = rv;
  for (int i = 0; i < bn; ++i) dims[i] = bd[i];
  *nd = bn;
  return true;
}

inline bool plan_geom_init(PlanGeom& g, int64_t n, const int* dims, int nd, int n_blocks = 1,
                           int64_t block_len = 0, int64_t lag0 = 0, int64_t n_lags = 0) {
  if (n < 1 || nd < 1 || nd > kMaxDims) return false;
  g.n = n;
  g.n_blocks = n_blocks < 1 ? 1 : n_blocks;
  g.block_len = g.n_blocks > 1 ? block_len : 0;
  g.lag0 = g.n_blocks > 1 ? lag0 : 0;
  g.alias_last = g.n_blocks > 1 && g.lag0 == 0;
  if (g.n_blocks > kMaxBlocks) return false;
  g.nd = nd;
  int64_t M = 1;
  for (int a = 0; a < nd; ++a) {
    if (!menu_find(dims[a])) return false;
    g.dims[a] = dims[a];
    M *= dims[a];
  }
  if (M >= 2147483648LL) return false;
  if (g.n_blocks == 1) {
    if (M < n) return false;
  } else {
    if (block_len < 1 || n_lags < 1 || block_len * g.n_blocks < n) return false;
    if (2 * M < block_len + n_lags) return false;
  }
  g.M = M;
  g.L = 2 * M;
  g.S[nd] = 1;
  for (int a = nd - 1; a >= 0; --a) g.S[a] = g.S[a + 1] * dims[a];
  int sh = 1;
  while ((1LL << (2 * sh)) < M) ++sh;
  g.tw_shift = sh;
  // mirrored-row regions (rows of length N_p): region r <-> first non-zero digit a0 = r+1
  const int64_t Np = dims[nd - 1];
  g.n_regions = nd - 1;
  int64_t next = 1;                       // pair 0 is row 0 paired with itself
  for (int r = 0; r < g.n_regions; ++r) {
    g.region_lo[r] = g.S[r + 1] / Np;
    g.region_hi[r] = g.S[r] / Np;
    g.pair_start[r] = next;
    next += (g.region_hi[r] - g.region_lo[r] + 1) / 2;
  }
  g.pair_start[g.n_regions] = next;
  g.n_pairs_rows = next;
  return true;
}

// ---- twiddle tables (computed in long double, stored in T) ----
template <class T> inline cx<T> unit_root(int64_t num, int64_t den) {   // exp(-2 pi i num/den)
  const long double two_pi = 6.283185307179586476925286766559005768L;
  long double a = two_pi * (long double)(num % den) / (long double)den;
  cx<T> r;
  r.x = (T)cosl(a);
  r.y = (T)(-sinl(a));
  return r;
}

template <class T> struct HostTables {
  std::vector<cx<T>> hi, lo;
  std::vector<cx<T>> inner[kMaxDims];
  std::vector<cx<T>> hi2, lo2, inner2;     // W_L^e (L = 2M) and W_{2 N_p}^t: real <-> packed conversion
};

template <class T> inline void build_tables(const PlanGeom& g, HostTables<T>& t) {
  const int64_t step = 1LL << g.tw_shift;
  t.lo.resize(step);
  for (int64_t j = 0; j < step; ++j) t.lo[j] = unit_root<T>(j, g.M);
  int64_t nh = (g.M + step - 1) / step;
  t.hi.resize(nh);
  for (int64_t h = 0; h < nh; ++h) t.hi[h] = unit_root<T>(h * step, g.M);
  for (int a = 0; a < g.nd; ++a) {
    t.inner[a].resize(g.dims[a]);
    for (int k = 0; k < g.dims[a]; ++k) t.inner[a][k] = unit_root<T>(k, g.dims[a]);
  }
  t.lo2.resize(step);
  for (int64_t j = 0; j < step; ++j) t.lo2[j] = unit_root<T>(j, g.L);
  t.hi2.resize(nh);
  for (int64_t h = 0; h < nh; ++h) t.hi2[h] = unit_root<T>(h * step, g.L);
  const int Np = g.dims[g.nd - 1];
  t.inner2.resize(Np);
  for (int k = 0; k < Np; ++k) t.inner2[k] = unit_root<T>(k, 2 * (int64_t)Np);
}

template <class T> struct DevTables {      // pointers valid where the passes run
  const cx<T>* hi = nullptr;
  const cx<T>* lo = nullptr;
  const cx<T>* inner[kMaxDims] = {nullptr, nullptr, nullptr, nullptr};
  const cx<T>* hi2 = nullptr;
  const cx<T>* lo2 = nullptr;
  const cx<T>* inner2 = nullptr;
};

template <class T> inline int tile_cols() { return 128 / (2 * (int)sizeof(T)); }

template <class T> inline void pass_common(PassParams<T>& p, const PlanGeom& g, const DevTables<T>& tb, int a) {
  p = PassParams<T>();
  const int C = tile_cols<T>();
  p.N = g.dims[a];
  p.M = g.M;
  p.tw_hi = tb.hi;
  p.tw_lo = tb.lo;
  p.tw_inner = tb.inner[a];
  p.tw_shift = g.tw_shift;
  p.tw_mask = (uint32_t)((1LL << g.tw_shift) - 1);
  p.tw2_hi = tb.hi2;
  p.tw2_lo = tb.lo2;
  p.tw2_inner = tb.inner2;
  p.n_regions = g.n_regions;
  for (int r = 0; r < kMaxDims; ++r) { p.region_lo[r] = g.region_lo[r]; p.region_hi[r] = g.region_hi[r]; }
  for (int r = 0; r <= kMaxDims; ++r) p.pair_start[r] = g.pair_start[r];
  p.n_pairs_rows = g.n_pairs_rows;
  p.nd = g.nd;
  for (int i = 0; i < kMaxDims; ++i) p.dims[i] = g.dims[i];
  p.n = g.n;
  p.L = g.L;
  if (a < g.nd - 1) {
    p.inner = g.S[a + 1];
    p.col_tiles = (p.inner + C - 1) / C;
    p.tiles_per_batch = (g.M / g.S[a]) * p.col_tiles;
    p.tw_scale = (uint32_t)(g.M / g.S[a]);
  } else {
    p.inner = 1;
    p.rows = g.M / p.N;
    p.tiles_per_batch = (p.rows + C - 1) / C;
  }
}

// forward pass a: x (real) -> spec for a == 0, spec in place otherwise
template <class T>
inline PassParams<T> make_forward_pass(const PlanGeom& g, const DevTables<T>& tb, int a, const T* x,
                                       int64_t x_stride, cx<T>* spec, int64_t spec_stride, int n_signals,
                                       double* sumsq_partial, int64_t n_real = -1, int64_t real_begin = 0) {
  PassParams<T> p;
  pass_common(p, g, tb, a);
  p.n_batch = n_signals;
  p.out = spec;
  p.out_stride = spec_stride;
  if (a == 0) {
    p.in = x;
    p.in_stride = x_stride;
    p.n_real = n_real >= 0 ? n_real : g.n;
    p.real_begin = real_begin;
    p.real_vec_ok = ((reinterpret_cast<uintptr_t>(x) % (2 * sizeof(T))) == 0 && x_stride % 2 == 0 &&
                     real_begin % 2 == 0) ? 1 : 0;
    p.sumsq_partial = sumsq_partial;
  } else {
    p.in = spec;
    p.in_stride = spec_stride;
  }
  if (a < g.nd - 1) {
    p.load_mode = (a == 0) ? CL_REAL : CL_CPLX;
    p.store_mode = CS_TW;
  } else {
    // last pass: tiles of mirrored row pairs, packed transform -> half spectrum on store
    p.load_mode = (a == 0) ? RL_REAL : RL_CPLX;
    p.store_mode = RS_HALF;
    p.pair_tiles = 1;
    const int C = tile_cols<T>();
    p.tiles_per_batch = (g.n_pairs_rows + (C / 2) - 1) / (C / 2);
  }
  return p;
}

struct CorrFlags {
  int64_t i0 = 0, i1 = 0;
  int normalize = 1, unbiased = 0, accumulate = 0;
  int n_stack = 1, stack_sig_stride = 0;       // spectral stacking over windows
  int64_t stack_spec_stride = 0;
  // fused arg-max instead of the lag store (last inverse pass, column kernels only)
  void* amax_val = nullptr;
  int64_t* amax_idx = nullptr;
  int32_t mask_w = 0, mask_t = 0, mask_f = 0;
  int use_blk_off = 0;                          // block mode: explicit per-block offsets
  int64_t blk_off_i[8] = {0, 0, 0, 0, 0, 0, 0, 0}, blk_off_j[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

// inverse pass a (run a = nd-1 down to 0) over `n_jobs` jobs using work[n_jobs][M]
template <class T>
inline PassParams<T> make_inverse_pass(const PlanGeom& g, const DevTables<T>& tb, int a, const cx<T>* spec,
                                       const double* rnorm, const Job* jobs, int n_jobs, cx<T>* work,
                                       T* ccf, const CorrFlags& f) {
  PassParams<T> p;
  pass_common(p, g, tb, a);
  const int C = tile_cols<T>();
  p.n_batch = n_jobs;
  p.in = work;
  p.out = work;
  p.in_stride = p.out_stride = g.M;
  p.jobs = jobs;
  p.rnorm = rnorm;
  p.spec = spec;
  p.n_stack = f.n_stack;
  p.stack_sig_stride = f.stack_sig_stride;
  p.stack_spec_stride = f.stack_spec_stride;
  p.use_blk_off = f.use_blk_off;
  for (int b = 0; b < 8; ++b) { p.blk_off_i[b] = f.blk_off_i[b]; p.blk_off_j[b] = f.blk_off_j[b]; }
  p.normalize = f.normalize;
  p.ccf = ccf;
  p.i0 = f.i0;
  p.i1 = f.i1;
  p.escale = (T)(1.0 / (double)g.M);
  p.unbiased = f.unbiased;
  p.accumulate = f.accumulate;
  // rows k of the last column pass (element m = k*S[1] + col) that can hold a kept lag
  p.keep_lo[0] = p.keep_lo[1] = 0;
  p.keep_hi[0] = p.keep_hi[1] = 0;
  p.epi_direct = g.n_blocks > 1 ? 1 : 0;
  if (g.n_blocks > 1) {
    if (g.nd > 1) p.keep_hi[0] = (int32_t)((((f.i1 - f.i0) + 1) / 2 - 1) / g.S[1] + 1);
  } else if (g.nd > 1) {
    const int64_t n = g.n, L = g.L, inner0 = g.S[1];
    // lags >= 0: i = l + n-1, l in [0, n)        -> l in [max(0, i0-(n-1)), min(n, i1-(n-1)))
    int64_t la = std::max<int64_t>(0, f.i0 - (n - 1)), lb = std::min<int64_t>(n, f.i1 - (n - 1));
    if (lb > la) { p.keep_lo[0] = (int32_t)((la / 2) / inner0); p.keep_hi[0] = (int32_t)(((lb - 1) / 2) / inner0 + 1); }
    // lags < 0: i = l - L + n-1, i in [0, n-1)   -> l = i + L - (n-1)
    int64_t ia = std::max<int64_t>(0, f.i0), ib = std::min<int64_t>(n - 1, f.i1);
    if (ib > ia) {
      int64_t l0 = ia + L - (n - 1), l1 = ib - 1 + L - (n - 1);
      p.keep_lo[1] = (int32_t)((l0 / 2) / inner0);
      p.keep_hi[1] = (int32_t)((l1 / 2) / inner0 + 1);
    }
  }
  if (a == g.nd - 1) {
    p.tile_major = 1;
    p.pair_tiles = 1;
    p.load_mode = RL_PROLOGUE;
    p.store_mode = (g.nd == 1) ? RS_EPI : RS_CONJ;
    p.tiles_per_batch = (g.n_pairs_rows + (C / 2) - 1) / (C / 2);
  } else {
    p.load_mode = CL_CPLX_INV;
    p.store_mode = (a == 0) ? CS_EPI : CS_CONJ;
    if (a == 0 && f.amax_idx) {
      p.store_mode = CS_AMAX;
      p.amax_val = f.amax_val;
      p.amax_idx = f.amax_idx;
      p.mask_w = f.mask_w; p.mask_t = f.mask_t; p.mask_f = f.mask_f;
      p.mask_shift = -1;
      if (f.mask_w > 0 && (f.mask_w & (f.mask_w - 1)) == 0) { int sh = 0; while ((1 << sh) < f.mask_w) ++sh; p.mask_shift = sh; }
    }
  }
  return p;
}

// ---------------------------------------------------------------------------
// Orchestration shared by the library (kernel launches) and the CPU emulator.
// `run(params, is_row_pass, kind)` executes one pass; kinds 0..nd-1 are forward
// passes, nd..2nd-1 inverse passes of the same axes.  Returns false on failure.
// ---------------------------------------------------------------------------
inline int64_t pass0_tiles(const PlanGeom& g, int C) {
  return g.nd > 1 ? (g.M / g.S[0]) * ((g.S[1] + C - 1) / C) : (g.n_pairs_rows + (C / 2) - 1) / (C / 2);
}

// Spectra per input signal: 1, or in block mode the segments that need their own
// transform (n_blocks, or n_blocks-1 when the last one aliases its block) followed
// by the n_blocks blocks.
inline int segments_per_signal(const PlanGeom& g) { return g.n_blocks - (g.alias_last ? 1 : 0); }
inline int spectra_per_signal(const PlanGeom& g) {
  return g.n_blocks > 1 ? segments_per_signal(g) + g.n_blocks : 1;
}

// Forward transforms of n_signals signals x[s][x_stride].  `partial` receives the
// per-tile sums of squares, [n_blocks][n_signals][pass0_tiles].
template <class T, class Run>
inline bool run_forward(const PlanGeom& g, const DevTables<T>& tb, const T* x, int64_t x_stride, int n_signals,
                        cx<T>* spec, double* partial, Run&& run) {
  const int C = tile_cols<T>();
  if (g.n_blocks == 1) {
    for (int a = 0; a < g.nd; ++a) {
      PassParams<T> p = make_forward_pass<T>(g, tb, a, x, x_stride, spec, g.M, n_signals, a == 0 ? partial : nullptr);
      if (!run(p, a == g.nd - 1, a)) return false;
    }
    return true;
  }
  const int B = g.n_blocks, nseg = segments_per_signal(g), per = spectra_per_signal(g);
  const int64_t t0 = pass0_tiles(g, C);
  for (int role = 0; role < 2; ++role)
    for (int b = 0; b < (role == 0 ? nseg : B); ++b) {
      // segment (role 0) starts lag0 samples after its block and may begin before the signal
      const int64_t start = (int64_t)b * g.block_len + (role == 0 ? g.lag0 : 0);
      int64_t valid = g.n - start;
      const int64_t want = role == 0 ? g.L : std::min<int64_t>(g.block_len, g.n - (int64_t)b * g.block_len);
      if (valid > want) valid = want;
      if (valid < 0) valid = 0;
      PassParams<T> p = make_forward_pass<T>(g, tb, 0, x, x_stride, spec + (int64_t)(role * nseg + b) * g.M,
                                             (int64_t)per * g.M, n_signals,
                                             role == 1 ? partial + (int64_t)b * n_signals * t0 : nullptr, valid,
                                             start);
      if (!run(p, g.nd == 1, 0)) return false;
    }
  for (int a = 1; a < g.nd; ++a) {
    PassParams<T> p = make_forward_pass<T>(g, tb, a, x, x_stride, spec, g.M, n_signals * per, nullptr);
    if (!run(p, a == g.nd - 1, a)) return false;
  }
  return true;
}

template <class T, class Run>
inline bool run_inverse(const PlanGeom& g, const DevTables<T>& tb, const cx<T>* spec, const double* rnorm,
                        const Job* jobs, int n_jobs, cx<T>* work, T* ccf, CorrFlags f, Run&& run) {
  if (g.n_blocks > 1) {                 // the block sum is the stacking loop of the pair product
    const int nseg = segments_per_signal(g);
    f.n_stack = g.n_blocks;
    f.stack_sig_stride = 0;
    f.stack_spec_stride = g.M;
    f.use_blk_off = 1;
    for (int b = 0; b < g.n_blocks; ++b) {
      // Job.spec_i is the signal's first segment, Job.spec_j its first block
      f.blk_off_i[b] = (int64_t)(b < nseg ? b : nseg + b) * g.M;     // aliased segment -> block b of signal i
      f.blk_off_j[b] = (int64_t)b * g.M;
    }
  }
  for (int a = g.nd - 1; a >= 0; --a) {
    PassParams<T> p = make_inverse_pass<T>(g, tb, a, spec, rnorm, jobs, n_jobs, work, ccf, f);
    if (!run(p, a == g.nd - 1, g.nd + a)) return false;
  }
  return true;
}

// Job of pair (signal si, signal sj) writing at out_off
inline Job make_job(const PlanGeom& g, int64_t si, int64_t sj, int64_t out_off) {
  Job j;
  const int64_t per = spectra_per_signal(g);
  j.sig_i = (int32_t)si;
  j.sig_j = (int32_t)sj;
  j.spec_i = si * per * g.M;                                   // segments of signal i
  j.spec_j = (sj * per + (g.n_blocks > 1 ? segments_per_signal(g) : 0)) * g.M;   // blocks of signal j
  j.out_off = out_off;
  return j;
}

// Grouping of the jobs of one launch for the staged pair product (xcb_tma.cuh):
// each group has at most max_jobs jobs and at most max_rows distinct spectra
// (first-role + second-role).  Greedy: two first-role spectra that share
// second-role partners form a 2 x k block (k + 2 rows for 2k jobs), otherwise one
// first-role spectrum with up to max_rows - 1 partners (k + 1 rows for k jobs).
inline void build_pair_groups(const Job* jobs, int n_jobs, int max_jobs, int max_rows, std::vector<PairGroup>& out) {
  out.clear();
  if (max_jobs > kGroupJobs) max_jobs = kGroupJobs;
  std::vector<char> used((size_t)n_jobs, 0);
  std::map<int64_t, std::vector<int>> by_i;              // first-role spectrum -> its jobs, in order
  for (int q = 0; q < n_jobs; ++q) by_i[jobs[q].spec_i].push_back(q);
  auto add = [&](PairGroup& g, int q) {
    const Job& jb = jobs[q];
    int a = 0, b = 0;
    while (a < g.n_i && g.spec_i[a] != jb.spec_i) ++a;
    if (a == g.n_i) g.spec_i[g.n_i++] = jb.spec_i;
    while (b < g.n_j && g.spec_j[b] != jb.spec_j) ++b;
    if (b == g.n_j) g.spec_j[g.n_j++] = jb.spec_j;
    g.job[g.n_jobs] = q;
    g.ji[g.n_jobs] = (uint8_t)a;
    g.jj[g.n_jobs] = (uint8_t)b;
    ++g.n_jobs;
    used[(size_t)q] = 1;
  };
  for (int q0 = 0; q0 < n_jobs;) {
    if (used[(size_t)q0]) { ++q0; continue; }              // (a block may leave q0 for a later group)
    PairGroup g;
    memset(&g, 0, sizeof(g));
    auto it1 = by_i.find(jobs[q0].spec_i);
    std::vector<int> mine;                                // unused jobs of this first-role spectrum
    for (int q : it1->second) if (!used[(size_t)q]) mine.push_back(q);
    // best partner among the following first-role spectra (jobs are window-major: neighbours share a window)
    const int kcols = std::min(max_jobs / 2, max_rows - 2);
    std::vector<std::pair<int, int>> best;                // (my job, partner's job) with the same second role
    if (kcols >= 2 && mine.size() >= 2) {
      auto it2 = it1;
      for (int tries = 0; tries < 16; ++tries) {
        if (++it2 == by_i.end()) break;
        std::vector<std::pair<int, int>> common;
        for (int qa : mine)
          for (int qb : it2->second) {
            if (used[(size_t)qb] || jobs[qb].spec_j != jobs[qa].spec_j) continue;
            bool taken = false;                           // (a pair listed twice must not match the same partner twice)
            for (const auto& c : common) taken = taken || c.second == qb;
            if (!taken) { common.push_back({qa, qb}); break; }
          }
        if (common.size() > best.size()) best = common;
        if ((int)best.size() >= kcols) break;
      }
    }
    if (best.size() >= 2) {
      const int k = std::min<int>((int)best.size(), kcols);
      for (int c = 0; c < k; ++c) { add(g, best[(size_t)c].first); add(g, best[(size_t)c].second); }
    } else {
      const int k = std::min<int>((int)mine.size(), std::min(max_jobs, max_rows - 1));
      for (int c = 0; c < k; ++c) add(g, mine[(size_t)c]);
    }
    out.push_back(g);
  }
  // second pass: merge the leftovers (partial blocks at the edge of the pair triangle, lone pairs) into
  // neighbouring groups while the union still fits -- every group costs a full CTA pass per row pair
  // whatever its size, so fewer, fuller groups are faster
  auto try_merge = [&](PairGroup& a, const PairGroup& b) {
    if (a.n_jobs + b.n_jobs > max_jobs) return false;
    PairGroup u = a;
    for (int w = 0; w < b.n_jobs; ++w) {
      const int64_t si = b.spec_i[b.ji[w]], sj = b.spec_j[b.jj[w]];
      int x = 0, y = 0;
      while (x < u.n_i && u.spec_i[x] != si) ++x;
      if (x == u.n_i) { if (u.n_i == kGroupJobs) return false; u.spec_i[u.n_i++] = si; }
      while (y < u.n_j && u.spec_j[y] != sj) ++y;
      if (y == u.n_j) { if (u.n_j == kGroupJobs) return false; u.spec_j[u.n_j++] = sj; }
      if (u.n_i + u.n_j > max_rows) return false;
      u.job[u.n_jobs] = b.job[w];
      u.ji[u.n_jobs] = (uint8_t)x;
      u.jj[u.n_jobs] = (uint8_t)y;
      ++u.n_jobs;
    }
    a = u;
    return true;
  };
  std::vector<char> dead(out.size(), 0);
  const int kWindow = 48;                                 // groups of one window are neighbours
  for (size_t b = 0; b < out.size(); ++b) {
    if (out[b].n_jobs >= max_jobs) continue;
    // best fit: the fullest neighbour that still takes this group
    int best = -1;
    const size_t lo = b > (size_t)kWindow ? b - kWindow : 0, hi = std::min(out.size(), b + kWindow + 1);
    for (size_t a = lo; a < hi; ++a) {
      if (a == b || dead[a] || out[a].n_jobs >= max_jobs || out[a].n_jobs + out[b].n_jobs > max_jobs) continue;
      PairGroup trial = out[a];
      if (!try_merge(trial, out[b])) continue;
      if (best < 0 || out[a].n_jobs > out[(size_t)best].n_jobs) best = (int)a;
    }
    if (best >= 0) { try_merge(out[(size_t)best], out[b]); dead[b] = 1; }
  }
  size_t k = 0;
  for (size_t a = 0; a < out.size(); ++a) if (!dead[a]) out[k++] = out[a];
  out.resize(k);
  // `pad` = bit mask of the staged rows this group is the FIRST to hold: the producer of that group requests those
  // spectra rows into L2 for a row pair well ahead of the grid's front, so every row is requested exactly once
  std::vector<int64_t> seen;
  for (PairGroup& g : out) {
    for (int a = 0; a < g.n_i; ++a) g.rows[a] = g.spec_i[a];
    for (int b = 0; b < g.n_j; ++b) g.rows[g.n_i + b] = g.spec_j[b];
    g.n_rows = g.n_i + g.n_j;
    g.pad = 0;
    for (int r = 0; r < g.n_rows; ++r) {
      if (std::find(seen.begin(), seen.end(), g.rows[r]) == seen.end()) {
        seen.push_back(g.rows[r]);
        g.pad |= 1 << r;
      }
    }
  }
}

// Block-mode planner: for kept lags [lag0, lag0 + K) (any sign of lag0) choose the
// number of blocks and the transform that minimise the HBM traffic model
//   P * ((2B + 4) F + K) + N * 2B * 5.5 F      (F = transform length in reals)
// against the classic  P * (6 L + K) + N * (n + 5 L).  Returns false when the
// classic single transform is at least as good.
inline bool plan_choose_blocks(int64_t n, int64_t n_lags, int n_recv, int n_pairs, int dims[kMaxDims], int* nd,
                               int* n_blocks, int64_t* block_len) {
  int cd[kMaxDims], cnd = 0;
  if (!plan_choose_dims(n, cd, &cnd)) return false;
  double Lc = 2.0;
  for (int a = 0; a < cnd; ++a) Lc *= cd[a];
  const double N = n_recv, P = n_pairs > 0 ? n_pairs : 1;
  double best = P * (6.0 * Lc + n_lags) + N * (n + 5.0 * Lc);
  bool found = false;
  for (int B = 2; B <= 6; ++B) {
    int bd[kMaxDims], bnd = 0;
    int64_t lb = (n + B - 1) / B;
    if (!plan_choose_dims((lb + n_lags + 1) / 2, bd, &bnd)) continue;
    int64_t M = 1;
    for (int a = 0; a < bnd; ++a) M *= bd[a];
    // TMA-legal pieces: whole rows of the first pass (row = 2*S[1] reals), else
    // whole 128-byte column tiles (32 reals covers both dtypes)
    int64_t row = 2 * (M / bd[0]);
    bool rounded = false;
    if (bnd > 1 && n % row == 0) {
      int64_t lb_rows = (lb + row - 1) / row * row;
      if (lb_rows + n_lags <= 2 * M) { lb = lb_rows; rounded = true; }
    }
    if (!rounded) {
      int64_t lb32 = (lb + 31) / 32 * 32;
      if (lb32 + n_lags <= 2 * M) lb = lb32;
    }
    if (lb + n_lags > 2 * M) continue;
    int Beff = (int)((n + lb - 1) / lb);
    if (Beff < 2) continue;
    const double F = 2.0 * M;
    const double cost = P * ((2.0 * Beff + 4.0) * F + n_lags) + N * 2.0 * Beff * 5.5 * F;
    if (cost < 0.95 * best) {
      best = cost;
      found = true;
      for (int a = 0; a < bnd; ++a) dims[a] = bd[a];
      *nd = bnd;
      *n_blocks = Beff;
      *block_len = lb;
    }
  }
  return found;
}

}  // namespace xcb
