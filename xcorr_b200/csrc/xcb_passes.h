// FFT pass bodies for the packed-real multi-pass correlation.
//
// A transform of M = N_1*...*N_p complex points is done as p passes over HBM.
// Pass a computes length-N_a sub-FFTs along axis a of the C-ordered array
// (N_1,...,N_p); the spectrum stays in digit-transposed order so no pass ever
// transposes through HBM.  Two kernel families:
//
//   ColPass  axis a < p : sub-FFT elements are `inner` apart; a CTA owns a
//            tile of C adjacent columns (C*sizeof(cx<T>) = 128 B) so every
//            global access is a full 128-byte line, and shared memory is
//            [N][C] = one bank sweep per row -> conflict free.
//   RowPass  axis p     : rows are contiguous; a CTA owns C rows, staged
//            through shared memory transposed ([N][C+1]).
//
// Each sub-FFT N = R1*R2 is two register-resident radix stages (Stockham):
// stage 1 reads x[j + r*R2], stage 2 reads s[j + r*R1] * W_N^(r*j).
//
// The bodies are split into phases separated by block barriers and keep all
// per-thread state in ThreadCtx, so that tests/host_emu can run the very same
// code thread-by-thread on the CPU (there is no GPU in the build container).
// That emulator is a test harness only; the product never runs it.
#pragma once
#include <math.h>
#include "xcb_radix_gen.h"
#include "xcb_types.h"

namespace xcb {

enum ColLoad : int32_t { CL_REAL = 0, CL_CPLX = 1, CL_CPLX_INV = 2 };
enum ColStore : int32_t { CS_TW = 0, CS_CONJ = 1, CS_EPI = 2, CS_AMAX = 3 };
enum RowLoad : int32_t { RL_CPLX = 0, RL_REAL = 1, RL_PROLOGUE = 2 };
enum RowStore : int32_t { RS_CPLX = 0, RS_CONJ = 1, RS_EPI = 2, RS_HALF = 3 };

constexpr int kMaxDims = 4;

// One (pair, window) correlation: where its two spectra live and where its
// cropped CCF goes.  Offsets are in elements.
struct Job {
  int64_t spec_i, spec_j;   // complex-element offsets into the spectrum buffer
  int64_t out_off;          // real-element offset into the CCF buffer
  int32_t sig_i, sig_j;     // indices into sumsq[]
};

// Up to 8 jobs of one launch that share spectra (pair product, grouped kernel):
// the distinct first-role / second-role spectra are staged once per row pair in
// shared memory and every job of the group (one warp each) reads them there.
constexpr int kGroupJobs = 8;
// distinct spectra a group may stage (shared-memory ring of the staged kernel: rows x 2 sides x N elements)
template <class T> constexpr int stage_rows() { return sizeof(T) == 4 ? 6 : 4; }
struct PairGroup {
  int32_t n_jobs, n_i, n_j, pad;
  int64_t spec_i[kGroupJobs], spec_j[kGroupJobs];   // distinct spectrum bases (complex-element offsets)
  int32_t job[kGroupJobs];                          // index of the job within the launch
  uint8_t ji[kGroupJobs], jj[kGroupJobs];           // staged row used by each job
  int64_t rows[2 * kGroupJobs];                     // staged rows in order: spec_i[0..n_i) then spec_j[0..n_j)
  int32_t n_rows, pad2;
};

// what a consumer warp of the staged pair product needs to know about its job of a group (copy of PairGroup / Job / norms)
struct alignas(16) PairRec {
  int32_t batch;        // job index within the launch, -1: this warp has no job in the group
  uint8_t row_i, row_j, pad0, pad1;
  double fs;            // 0.5 / (||x_i|| ||x_j||) (0.5 when not normalised or when the norms are per term)
};

template <class T> struct PassParams {
  // ---- geometry ----
  int32_t N;                 // sub-FFT length of this pass
  int32_t load_mode, store_mode;
  int32_t n_batch;           // signals (forward) or jobs (inverse) in this launch
  int64_t M;                 // complex points per transform
  int64_t inner;             // col pass: stride between sub-FFT elements
  int64_t col_tiles;         // col pass: ceil(inner / C)
  int64_t rows;              // row pass: M / N
  int64_t tiles_per_batch;   // CTAs per signal/job
  // ---- buffers ----
  const void* in;            // cx<T>[n_batch][in_stride] (or T[...] for *_REAL)
  void* out;                 // cx<T>[n_batch][out_stride]
  int64_t in_stride, out_stride;
  // ---- real input (forward first pass) ----
  int64_t n_real;            // transform samples [0, n_real) may be non-zero; the rest of 2M is zero
  int64_t real_begin;        // transform sample q is signal sample real_begin + q (zero where that is < 0)
  int32_t real_vec_ok;       // 1: signal sample pairs (real_begin + 2m, +1) are 2*sizeof(T)-aligned
  int32_t pad0;
  double* sumsq_partial;     // [n_batch][tiles_per_batch] or null
  // ---- twiddles: W_M^e = tw_hi[e >> shift] * tw_lo[e & mask] ----
  const cx<T>* tw_hi;
  const cx<T>* tw_lo;
  const cx<T>* tw_inner;     // W_N^t, t in [0, N)
  int32_t tw_shift;
  uint32_t tw_mask;
  uint32_t tw_scale;         // col pass: exponent = tw_scale * column * index
  // ---- half-step twiddles of the real <-> packed conversion: W_L^e (L = 2M), same split ----
  const cx<T>* tw2_hi;
  const cx<T>* tw2_lo;
  const cx<T>* tw2_inner;    // W_{2N}^t, t in [0, N) (row axis)
  // row pass: 1 = a tile is C/2 mirrored row pairs (ra, rb) in columns (2q, 2q+1)
  // (forward last pass with the half-spectrum store, and the pair product)
  int32_t pair_tiles;
  int32_t n_groups;          // pair product: job groups of this launch (0 = ungrouped kernel)
  const PairGroup* groups;
  unsigned long long* work_counter;   // staged pair product: next work item (zeroed before the launch)
  struct PairRec* recs_global;        // per (group, consumer warp) records in the caller's workspace, for launches with
                                      // more groups than the kernel's shared-memory table holds (filled before the launch)
  // ---- pair product (inverse first pass) ----
  int32_t nd;                // number of passes p
  int32_t dims[kMaxDims];    // N_1..N_p
  int32_t n_regions;         // p - 1 mirror regions over rows
  int32_t normalize;
  int64_t region_lo[kMaxDims], region_hi[kMaxDims];  // row ranges
  int64_t pair_start[kMaxDims + 1];                  // first pair id of region g (pair 0 = row 0)
  int64_t n_pairs_rows;      // total row pairs
  const Job* jobs;           // [n_batch]
  const double* rnorm;       // per signal: 1/||x|| (= rsqrt of the sum of squares)
  const cx<T>* spec;         // spectrum buffer base
  // spectral stacking: the product is summed over n_stack windows before the
  // inverse transform; window w of a job adds w*stack_spec_stride elements to
  // its spectrum offsets and w*stack_sig_stride to its signal indices
  int32_t n_stack;
  int32_t stack_sig_stride;
  int64_t stack_spec_stride;
  // block mode (n_stack <= kMaxBlocks): explicit per-block offsets instead of the
  // stride, so that a segment whose samples equal a block's can share its spectrum
  int32_t use_blk_off;
  int32_t pad2;
  int64_t blk_off_i[8], blk_off_j[8];
  // ---- lag epilogue (inverse last pass) ----
  int64_t n;                 // samples per window
  int64_t L;                 // 2M
  int64_t i0, i1;            // kept output indices [i0, i1) of the 2n-1 lags
  T* ccf;                    // CCF buffer base (jobs[].out_off applies)
  T escale;                  // 1/M
  int32_t unbiased;          // multiply by n/(n-|k|)
  int32_t accumulate;        // ccf += instead of ccf =
  // last inverse column pass: output rows k (element m = k*inner + col) that
  // can hold a kept lag: [keep_lo[0], keep_hi[0]) U [keep_lo[1], keep_hi[1])
  int32_t keep_lo[2], keep_hi[2];
  // block mode: circular lag l of the (short) transform IS kept lag l (output index l)
  int32_t epi_direct;
  // row pass: 1 = consecutive CTAs take the SAME tile of consecutive jobs (tile
  // major).  For the pair product all pairs of a frequency tile then run
  // together and the spectra rows they share are served by L2 instead of HBM.
  int32_t tile_major;
  // ---- fused arg-max epilogue (CS_AMAX: the kept lags are reduced, not stored) ----
  // per tile: the first maximal kept output index (np.argmax order: a NaN wins, ties -> lowest index) and its
  // value; -1 when the tile holds no kept lag.  A second kernel reduces the tiles of a job.
  void* amax_val;            // T[n_batch][tiles_per_batch]
  int64_t* amax_idx;         // [n_batch][tiles_per_batch]
  // 2-D validity of lag = dF * mask_w + dT (row-embedded images, signal/correlate.py): kept iff |dT| < mask_t
  // and |dF| < mask_f; mask_w == 0 keeps every lag of [i0, i1)
  int32_t mask_w, mask_t, mask_f;
  int32_t mask_shift;        // log2(mask_w) when mask_w is a power of two, else -1
};

template <class T, int R1_, int R2_> struct Cfg {
  static constexpr int R1 = R1_, R2 = R2_;
  static constexpr int N = R1 * R2;
  static constexpr int C = 128 / (2 * (int)sizeof(T));   // 16 (f32) / 8 (f64)
  static constexpr int TJ = R1 > R2 ? R1 : R2;
  static constexpr int THREADS = C * TJ;
  static constexpr int CP = C + 1;                        // padded row-pass pitch
  static constexpr int COL_SMEM = N * C * 2 * (int)sizeof(T) + THREADS * 8;
  static constexpr int ROW_SMEM = N * CP * 2 * (int)sizeof(T) + C * 32 + THREADS * 8;
};

template <class T, int RMAX> struct ThreadCtx {
  int32_t tid;
  int64_t bid;
  cx<T> v[RMAX];
  double ss;
};

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
template <class T> XCB_HD cx<T> ld_ro(const cx<T>* p) {
#if defined(__CUDA_ARCH__)
  if constexpr (sizeof(T) == 4) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    cx<T> r; r.x = t.x; r.y = t.y; return r;
  } else {
    double2 t = __ldg(reinterpret_cast<const double2*>(p));
    cx<T> r; r.x = t.x; r.y = t.y; return r;
  }
#else
  return *p;
#endif
}

template <class T> XCB_HD cx<T> tw_big(const PassParams<T>& p, uint32_t e) {
  return cmul(ld_ro(p.tw_hi + (e >> p.tw_shift)), ld_ro(p.tw_lo + (e & p.tw_mask)));
}

// packed real load: complex index m of the transform -> samples (2m, 2m+1) of the
// window [real_begin, real_begin + n_real) of signal `x`; zero outside it and
// before the start of the signal
template <class T> XCB_HD cx<T> load_real_packed(const PassParams<T>& p, const T* x, int64_t m, double& ss) {
  cx<T> r; r.x = T(0); r.y = T(0);
  const int64_t q = 2 * m, i = p.real_begin + q;
  if (q + 1 < p.n_real && i >= 0) {
    if (p.real_vec_ok) r = *reinterpret_cast<const cx<T>*>(x + i);
    else { r.x = x[i]; r.y = x[i + 1]; }
  } else {
    if (q < p.n_real && i >= 0) r.x = x[i];
    if (q + 1 < p.n_real && i + 1 >= 0) r.y = x[i + 1];
  }
  ss += (double)(r.x * r.x + r.y * r.y);
  return r;
}

// final packed element m = (r[2m], r[2m+1]) -> cropped, scaled lag samples.
// All indices fit 32 bits (L = 2M < 2^32).
template <class T> XCB_HD void epilogue_store(const PassParams<T>& p, T* outp, uint32_t m, T re, T im) {
  const uint32_t n = (uint32_t)p.n, L = (uint32_t)p.L;
  const uint32_t i0 = (uint32_t)p.i0, K = (uint32_t)(p.i1 - p.i0);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t l = 2u * m + (uint32_t)h;
    uint32_t i, o;
    if (p.epi_direct) {
      o = l;
      i = l + i0;
    } else {
      if (l < n) i = l + (n - 1u);
      else if (l > L - n) i = l - (L - n + 1u);
      else continue;
      o = i - i0;                   // wraps to a huge value when i < i0
    }
    if (o >= K) continue;
    T val = (h ? im : re) * p.escale;
    if (p.unbiased) {
      uint32_t k = i >= n - 1u ? i - (n - 1u) : (n - 1u) - i;
      val *= T((double)n / (double)(n - k));
    }
    T* dst = outp + o;
    *dst = p.accumulate ? (*dst + val) : val;
  }
}

// Lean epilogue for the common block-mode case (lag l = output index l, no
// weights, plain store): one aligned 2-sample store per complex element.
template <class T> XCB_HD bool epilogue_fast_ok(const PassParams<T>& p) {
  return p.epi_direct && !p.unbiased && !p.accumulate;
}
template <class T> XCB_HD void epilogue_store_fast(const PassParams<T>& p, T* outp, uint32_t m, T re, T im) {
  const uint32_t K = (uint32_t)(p.i1 - p.i0), l = 2u * m;
  if (l + 1u < K) {
    cx<T> v; v.x = re * p.escale; v.y = im * p.escale;
    *reinterpret_cast<cx<T>*>(outp + l) = v;           // l even, rows start 2*sizeof(T)-aligned (K even or job 0)
  } else if (l < K) {
    outp[l] = re * p.escale;
  }
}

// ---- fused arg-max (CS_AMAX) ----
constexpr uint32_t kAmaxNone = 0xffffffffu;
// np.argmax order on (value, output index) candidates: a NaN beats any number, then the larger value, then the lower index
template <class T> XCB_HD bool amax_better(T vb, uint32_t ob, T va, uint32_t oa) {
  if (ob == kAmaxNone) return false;
  if (oa == kAmaxNone) return true;
  const bool nb = vb != vb, na = va != va;
  if (nb != na) return nb;
  if (nb) return ob < oa;
  return vb > va || (vb == va && ob < oa);
}
// circular lag l of the transform with raw value x: its cropped output index and scaled value, if the lag is kept
// (inside [i0, i1), not masked)
template <class T> XCB_HD bool amax_lag(const PassParams<T>& p, uint32_t l, T x, uint32_t& o, T& val,
                                        bool prescaled = false) {
  const uint32_t n = (uint32_t)p.n, L = (uint32_t)p.L;
  const uint32_t i0 = (uint32_t)p.i0, K = (uint32_t)(p.i1 - p.i0);
  uint32_t i;
  if (p.epi_direct) {
    o = l;
    i = l + i0;
  } else {
    if (l < n) i = l + (n - 1u);
    else if (l > L - n) i = l - (L - n + 1u);
    else return false;
    o = i - i0;
  }
  if (o >= K) return false;
  if (p.mask_w) {
    // lag = i - (n-1) = dF * W + dT;  t = lag + (T-1) + F*W > 0  ->  dF + F = t / W,  dT + T-1 = t % W
    const uint32_t W = (uint32_t)p.mask_w, Tm = (uint32_t)p.mask_t, F = (uint32_t)p.mask_f;
    const uint32_t t = i + (Tm - 1u) + F * W - (n - 1u);
    const uint32_t q = p.mask_shift >= 0 ? (t >> p.mask_shift) : t / W;
    const uint32_t dt = t - q * W;
    if (dt >= 2u * Tm - 1u || q < 1u || q > 2u * F - 1u) return false;
  }
  val = prescaled ? x : x * p.escale;
  if (p.unbiased) {
    const uint32_t k = i >= n - 1u ? i - (n - 1u) : (n - 1u) - i;
    val *= T((double)n / (double)(n - k));
  }
  return true;
}
// final packed element m = (r[2m], r[2m+1]): offer its kept, unmasked lags to the running candidate (bv, bo)
template <class T> XCB_HD void amax_consider(const PassParams<T>& p, uint32_t m, T re, T im, T& bv, uint32_t& bo) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t o;
    T val;
    if (amax_lag(p, 2u * m + (uint32_t)h, h ? im : re, o, val) && amax_better(val, o, bv, bo)) { bv = val; bo = o; }
  }
}

// v[r] <- (conj?)(v[r]) * W_M^(e0 + r*de), r = 0..R-1, with the twiddle advanced
// by a running product instead of R table lookups (2 lookups per thread).
template <int R, bool CONJ_IN, class T>
XCB_HD void twiddle_chain(const PassParams<T>& p, uint32_t e0, uint32_t de, cx<T>* v) {
  cx<T> w = tw_big(p, e0);
  const cx<T> step = tw_big(p, de);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    cx<T> x = v[r];
    if (CONJ_IN) x.y = -x.y;
    v[r] = cmul(x, w);
    if (r + 1 < R) w = cmul(w, step);
  }
}

// the same with the two twiddles already looked up (the caller fetched them ahead of a wait)
template <int R, bool CONJ_IN, class T> XCB_HD void twiddle_chain_from(cx<T> w, const cx<T> step, cx<T>* v) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    cx<T> x = v[r];
    if (CONJ_IN) x.y = -x.y;
    v[r] = cmul(x, w);
    if (r + 1 < R) w = cmul(w, step);
  }
}

// v[r] <- v[r] * w^r, r = 1..R-1 (stage-2 twiddles W_N^(r*j), w = W_N^j)
template <int R, bool PK = false, class T> XCB_HD void twiddle_powers(cx<T> w, cx<T>* v) {
  cx<T> wr = w;
#pragma unroll
  for (int r = 1; r < R; ++r) {
    v[r] = xmul<PK>(v[r], wr);
    if (r + 1 < R) wr = xmul<PK>(wr, w);
  }
}


// per-CTA sum of the threads' partial sums of squares (deterministic order);
// `red` holds one double per thread, written before the preceding barrier.
template <int THREADS> XCB_HD void block_sumsq(double* out, const double* red, int tid) {
#if defined(__CUDA_ARCH__)
  if (tid < 32) {
    double a = 0.0;
    for (int i = tid; i < THREADS; i += 32) a += red[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (tid == 0) *out = a;
  }
#else
  if (tid == 0) {
    // same association order as the device path: 32 strided lanes, then a tree
    double lane[32];
    for (int l = 0; l < 32; ++l) { lane[l] = 0.0; for (int i = l; i < THREADS; i += 32) lane[l] += red[i]; }
    for (int o = 16; o > 0; o >>= 1) for (int l = 0; l < o; ++l) lane[l] += lane[l + o];
    *out = lane[0];
  }
#endif
}

// ---------------------------------------------------------------------------
// Column pass
// ---------------------------------------------------------------------------
template <class T, int R1, int R2> struct ColPass {
  using G = Cfg<T, R1, R2>;
  using Ctx = ThreadCtx<T, G::TJ>;
  static constexpr int N = G::N, C = G::C;

  struct Tile {
    int64_t batch, base, col0;   // base: element offset of (n=0, c=0) within the signal
    bool col_ok;
    int c, j;
  };

  static XCB_HD Tile decode(const PassParams<T>& p, const Ctx& ctx) {
    Tile t;
    t.c = ctx.tid % C;
    t.j = ctx.tid / C;
    t.batch = ctx.bid / p.tiles_per_batch;
    int64_t rem = ctx.bid - t.batch * p.tiles_per_batch;
    int64_t o = rem / p.col_tiles;
    int64_t cb = rem - o * p.col_tiles;
    t.col0 = cb * C;
    t.base = o * (int64_t)N * p.inner + t.col0;
    t.col_ok = (t.col0 + t.c) < p.inner;
    return t;
  }

  // stage 1: global -> registers -> radix-R1 -> shared
  static XCB_HD void phase_a(const PassParams<T>& p, Ctx& ctx, unsigned char* smem_raw) {
    cx<T>* smem = reinterpret_cast<cx<T>*>(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw + (size_t)N * C * sizeof(cx<T>));
    Tile t = decode(p, ctx);
    ctx.ss = 0.0;
    if (t.j < R2 && t.col_ok) {
      cx<T>* v = ctx.v;
      if (p.load_mode == CL_REAL) {
        const T* x = reinterpret_cast<const T*>(p.in) + t.batch * p.in_stride;
#pragma unroll
        for (int r = 0; r < R1; ++r) {
          int64_t m = t.base + (int64_t)(t.j + r * R2) * p.inner + t.c;
          v[r] = load_real_packed(p, x, m, ctx.ss);
        }
      } else {
        const cx<T>* src = reinterpret_cast<const cx<T>*>(p.in) + t.batch * p.in_stride + t.base + t.c;
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = src[(int64_t)(t.j + r * R2) * p.inner];
        if (p.load_mode == CL_CPLX_INV) {
          uint32_t col = (uint32_t)(t.col0 + t.c) * p.tw_scale;
          twiddle_chain<R1, true>(p, col * (uint32_t)t.j, col * (uint32_t)R2, v);
        }
      }
      fft_reg<R1>(v);
#pragma unroll
      for (int r = 0; r < R1; ++r) smem[(t.j * R1 + r) * C + t.c] = v[r];
    }
    if (p.sumsq_partial) red[ctx.tid] = ctx.ss;
  }

  // stage 2: shared -> registers -> radix-R2 -> global
  static XCB_HD void phase_b(const PassParams<T>& p, Ctx& ctx, unsigned char* smem_raw) {
    cx<T>* smem = reinterpret_cast<cx<T>*>(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw + (size_t)N * C * sizeof(cx<T>));
    Tile t = decode(p, ctx);
    if (t.j < R1 && t.col_ok) {
      cx<T>* v = ctx.v;
#pragma unroll
      for (int r = 0; r < R2; ++r) v[r] = smem[(t.j + r * R1) * C + t.c];
      twiddle_powers<R2>(ld_ro(p.tw_inner + t.j), v);
      fft_reg<R2>(v);
      if (p.store_mode == CS_TW) {
        cx<T>* dst = reinterpret_cast<cx<T>*>(p.out) + t.batch * p.out_stride + t.base + t.c;
        uint32_t col = (uint32_t)(t.col0 + t.c) * p.tw_scale;
        twiddle_chain<R2, false>(p, col * (uint32_t)t.j, col * (uint32_t)R1, v);
#pragma unroll
        for (int r = 0; r < R2; ++r) dst[(int64_t)(t.j + r * R1) * p.inner] = v[r];
      } else if (p.store_mode == CS_CONJ) {
        cx<T>* dst = reinterpret_cast<cx<T>*>(p.out) + t.batch * p.out_stride + t.base + t.c;
#pragma unroll
        for (int r = 0; r < R2; ++r) dst[(int64_t)(t.j + r * R1) * p.inner] = cconj(v[r]);
      } else {
        const int64_t out_off = p.jobs[t.batch].out_off;
        T* outp = p.ccf + out_off;
        const uint32_t m0 = (uint32_t)(t.base + t.c), inner = (uint32_t)p.inner;
        if (epilogue_fast_ok(p) && (out_off & 1) == 0) {
#pragma unroll
          for (int r = 0; r < R2; ++r) {
            const int k = t.j + r * R1;
            if (k < p.keep_hi[0]) epilogue_store_fast(p, outp, m0 + (uint32_t)k * inner, v[r].x, -v[r].y);
          }
        } else {
#pragma unroll
          for (int r = 0; r < R2; ++r) {
            const int k = t.j + r * R1;
            if ((k >= p.keep_lo[0] && k < p.keep_hi[0]) || (k >= p.keep_lo[1] && k < p.keep_hi[1]))
              epilogue_store(p, outp, m0 + (uint32_t)k * inner, v[r].x, -v[r].y);
          }
        }
      }
    }
    if (p.sumsq_partial) block_sumsq<G::THREADS>(p.sumsq_partial + ctx.bid, red, ctx.tid);
  }
};

// ---------------------------------------------------------------------------
// Row pass
// ---------------------------------------------------------------------------
template <class T> struct RowInfo {   // per tile column, kept in shared memory
  int64_t row;      // global row index, or -1 if the column is unused
  int32_t store;    // 1: this column's result is written back
  int32_t pad;
  cx<T> wrow;       // W_M^{klow(row)} for the pair product
};

template <class T, int R1, int R2> struct RowPass {
  using G = Cfg<T, R1, R2>;
  using Ctx = ThreadCtx<T, G::TJ>;
  static constexpr int N = G::N, C = G::C, CP = G::CP;

  static XCB_HD RowInfo<T>* info_ptr(unsigned char* smem_raw) {
    return reinterpret_cast<RowInfo<T>*>(smem_raw + (size_t)N * CP * sizeof(cx<T>));
  }
  static XCB_HD double* red_ptr(unsigned char* smem_raw) {
    return reinterpret_cast<double*>(smem_raw + (size_t)N * CP * sizeof(cx<T>) + C * 32);
  }

  // frequency index contributed by the leading p-1 digits of a row
  static XCB_HD uint32_t klow_of_row(const PassParams<T>& p, int64_t row) {
    // row = ((k_1 N_2 + k_2) N_3 + ...) + k_{p-1};  klow = sum k_a * (N_1..N_{a-1})
    uint32_t digits[kMaxDims];
    uint32_t r = (uint32_t)row;            // rows < M < 2^31
    for (int a = p.nd - 2; a >= 0; --a) {
      const uint32_t d = (uint32_t)p.dims[a], qq = r / d;
      digits[a] = r - qq * d;
      r = qq;
    }
    uint32_t k = 0, base = 1;
    for (int a = 0; a < p.nd - 1; ++a) { k += digits[a] * base; base *= (uint32_t)p.dims[a]; }
    return k;
  }

  static XCB_HD void decode(const PassParams<T>& p, int64_t bid, int64_t& batch, int64_t& tile) {
    if (p.tile_major) { tile = bid / p.n_batch; batch = bid - tile * p.n_batch; }
    else { batch = bid / p.tiles_per_batch; tile = bid - batch * p.tiles_per_batch; }
  }

  // row pair q -> its two mirrored rows (pair 0 = row 0 with itself); false past the last pair
  static XCB_HD bool pair_rows(const PassParams<T>& p, int64_t q, int64_t& ra, int64_t& rb) {
    ra = rb = 0;
    if (q >= p.n_pairs_rows) return false;
    if (q > 0) {
      int g = 0;
      while (g + 1 < p.n_regions && q >= p.pair_start[g + 1]) ++g;
      const int64_t tt = q - p.pair_start[g];
      ra = p.region_lo[g] + tt;
      rb = p.region_hi[g] - 1 - tt;
    }
    return true;
  }

  // column c of tile `tile` -> global row.  Pair tiles: columns (2q, 2q+1) hold
  // the mirrored rows (ra, rb) of pair slot q: element e of ra is frequency k,
  // element em of rb is frequency M - k.
  static XCB_HD RowInfo<T> row_info(const PassParams<T>& p, int64_t tile, int c) {
    RowInfo<T> ri;
    ri.row = -1; ri.store = 0; ri.pad = 0; ri.wrow.x = T(1); ri.wrow.y = T(0);
    if (p.pair_tiles) {
      int64_t ra, rb;
      if (pair_rows(p, tile * (C / 2) + (c >> 1), ra, rb)) {
        const bool self_dup = (c & 1) && ra == rb;     // a self-mirrored row lives in column 2q only
        ri.row = (c & 1) ? rb : ra;
        ri.store = self_dup ? 0 : 1;
        // forward: the duplicate column is not loaded at all (it would count twice in the norms)
        if (self_dup && p.load_mode != RL_PROLOGUE) ri.row = -1;
        // both columns of a pair carry the half-step twiddle W_L^klow of the first row
        const uint32_t kl = klow_of_row(p, ra);
        ri.wrow = cmul(ld_ro(p.tw2_hi + (kl >> p.tw_shift)), ld_ro(p.tw2_lo + (kl & p.tw_mask)));
      }
    } else {
      int64_t row = tile * C + c;
      if (row < p.rows) { ri.row = row; ri.store = 1; }
    }
    return ri;
  }

  // phase 0: column -> row bookkeeping
  static XCB_HD void phase_info(const PassParams<T>& p, Ctx& ctx, unsigned char* smem_raw) {
    if (ctx.tid >= C) return;
    int64_t batch, tile;
    decode(p, ctx.bid, batch, tile);
    info_ptr(smem_raw)[ctx.tid] = row_info(p, tile, ctx.tid);
  }

  // Spectra are stored as HALF SPECTRA of the real signals (store mode RS_HALF):
  // address of frequency k holds X[k], k = 0..M-1, slot 0 holds (X[0], X[M]) (both
  // real).  With Z the packed transform, S = Z[k] + conj(Z[M-k]), D = Z[k] -
  // conj(Z[M-k]) and W = W_L^k:   X[k] = (S - i W D)/2,  X[M-k] = conj((S + i W D)/2).
  static XCB_HD void untangle(const cx<T>& Zk, const cx<T>& Zm, const cx<T>& W, cx<T>& Xk, cx<T>& Xm) {
    cx<T> S, D;
    S.x = Zk.x + Zm.x; S.y = Zk.y - Zm.y;
    D.x = Zk.x - Zm.x; D.y = Zk.y + Zm.y;
    const cx<T> WD = cmul(W, D);                   // -i W D = (WD.y, -WD.x)
    Xk.x = T(0.5) * (S.x + WD.y); Xk.y = T(0.5) * (S.y - WD.x);
    Xm.x = T(0.5) * (S.x - WD.y); Xm.y = T(-0.5) * (S.y + WD.x);
  }

  // Cross-spectrum C[k] = sum X_i[k] conj(X_j[k]) of a real correlation r (length
  // L) -> packed spectrum V of z[m] = r[2m] + i r[2m+1]:  E = (C[k] + conj(C[M-k]))/2,
  // O = conj(W_L^k) (C[k] - conj(C[M-k]))/2,  V[k] = E + iO,  V[M-k] = conj(E) + i conj(O).
  // Returns fs * conj(V[k]) and fs * conj(V[M-k]) with fs carrying the 1/2 (the
  // inverse runs as conj . FFT . conj).
  static XCB_HD void tangle(const cx<T>& Ck, const cx<T>& Cm, const cx<T>& W, T fs, cx<T>& va, cx<T>& vb) {
    cx<T> E, D;
    E.x = Ck.x + Cm.x; E.y = Ck.y - Cm.y;
    D.x = Ck.x - Cm.x; D.y = Ck.y + Cm.y;
    const cx<T> O = cmulc(D, W);                   // conj(W) D
    va.x = fs * (E.x - O.y); va.y = -fs * (E.y + O.x);
    vb.x = fs * (E.x + O.y); vb.y = fs * (E.y - O.x);
  }

  // Thread <-> data mapping of the cooperative phases: every column c of the
  // tile (= one global row) is owned by TJ consecutive threads which stride over
  // its N elements (N / TJ = RMIN iterations, exact).  Row bookkeeping is read
  // once per thread; global accesses of a warp are contiguous runs of TJ
  // elements; shared accesses have pitch CP (conflict free).
  static constexpr int TJ = G::TJ;
  static constexpr int RMIN = N / TJ;             // = min(R1, R2)
  static constexpr int UB = 8;

  // phase 1: global -> shared (transposed)
  // MODE 0: role from p.load_mode (emulator); 1: forward / plain complex pass; 2: pair product.  The kernels
  // are compiled per role.
  template <int MODE = 0>
  static XCB_HD void phase_load(const PassParams<T>& p, Ctx& ctx, unsigned char* smem_raw) {
    cx<T>* smem = reinterpret_cast<cx<T>*>(smem_raw);
    const RowInfo<T>* info = info_ptr(smem_raw);
    int64_t batch, tile_;
    decode(p, ctx.bid, batch, tile_);
    ctx.ss = 0.0;
    if (MODE == 2 || (MODE == 0 && p.load_mode == RL_PROLOGUE)) {
      // Pair slot lp (rows ra, rb) is owned by 2*TJ threads; nothing here reads
      // `info`, so the kernel does not wait for phase_info before loading.
      constexpr int TS = 2 * TJ;
      constexpr int ITER = (N + TS - 1) / TS;
      constexpr bool EXACT = (N % TS) == 0;
      constexpr int PB = (ITER >= 4 && sizeof(T) == 4) ? 4 : (ITER >= 2 ? 2 : 1);
      constexpr bool WHOLE = (ITER % PB) == 0;           // else the last batch is partial
      const int lp = ctx.tid / TS, t = ctx.tid - lp * TS;
      int64_t ra, rb;
      if (!pair_rows(p, tile_ * (C / 2) + lp, ra, rb)) {
        cx<T> z; z.x = z.y = T(0);
        for (int e = t; e < N; e += TS) { smem[e * CP + 2 * lp] = z; smem[e * CP + 2 * lp + 1] = z; }
        return;
      }
      const bool row0 = ra == 0;
      const int emb = row0 ? N : N - 1;                  // mirrored element: emb - e (0 for element 0 of row 0)
      const uint32_t kl = klow_of_row(p, ra);
      const cx<T> wh = ld_ro(p.tw2_hi + (kl >> p.tw_shift)), wl = ld_ro(p.tw2_lo + (kl & p.tw_mask));
      const Job jb = p.jobs[batch];
      // one scale for all terms (block mode, plain pairs) is applied once at the end
      const bool per_term = p.normalize && p.stack_sig_stride != 0;
      T fs = T(0.5);
      if (p.normalize && !per_term) fs = T(0.5 * p.rnorm[jb.sig_i] * p.rnorm[jb.sig_j]);
      const cx<T>* za = p.spec + ra * N + t;             // frequency k side
      const cx<T>* zb = p.spec + rb * N + (emb - t);     // frequency M - k side (walks backwards)
      const int fix0 = (row0 && t == 0) ? -N : 0;        // element 0 of row 0 mirrors itself
      const cx<T> wrow = cmul(wh, wl);
      for (int it0 = 0; it0 < ITER; it0 += PB) {
        cx<T> ca[PB], cb[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) ca[u].x = ca[u].y = cb[u].x = cb[u].y = T(0);
        for (int w = 0; w < p.n_stack; ++w) {
          const int64_t oi = jb.spec_i + (p.use_blk_off ? p.blk_off_i[w] : w * p.stack_spec_stride);
          const int64_t oj = jb.spec_j + (p.use_blk_off ? p.blk_off_j[w] : w * p.stack_spec_stride);
          T sw = T(1);
          if (per_term) sw = T(p.rnorm[jb.sig_i + w * p.stack_sig_stride] * p.rnorm[jb.sig_j + w * p.stack_sig_stride]);
          cx<T> A[PB], Am[PB], B[PB], Bm[PB];
#pragma unroll
          for (int u = 0; u < PB; ++u) {
            const int de = (it0 + u) * TS;
            if ((WHOLE || it0 + u < ITER) && (EXACT || t + de < N)) {
              const int dm = (it0 + u == 0) ? fix0 - de : -de;
              A[u] = za[oi + de]; Am[u] = zb[oi + dm]; B[u] = za[oj + de]; Bm[u] = zb[oj + dm];
            }
          }
#pragma unroll
          for (int u = 0; u < PB; ++u) {
            if ((WHOLE || it0 + u < ITER) && (EXACT || t + (it0 + u) * TS < N)) {
              const T ax = A[u].x * sw, ay = A[u].y * sw, mx = Am[u].x * sw, my = Am[u].y * sw;
              ca[u].x += ax * B[u].x + ay * B[u].y;  ca[u].y += ay * B[u].x - ax * B[u].y;
              cb[u].x += mx * Bm[u].x + my * Bm[u].y; cb[u].y += my * Bm[u].x - mx * Bm[u].y;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
          const int e = t + (it0 + u) * TS;
          if ((WHOLE || it0 + u < ITER) && (EXACT || e < N)) {
            const int em = (it0 + u == 0 && fix0) ? 0 : emb - e;
            cx<T> va, vb;
            tangle(ca[u], cb[u], cmul(wrow, ld_ro(p.tw2_inner + e)), fs, va, vb);
            smem[e * CP + 2 * lp] = va;
            smem[em * CP + 2 * lp + 1] = vb;
          }
        }
      }
      if (row0 && t == 0) {
        // slot 0 holds (X[0], X[M]), both real: C[0] and C[M] are plain products and
        // V[0] = (C[0] + C[M])/2 + i (C[0] - C[M])/2
        T c0 = T(0), cm = T(0);
        for (int w = 0; w < p.n_stack; ++w) {
          const cx<T> a = p.spec[jb.spec_i + (p.use_blk_off ? p.blk_off_i[w] : w * p.stack_spec_stride)];
          const cx<T> b = p.spec[jb.spec_j + (p.use_blk_off ? p.blk_off_j[w] : w * p.stack_spec_stride)];
          T sw = T(1);
          if (per_term) sw = T(p.rnorm[jb.sig_i + w * p.stack_sig_stride] * p.rnorm[jb.sig_j + w * p.stack_sig_stride]);
          c0 += a.x * sw * b.x;
          cm += a.y * sw * b.y;
        }
        cx<T> v0; v0.x = fs * (c0 + cm); v0.y = -fs * (c0 - cm);
        smem[2 * lp] = v0;
      }
    } else if (MODE != 2 && p.load_mode == RL_REAL) {
      const T* x = reinterpret_cast<const T*>(p.in) + batch * p.in_stride;
      for (int idx = ctx.tid; idx < C * N; idx += G::THREADS) {
        int c = idx / N, e = idx - c * N;
        cx<T> v; v.x = v.y = T(0);
        int64_t row = info[c].row;
        if (row >= 0) v = load_real_packed(p, x, row * N + e, ctx.ss);
        smem[e * CP + c] = v;
      }
      if (p.sumsq_partial) red_ptr(smem_raw)[ctx.tid] = ctx.ss;
    } else {
      const int c = ctx.tid / TJ, t = ctx.tid - c * TJ;
      const int64_t row = info[c].row;
      const cx<T>* src = reinterpret_cast<const cx<T>*>(p.in) + batch * p.in_stride + (row >= 0 ? row : 0) * N + t;
      cx<T>* dst = smem + t * CP + c;
      for (int it0 = 0; it0 < RMIN; it0 += UB) {
        cx<T> buf[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          buf[u].x = buf[u].y = T(0);
          if (it0 + u < RMIN && row >= 0) buf[u] = src[(it0 + u) * TJ];
        }
#pragma unroll
        for (int u = 0; u < UB; ++u)
          if (it0 + u < RMIN) dst[(it0 + u) * TJ * CP] = buf[u];
      }
    }
  }

  static XCB_HD void phase_s1_read(const PassParams<T>&, Ctx& ctx, unsigned char* smem_raw) {
    const cx<T>* smem = reinterpret_cast<const cx<T>*>(smem_raw);
    int c = ctx.tid % C, j = ctx.tid / C;
    if (j < R2) {
#pragma unroll
      for (int r = 0; r < R1; ++r) ctx.v[r] = smem[(j + r * R2) * CP + c];
    }
  }

  static XCB_HD void phase_s1_write(const PassParams<T>&, Ctx& ctx, unsigned char* smem_raw) {
    cx<T>* smem = reinterpret_cast<cx<T>*>(smem_raw);
    int c = ctx.tid % C, j = ctx.tid / C;
    if (j < R2) {
      fft_reg<R1>(ctx.v);
#pragma unroll
      for (int r = 0; r < R1; ++r) smem[(j * R1 + r) * CP + c] = ctx.v[r];
    }
  }

  static XCB_HD void phase_s2(const PassParams<T>& p, Ctx& ctx, unsigned char* smem_raw) {
    cx<T>* smem = reinterpret_cast<cx<T>*>(smem_raw);
    int c = ctx.tid % C, j = ctx.tid / C;
    if (j < R1) {
      cx<T>* v = ctx.v;
#pragma unroll
      for (int r = 0; r < R2; ++r) v[r] = smem[(j + r * R1) * CP + c];
      twiddle_powers<R2>(ld_ro(p.tw_inner + j), v);
      fft_reg<R2>(v);
#pragma unroll
      for (int r = 0; r < R2; ++r) smem[(j + r * R1) * CP + c] = v[r];
    }
  }

  template <int MODE = 0>
  static XCB_HD void phase_store(const PassParams<T>& p, Ctx& ctx, unsigned char* smem_raw) {
    const RowInfo<T>* info = info_ptr(smem_raw);
    const cx<T>* smem = reinterpret_cast<const cx<T>*>(smem_raw);
    int64_t batch, tile_;
    decode(p, ctx.bid, batch, tile_);
    if (MODE != 2 && p.store_mode == RS_HALF) {
      // forward last pass: packed transform Z of the pair tile -> half spectrum X
      const int lp = ctx.tid / (2 * TJ), t = ctx.tid - lp * (2 * TJ);
      const RowInfo<T> ia = info[2 * lp];
      const int64_t rb = info[2 * lp + 1].row;
      if (ia.row >= 0) {
        const bool self = rb < 0, row0 = ia.row == 0;
        const int cm = self ? 2 * lp : 2 * lp + 1;
        cx<T>* base = reinterpret_cast<cx<T>*>(p.out) + batch * p.out_stride;
        cx<T>* dst_a = base + ia.row * N;
        cx<T>* dst_b = base + (self ? ia.row : rb) * N;
        constexpr int TS = 2 * TJ;
        constexpr int ITER = (N + TS - 1) / TS;
        const int emb = row0 ? N : N - 1;
        const cx<T>* sa = smem + t * CP + 2 * lp;
        const cx<T>* sb = smem + (emb - t) * CP + cm;
        dst_a += t;
        dst_b += emb - t;
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
          const int de = it * TS;
          if ((N % TS) == 0 || t + de < N) {
            const bool first = it == 0 && row0 && t == 0;        // element 0 of row 0 mirrors itself
            cx<T> Xk, Xm;
            untangle(sa[de * CP], first ? sa[0] : sb[-de * CP], cmul(ia.wrow, ld_ro(p.tw2_inner + t + de)), Xk, Xm);
            if (first) Xk.y = Xm.x;                              // slot 0 = (X[0], X[M])
            dst_a[de] = Xk;
            if (!self) dst_b[-de] = Xm;
          }
        }
      }
    } else if (p.store_mode == RS_EPI) {
      T* outp = p.ccf + p.jobs[batch].out_off;
      for (int idx = ctx.tid; idx < C * N; idx += G::THREADS) {
        int c = idx / N, e = idx - c * N;
        const RowInfo<T> ri = info[c];
        if (ri.row < 0 || !ri.store) continue;
        cx<T> v = smem[e * CP + c];
        epilogue_store(p, outp, (uint32_t)(ri.row * N + e), v.x, -v.y);
      }
    } else {
      const int c = ctx.tid / TJ, t = ctx.tid - c * TJ;
      const RowInfo<T> ri = info[c];
      if (ri.row >= 0 && ri.store) {
        cx<T>* dst = reinterpret_cast<cx<T>*>(p.out) + batch * p.out_stride + ri.row * N + t;
        const cx<T>* src = smem + t * CP + c;
        const T sgn = p.store_mode == RS_CONJ ? T(-1) : T(1);
#pragma unroll
        for (int it = 0; it < RMIN; ++it) {
          cx<T> v = src[it * TJ * CP];
          v.y *= sgn;
          dst[it * TJ] = v;
        }
      }
    }
    if (MODE != 2 && p.sumsq_partial && p.load_mode == RL_REAL)
      block_sumsq<G::THREADS>(p.sumsq_partial + ctx.bid, red_ptr(smem_raw), ctx.tid);
  }
};

// ---------------------------------------------------------------------------
// Row pass, warp-synchronous variant (R1 == 16)
// ---------------------------------------------------------------------------
// One warp owns one mirrored row pair: lanes 0-15 the row ra (frequencies k),
// lanes 16-31 its mirror rb (frequencies M - k).  Lane j of a half loads the
// stage-1 inputs e = j + R2 r of its row straight from HBM into registers (a
// half-warp reads 16 consecutive elements = one 128-byte line per r), runs
// radix-R1, swaps through a padded per-row buffer in shared memory, runs
// radix-R2 and stores k = j + R1 r straight from registers (again full lines).
// The only synchronisation is __syncwarp: warps never wait for each other, and
// every element makes ONE shared-memory round trip per exchange instead of
// three.  The real <-> packed conversion needs the mirrored element (row rb,
// N-1-e): both halves publish their values to the row buffers (linear index)
// and read the partner's.
//
//   forward last pass : A load, radix-R1, write stage layout | B read, twiddle,
//                       radix-R2 | C publish Z | D read mirror, untangle, store X
//   pair product      : A X_i conj(X_j) summed over blocks / stacked windows,
//                       publish C | B read mirror, tangle, radix-R1 | C write
//                       stage layout | D read, twiddle, radix-R2, store
// which (R1, R2) run warp-synchronously: LN = R1 lanes per row -- a half-warp for the 16 x R2 lengths, a
// quarter-warp (two mirrored row pairs per warp) for 8 x 8
template <int R1, int R2> struct RowWarpOk { static constexpr bool value = (R1 == 16 && R2 <= 16) || (R1 == 8 && R2 == 8); };

template <class T, int R1, int R2> struct RowWarp {
  static_assert(RowWarpOk<R1, R2>::value, "R1 lanes per row, R2 <= R1");
  using G = Cfg<T, R1, R2>;
  using Ctx = ThreadCtx<T, 16>;
  static constexpr int LN = R1;                      // lanes per row
  static constexpr int PP = 32 / (2 * LN);           // mirrored row pairs per warp
  static constexpr int N = R1 * R2, C = G::C, THREADS = LN * C;
  static constexpr int P1 = R1 + 1;                  // padded pitch of the stage exchange
  static constexpr int COLBUF = R2 * P1;             // complex elements per row buffer (>= N)
  static constexpr int SMEM = C * COLBUF * 2 * (int)sizeof(T) + THREADS * 8;
  static_assert(SMEM <= G::ROW_SMEM, "fits the row-pass allocation");

  // MODE 0: the role is read from p.load_mode (emulator); 1: forward pass only; 2: pair product only --
  // the kernels are compiled per role so that neither carries the other's registers and code
  template <int MODE> static XCB_HD bool is_pro(const PassParams<T>& p) {
    return MODE == 0 ? p.load_mode == RL_PROLOGUE : MODE == 2;
  }
  // the pair-product kernels (MODE 2) are bound by instruction issue: packed f32x2 complex arithmetic there
  // (xcb_types.h); the forward pass measured slower with it (1.34 -> 1.40 ms) and keeps the scalar form
  template <int MODE> struct Pk { static constexpr bool value = MODE == 2; };

  struct Slot {
    int32_t batch, row;   // (32 bits: a launch has fewer than 2^31 jobs and rows; 64-bit products at the uses)
    cx<T>* buf;        // this half's row buffer
    const cx<T>* mbuf; // buffer holding the mirrored row
    cx<T> wrow;        // W_L^klow(row): half-step twiddle of the row's leading digits
    cx<T> w_inner;     // W_N^j: inner twiddle of the second radix stage
    const cx<T>* tw2;  // W_2N^e table (global, or the grouped kernel's shared-memory copy)
    T fs;              // pair product: scale of the job when known up front, < 0: derive from the norms
    int j;
    bool active;       // this half-warp has a row
    bool row0;         // row 0 mirrors itself with e -> (N - e) mod N
  };

  static XCB_HD Slot decode(const PassParams<T>& p, const Ctx& ctx, unsigned char* smem_raw) {
    Slot s;
    // grids stay below 2^31 CTAs: 32-bit division
    const uint32_t bid = (uint32_t)ctx.bid;
    uint32_t tile, batch;
    if (p.tile_major) { tile = bid / (uint32_t)p.n_batch; batch = bid - tile * (uint32_t)p.n_batch; }
    else { batch = bid / (uint32_t)p.tiles_per_batch; tile = bid - batch * (uint32_t)p.tiles_per_batch; }
    s.batch = (int32_t)batch;
    const int col = ctx.tid / LN, side = col & 1;
    s.j = ctx.tid % LN;
    int64_t ra, rb;
    const bool valid = RowPass<T, R1, R2>::pair_rows(p, (int64_t)tile * (C / 2) + (col >> 1), ra, rb);
    const bool self = ra == rb;
    s.active = valid && !(side && self);
    s.row = (int32_t)(side ? rb : ra);
    s.row0 = ra == 0;
    cx<T>* base = reinterpret_cast<cx<T>*>(smem_raw);
    s.buf = base + col * COLBUF;
    s.mbuf = base + (self ? col : (col ^ 1)) * COLBUF;
    s.wrow.x = T(1); s.wrow.y = T(0);
    if (s.active) s.wrow = row_twiddle(p, s.row);
    s.w_inner = ld_ro(p.tw_inner + s.j);
    s.tw2 = p.tw2_inner;
    s.fs = T(-1);
    return s;
  }
  static XCB_HD double* red_ptr(unsigned char* smem_raw) {
    return reinterpret_cast<double*>(smem_raw + (size_t)C * COLBUF * sizeof(cx<T>));
  }
  static XCB_HD int mirror(const Slot& s, int e) { return s.row0 ? (e == 0 ? 0 : N - e) : N - 1 - e; }
  static XCB_HD cx<T> row_twiddle(const PassParams<T>& p, int64_t row) {
    const uint32_t kl = RowPass<T, R1, R2>::klow_of_row(p, row);
    return cmul(ld_ro(p.tw2_hi + (kl >> p.tw_shift)), ld_ro(p.tw2_lo + (kl & p.tw_mask)));
  }
  // offsets of term w (block / stacked window) of a job's two spectra, and its scale
  static XCB_HD void term(const PassParams<T>& p, const Job& jb, int w, bool per_term, int64_t& oi, int64_t& oj, T& sw) {
    oi = jb.spec_i + (p.use_blk_off ? p.blk_off_i[w] : w * p.stack_spec_stride);
    oj = jb.spec_j + (p.use_blk_off ? p.blk_off_j[w] : w * p.stack_spec_stride);
    sw = T(1);
    if (per_term) sw = T(p.rnorm[jb.sig_i + w * p.stack_sig_stride] * p.rnorm[jb.sig_j + w * p.stack_sig_stride]);
  }

  // v[r] += sw * a[r R2] * conj(b[r R2]), r < R1: this lane's share of one term
  // (block / stacked window) of the cross-spectrum; a, b point at the lane's first
  // element of the two rows (global memory, or rows staged in shared memory)
  template <bool SCALED, bool PK>
  static XCB_HD void accumulate_t(cx<T>* v, const cx<T>* a_row, const cx<T>* b_row, T sw) {
    constexpr int H = R1 / 2;                       // two batches of loads bound the registers
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      cx<T> a[H], b[H];
#pragma unroll
      for (int r = 0; r < H; ++r) { a[r] = a_row[(h * H + r) * R2]; b[r] = b_row[(h * H + r) * R2]; }
#pragma unroll
      for (int r = 0; r < H; ++r)
        v[h * H + r] = xmulc_acc<PK>(SCALED ? xscale<PK>(a[r], sw) : a[r], b[r], v[h * H + r]);   // += a conj(b)
    }
  }
  // first unscaled term: v[r] = a[r R2] conj(b[r R2]).  The a's are loaded straight INTO the accumulators (all
  // R1 loads in flight without staging registers), the b's follow in two batches.  (ncu, grouped kernel: with
  // zeroed accumulators and a multiply-add per term ptxas, at its register limit, reused one register pair for
  // every element -- LDS, LDS, FFMA2, FFMA2 sixteen times in series, 20 % of the kernel's stall samples.)
  template <bool PK>
  static XCB_HD void accumulate_first(cx<T>* v, const cx<T>* a_row, const cx<T>* b_row) {
#pragma unroll
    for (int r = 0; r < R1; ++r) v[r] = a_row[r * R2];
    constexpr int H = R1 / 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      cx<T> b[H];
#pragma unroll
      for (int r = 0; r < H; ++r) b[r] = b_row[(h * H + r) * R2];
#pragma unroll
      for (int r = 0; r < H; ++r) v[h * H + r] = xmulc<PK>(v[h * H + r], b[r]);
    }
  }
  // further unscaled terms in batches of Q elements (Q a and Q b values staged)
  template <bool PK, int Q>
  static XCB_HD void accumulate_next(cx<T>* v, const cx<T>* a_row, const cx<T>* b_row) {
#pragma unroll
    for (int h = 0; h < R1 / Q; ++h) {
      cx<T> a[Q], b[Q];
#pragma unroll
      for (int r = 0; r < Q; ++r) { a[r] = a_row[(h * Q + r) * R2]; b[r] = b_row[(h * Q + r) * R2]; }
#pragma unroll
      for (int r = 0; r < Q; ++r) v[h * Q + r] = xmulc_acc<PK>(a[r], b[r], v[h * Q + r]);
    }
  }
  // two unscaled terms at once (block mode with two blocks): v[r] = a0 conj(b0) + a1 conj(b1), Q elements per batch
  template <bool PK, int Q>
  static XCB_HD void accumulate_two(cx<T>* v, const cx<T>* a0, const cx<T>* b0, const cx<T>* a1, const cx<T>* b1) {
#pragma unroll
    for (int r = 0; r < R1; ++r) v[r] = a0[r * R2];
#pragma unroll
    for (int h = 0; h < R1 / Q; ++h) {
      cx<T> x[Q], y[Q], z[Q];
#pragma unroll
      for (int r = 0; r < Q; ++r) {
        x[r] = b0[(h * Q + r) * R2]; y[r] = a1[(h * Q + r) * R2]; z[r] = b1[(h * Q + r) * R2];
      }
#pragma unroll
      for (int r = 0; r < Q; ++r) v[h * Q + r] = xmulc_acc<PK>(y[r], z[r], xmulc<PK>(v[h * Q + r], x[r]));
    }
  }
  // `scaled`: per-term norms (stacked windows); else the single scale is applied after the sum
  template <bool PK = false>
  static XCB_HD void accumulate(cx<T>* v, const cx<T>* a_row, const cx<T>* b_row, T sw, bool scaled = true) {
    if (scaled) accumulate_t<true, PK>(v, a_row, b_row, sw);
    else accumulate_t<false, PK>(v, a_row, b_row, sw);
  }

  // slot 0 of a spectrum holds (X[0], X[M]), both real: C[0] and C[M] are plain products
  static XCB_HD cx<T> packed_dc(const PassParams<T>& p, const Job& jb, bool per_term) {
    T c0 = T(0), cm = T(0);
    for (int w = 0; w < p.n_stack; ++w) {
      int64_t oi, oj;
      T sw;
      term(p, jb, w, per_term, oi, oj, sw);
      const cx<T> a = p.spec[oi], b = p.spec[oj];
      c0 += a.x * sw * b.x;
      cm += a.y * sw * b.y;
    }
    cx<T> r; r.x = c0; r.y = cm;
    return r;
  }

  template <int MODE = 0>
  static XCB_HD void phase_a(const PassParams<T>& p, Ctx& ctx, const Slot& s, unsigned char* smem_raw) {
    cx<T>* v = ctx.v;
    ctx.ss = 0.0;
    if (s.active && s.j < R2) {
      if (is_pro<MODE>(p)) {
        const Job jb = p.jobs[s.batch];
        const bool per_term = p.normalize && p.stack_sig_stride != 0;
        const cx<T>* z = p.spec + (int64_t)s.row * N + s.j;
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r].x = v[r].y = T(0);
        for (int w = 0; w < p.n_stack; ++w) {
          int64_t oi, oj;
          T sw;
          term(p, jb, w, per_term, oi, oj, sw);
          accumulate<Pk<MODE>::value>(v, z + oi, z + oj, sw, per_term);
        }
        if (s.row0 && s.j == 0) v[0] = packed_dc(p, jb, per_term);
#pragma unroll
        for (int r = 0; r < R1; ++r) s.buf[s.j + r * R2] = v[r];       // publish C (linear index)
      } else {
        if (p.load_mode == RL_REAL) {
          const T* x = reinterpret_cast<const T*>(p.in) + (int64_t)s.batch * p.in_stride;
#pragma unroll
          for (int r = 0; r < R1; ++r) v[r] = load_real_packed(p, x, (int64_t)s.row * N + s.j + r * R2, ctx.ss);
        } else {
          const cx<T>* src = reinterpret_cast<const cx<T>*>(p.in) + (int64_t)s.batch * p.in_stride + (int64_t)s.row * N + s.j;
#pragma unroll
          for (int r = 0; r < R1; ++r) v[r] = src[r * R2];
        }
        fft_reg<R1>(v);
#pragma unroll
        for (int r = 0; r < R1; ++r) s.buf[s.j * P1 + r] = v[r];       // stage layout
      }
    }
    if (MODE != 2 && p.sumsq_partial && p.load_mode == RL_REAL) red_ptr(smem_raw)[ctx.tid] = ctx.ss;
  }

  // stage 2: read the stage layout, inner twiddles, radix-R2 -> v[r] = element j + R1 r
  template <bool PK = false> static XCB_HD void stage2(const PassParams<T>& p, const Slot& s, cx<T>* v) {
#pragma unroll
    for (int r = 0; r < R2; ++r) v[r] = s.buf[r * P1 + s.j];
    twiddle_powers<R2, PK>(s.w_inner, v);
    fft_reg<R2, PK>(v);
  }

  // W_2N^(r R2) = exp(-i pi r / R1) (table of multiples of pi/16): the half-step twiddle of element e = j + r R2 is W_2N^j times this constant,
  // so the pair product needs one table value per lane instead of one per element (the grouped kernel is bound
  // by shared-memory wavefronts: ncu L1/TEX throughput 77 %)
  static XCB_HD cx<T> half_step(int r) {
    constexpr double c[16] = {1, 0.98078528040323043, 0.92387953251128674, 0.83146961230254524, 0.70710678118654757,
                              0.55557023301960229, 0.38268343236508984, 0.19509032201612833, 0.0,
                              -0.19509032201612819, -0.38268343236508973, -0.55557023301960196, -0.70710678118654746,
                              -0.83146961230254535, -0.92387953251128674, -0.98078528040323043};
    constexpr double sn[16] = {0, 0.19509032201612825, 0.38268343236508978, 0.55557023301960218, 0.70710678118654746,
                               0.83146961230254524, 0.92387953251128674, 0.98078528040323043, 1, 0.98078528040323043,
                               0.92387953251128674, 0.83146961230254546, 0.70710678118654757, 0.55557023301960218,
                               0.38268343236508989, 0.19509032201612861};
    cx<T> w; w.x = (T)c[r * (16 / R1)]; w.y = (T)-sn[r * (16 / R1)];
    return w;
  }
  template <bool PK> static XCB_HD cx<T> row_w(const cx<T>& wb, int r) {
    if (r == 0) return wb;
    if (r == R1 / 2) return cmi(wb);                 // exp(-i pi/2) = -i
    return xmul<PK>(wb, half_step(r));
  }
  // E = C[k] + conj(C[M-k]), D = C[k] - conj(C[M-k]), O = conj(W) D, V = (E + iO)/2; returns fs conj(V)
  template <bool PK> static XCB_HD cx<T> tangle_one(const cx<T>& c, const cx<T>& cm, const cx<T>& W, T fs) {
    const cx<T> cmc = cconj(cm);
    const cx<T> E = xadd<PK>(c, cmc), D = xsub<PK>(c, cmc);
    const cx<T> O = xmulc<PK>(D, W);
    cx<T> o = xscale<PK>(xadd<PK>(E, cpi(O)), fs);   // fs (E + iO) ...
    o.y = -o.y;                                       // ... conjugated
    return o;
  }
  static XCB_HD T job_scale(const PassParams<T>& p, const Slot& s) {
    T fs = s.fs;
    if (fs < T(0)) {
      const Job jb = p.jobs[s.batch];
      const bool per_term = p.normalize && p.stack_sig_stride != 0;
      fs = T(0.5);
      if (p.normalize && !per_term) fs = T(0.5 * p.rnorm[jb.sig_i] * p.rnorm[jb.sig_j]);
    }
    return fs;
  }

  template <int MODE = 0>
  static XCB_HD void phase_b(const PassParams<T>& p, Ctx& ctx, const Slot& s) {
    cx<T>* v = ctx.v;
    if (!s.active) return;
    if (is_pro<MODE>(p)) {
      if (s.j >= R2) return;
      const T fs = job_scale(p, s);
      constexpr bool PK = Pk<MODE>::value;
      const cx<T> wb = xmul<PK>(s.wrow, s.tw2[s.j]);
#pragma unroll
      for (int r = 0; r < R1; ++r) {
        const int e = s.j + r * R2;
        const cx<T> cm = s.mbuf[mirror(s, e)];
        cx<T> o = tangle_one<PK>(v[r], cm, row_w<PK>(wb, r), fs);
        if (r == 0 && s.row0 && s.j == 0) {          // V[0] = (C[0] + C[M])/2 + i (C[0] - C[M])/2
          o.x = fs * (v[0].x + v[0].y); o.y = -fs * (v[0].x - v[0].y);
        }
        v[r] = o;
      }
      fft_reg<R1, Pk<MODE>::value>(v);
    } else {
      stage2(p, s, v);
    }
  }

#if XCB_PK_DEVICE
  // Pair product, rows other than the self-mirrored ones: the mirrored element of (row ra, e = j + r R2) is
  // (row rb, N - 1 - e) = lane R2 - 1 - j of the OTHER half-warp, register R1 - 1 - r -- a warp shuffle
  // instead of a round trip through the row buffers (16 stores + 16 loads per lane).  Called by ALL lanes of a
  // warp that has a job (lanes j >= R2 only serve the shuffles).
  static __device__ __forceinline__ void phase_b_shfl(const PassParams<T>& p, Ctx& ctx, const Slot& s) {
    cx<T>* v = ctx.v;
    constexpr bool PK = true;
    const bool on = s.j < R2;
    const int lane = ctx.tid & 31;                    // partner: same row pair, other side, lane R2 - 1 - j
    const int src = (lane & ~(2 * LN - 1)) + ((lane & LN) ^ LN) + (R2 - 1 - s.j < 0 ? 0 : R2 - 1 - s.j);
    const T fs = job_scale(p, s);
    cx<T> wb;
    wb.x = T(1); wb.y = T(0);
    if (on) wb = xmul<PK>(s.wrow, s.tw2[s.j]);
#pragma unroll
    for (int r = 0; r < R1 / 2; ++r) {
      const int q = R1 - 1 - r;
      cx<T> ma, mb;                                  // mirrors of elements r and q: the partner's registers q and r
      ma.x = __shfl_sync(0xffffffffu, v[q].x, src); ma.y = __shfl_sync(0xffffffffu, v[q].y, src);
      mb.x = __shfl_sync(0xffffffffu, v[r].x, src); mb.y = __shfl_sync(0xffffffffu, v[r].y, src);
      v[r] = tangle_one<PK>(v[r], ma, row_w<PK>(wb, r), fs);
      v[q] = tangle_one<PK>(v[q], mb, row_w<PK>(wb, q), fs);
    }
    if (on) fft_reg<R1, PK>(v);
  }
#endif

  template <int MODE = 0>
  static XCB_HD void phase_c(const PassParams<T>& p, Ctx& ctx, const Slot& s) {
    const cx<T>* v = ctx.v;
    if (!s.active) return;
    if (is_pro<MODE>(p)) {
      if (s.j >= R2) return;
#pragma unroll
      for (int r = 0; r < R1; ++r) s.buf[s.j * P1 + r] = v[r];         // stage layout
    } else if (p.store_mode == RS_HALF) {
#pragma unroll
      for (int r = 0; r < R2; ++r) s.buf[s.j + r * R1] = v[r];         // publish Z (linear index)
    }
  }

  template <int MODE = 0>
  static XCB_HD void phase_d(const PassParams<T>& p, Ctx& ctx, const Slot& s, unsigned char* smem_raw) {
    cx<T>* v = ctx.v;
    if (s.active) {
      if (is_pro<MODE>(p)) {
        stage2<Pk<MODE>::value>(p, s, v);
        if (p.store_mode == RS_EPI) {
          T* outp = p.ccf + p.jobs[s.batch].out_off;
#pragma unroll
          for (int r = 0; r < R2; ++r)
            epilogue_store(p, outp, (uint32_t)s.row * (uint32_t)N + (uint32_t)(s.j + r * R1), v[r].x, -v[r].y);
        } else {
          cx<T>* dst = reinterpret_cast<cx<T>*>(p.out) + (int64_t)s.batch * p.out_stride + (int64_t)s.row * N + s.j;
#pragma unroll
          for (int r = 0; r < R2; ++r) dst[r * R1] = cconj(v[r]);
        }
      } else if (p.store_mode == RS_EPI) {
        // plain complex transform, last inverse pass of a single-pass plan: interleaved output
        T* outp = p.ccf + p.jobs[s.batch].out_off;
#pragma unroll
        for (int r = 0; r < R2; ++r)
          epilogue_store(p, outp, (uint32_t)s.row * (uint32_t)N + (uint32_t)(s.j + r * R1), v[r].x, -v[r].y);
      } else if (p.store_mode != RS_HALF) {
        // plain complex transform (xcb_fft_*): no real <-> packed conversion
        cx<T>* dst = reinterpret_cast<cx<T>*>(p.out) + (int64_t)s.batch * p.out_stride + (int64_t)s.row * N + s.j;
#pragma unroll
        for (int r = 0; r < R2; ++r) dst[r * R1] = p.store_mode == RS_CONJ ? cconj(v[r]) : v[r];
      } else {
        // X[k] = (S - i W D)/2 with S = Z[k] + conj(Z[M-k]), D = Z[k] - conj(Z[M-k]), W = W_L^k
        const cx<T> wrow = s.wrow;
        cx<T>* dst = reinterpret_cast<cx<T>*>(p.out) + (int64_t)s.batch * p.out_stride + (int64_t)s.row * N + s.j;
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          const int k = s.j + r * R1;
          const cx<T> zm = s.mbuf[mirror(s, k)];
          const cx<T> W = cmul(wrow, s.tw2[k]);
          cx<T> S, D;
          S.x = v[r].x + zm.x; S.y = v[r].y - zm.y;
          D.x = v[r].x - zm.x; D.y = v[r].y + zm.y;
          const cx<T> WD = cmul(W, D);
          cx<T> X;
          X.x = T(0.5) * (S.x + WD.y); X.y = T(0.5) * (S.y - WD.x);
          if (r == 0 && s.row0 && s.j == 0) { X.x = v[0].x + v[0].y; X.y = v[0].x - v[0].y; }   // (X[0], X[M])
          dst[r * R1] = X;
        }
      }
    }
    if (MODE != 2 && p.sumsq_partial && p.load_mode == RL_REAL)
      block_sumsq<THREADS>(p.sumsq_partial + ctx.bid, red_ptr(smem_raw), ctx.tid);
  }
};


}  // namespace xcb
