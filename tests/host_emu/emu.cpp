// Host thread-emulator for the FFT pass bodies (TEST HARNESS ONLY).
//
// The build container has no GPU, so the device code in xcb_passes.h is written
// as barrier-separated phases over an explicit per-thread context.  This file
// runs those same phases thread-by-thread on the CPU so that the index algebra,
// twiddles, pair product and lag epilogue can be checked against the oracle
// before any GPU time is spent.  It is compiled only by tests/ (g++), is not
// part of libxcorr_b200.so, and is never used by the product path.
#include <stdio.h>
#include <string.h>
#include <vector>
#include "../../xcorr_b200/csrc/xcb_plan.h"

using namespace xcb;

template <class T, int R1, int R2> static void emu_col(const PassParams<T>& p) {
  using P = ColPass<T, R1, R2>;
  using G = Cfg<T, R1, R2>;
  std::vector<unsigned char> smem(G::COL_SMEM + 64);
  std::vector<typename P::Ctx> ctx(G::THREADS);
  int64_t nblocks = (int64_t)p.n_batch * p.tiles_per_batch;
  for (int64_t b = 0; b < nblocks; ++b) {
    for (int t = 0; t < G::THREADS; ++t) { ctx[t].tid = t; ctx[t].bid = b; P::phase_a(p, ctx[t], smem.data()); }
    for (int t = 0; t < G::THREADS; ++t) P::phase_b(p, ctx[t], smem.data());
  }
}

template <class T, int R1, int R2> static void emu_row(const PassParams<T>& p) {
  using G = Cfg<T, R1, R2>;
  std::vector<unsigned char> smem(G::ROW_SMEM + 64);
  int64_t nblocks = (int64_t)p.n_batch * p.tiles_per_batch;
  if constexpr (RowWarpOk<R1, R2>::value) {          // warp-synchronous variant: phases are separated by __syncwarp on the device
    using P = RowWarp<T, R1, R2>;
    std::vector<typename P::Ctx> ctx(G::THREADS);
    std::vector<typename P::Slot> slot(G::THREADS);
    for (int64_t b = 0; b < nblocks; ++b) {
      for (int t = 0; t < G::THREADS; ++t) {
        ctx[t].tid = t; ctx[t].bid = b;
        slot[t] = P::decode(p, ctx[t], smem.data());
        P::phase_a(p, ctx[t], slot[t], smem.data());
      }
      for (int t = 0; t < G::THREADS; ++t) P::phase_b(p, ctx[t], slot[t]);
      for (int t = 0; t < G::THREADS; ++t) P::phase_c(p, ctx[t], slot[t]);
      for (int t = 0; t < G::THREADS; ++t) P::phase_d(p, ctx[t], slot[t], smem.data());
    }
  } else {
    using P = RowPass<T, R1, R2>;
    std::vector<typename P::Ctx> ctx(G::THREADS);
    for (int64_t b = 0; b < nblocks; ++b) {
      for (int t = 0; t < G::THREADS; ++t) { ctx[t].tid = t; ctx[t].bid = b; P::phase_info(p, ctx[t], smem.data()); }
      for (int t = 0; t < G::THREADS; ++t) P::phase_load(p, ctx[t], smem.data());
      for (int t = 0; t < G::THREADS; ++t) P::phase_s1_read(p, ctx[t], smem.data());
      for (int t = 0; t < G::THREADS; ++t) P::phase_s1_write(p, ctx[t], smem.data());
      for (int t = 0; t < G::THREADS; ++t) P::phase_s2(p, ctx[t], smem.data());
      for (int t = 0; t < G::THREADS; ++t) P::phase_store(p, ctx[t], smem.data());
    }
  }
}

template <class T> static bool run_pass(const PassParams<T>& p, bool row) {
  switch (p.N) {
#define X(n, a, b) case n: if (row) emu_row<T, a, b>(p); else emu_col<T, a, b>(p); return true;
    XCB_MENU(X)
#undef X
  }
  return false;
}

template <class T>
static int emu_cc_t(int64_t n, const int* dims, int nd, const T* x, int64_t x_stride, int n_signals,
                    const int* pairs, int n_pairs, int64_t i0, int64_t i1, int normalize, int unbiased,
                    T* out, double* sumsq_out, T* spec_out, int n_stack, int n_blocks, int64_t block_len) {
  PlanGeom g;
  int d[kMaxDims], k = nd;
  const int64_t K = i1 - i0;
  if (n_blocks < 0) {          // automatic block planning, as xcb_plan_create does
    n_blocks = 1;
    int bd[kMaxDims], bnd = 0, nb = 0;
    int64_t bl = 0;
    if (nd == 0 && plan_choose_blocks(n, K, n_signals, n_pairs, bd, &bnd, &nb, &bl)) {
      n_blocks = nb; block_len = bl; k = bnd;
      for (int a = 0; a < bnd; ++a) d[a] = bd[a];
    }
  }
  if (n_blocks <= 1) {
    if (nd == 0) { if (!plan_choose_dims(n, d, &k)) return -1; }
    else for (int a = 0; a < nd; ++a) d[a] = dims[a];
    if (!plan_geom_init(g, n, d, k)) return -2;
  } else {
    if (nd > 0) for (int a = 0; a < nd; ++a) d[a] = dims[a];
    if (!plan_geom_init(g, n, d, k, n_blocks, block_len, i0 - (n - 1), K)) return -2;
  }
  HostTables<T> ht;
  build_tables(g, ht);
  DevTables<T> tb;
  tb.hi = ht.hi.data(); tb.lo = ht.lo.data();
  for (int a = 0; a < g.nd; ++a) tb.inner[a] = ht.inner[a].data();
  tb.hi2 = ht.hi2.data(); tb.lo2 = ht.lo2.data(); tb.inner2 = ht.inner2.data();

  auto runner = [&](const PassParams<T>& p, bool row, int) { return run_pass(p, row); };
  const int per = spectra_per_signal(g);
  const int64_t t0 = pass0_tiles(g, tile_cols<T>());
  std::vector<cx<T>> spec((size_t)n_signals * per * g.M);
  std::vector<double> partial((size_t)g.n_blocks * n_signals * t0, 0.0);
  if (!run_forward<T>(g, tb, x, x_stride, n_signals, spec.data(), partial.data(), runner)) return -3;
  std::vector<double> sumsq(n_signals, 0.0), rnorm(n_signals);
  for (int b = 0; b < g.n_blocks; ++b)
    for (int s = 0; s < n_signals; ++s)
      for (int64_t t = 0; t < t0; ++t) sumsq[s] += partial[((size_t)b * n_signals + s) * t0 + t];
  for (int s = 0; s < n_signals; ++s) rnorm[s] = 1.0 / sqrt(sumsq[s]);
  if (sumsq_out) memcpy(sumsq_out, sumsq.data(), sizeof(double) * n_signals);
  if (spec_out && per == 1) memcpy(spec_out, spec.data(), sizeof(cx<T>) * spec.size());

  std::vector<Job> jobs(n_pairs);
  // signals are [receiver][window] when stacking (n_stack windows per receiver)
  for (int q = 0; q < n_pairs; ++q)
    jobs[q] = make_job(g, (int64_t)pairs[2 * q] * n_stack, (int64_t)pairs[2 * q + 1] * n_stack, (int64_t)q * K);
  std::vector<cx<T>> work((size_t)n_pairs * g.M);
  CorrFlags f; f.i0 = i0; f.i1 = i1; f.normalize = normalize; f.unbiased = unbiased; f.accumulate = 0;
  f.n_stack = n_stack; f.stack_sig_stride = 1; f.stack_spec_stride = g.M;
  if (!run_inverse<T>(g, tb, spec.data(), rnorm.data(), jobs.data(), n_pairs, work.data(), out, f, runner)) return -4;
  return g.n_blocks;
}

extern "C" {

int emu_choose_dims(int64_t n, int* dims, int* nd) { return plan_choose_dims(n, dims, nd) ? 0 : -1; }

// block planner as xcb_plan_create runs it; returns the number of blocks (1 = classic transform)
int emu_choose_blocks(int64_t n, int64_t n_lags, int n_recv, int n_pairs, int* dims, int* nd, int64_t* block_len) {
  int nb = 1;
  if (!plan_choose_blocks(n, n_lags, n_recv, n_pairs, dims, nd, &nb, block_len)) return 1;
  return nb;
}

// Job grouping of the staged pair product for `n_pairs` pairs of one window (classic plan over M points).
// Writes per job its group id; returns the number of groups, or a negative code when a group is inconsistent.
int emu_pair_groups(const int* pairs, int n_pairs, int max_jobs, int max_rows, int* group_of_job, int* rows_total) {
  PlanGeom g;
  int d[1] = {64};
  plan_geom_init(g, 50, d, 1);
  std::vector<Job> jobs(n_pairs);
  for (int q = 0; q < n_pairs; ++q) jobs[q] = make_job(g, pairs[2 * q], pairs[2 * q + 1], (int64_t)q * 99);
  std::vector<PairGroup> groups;
  build_pair_groups(jobs.data(), n_pairs, max_jobs, max_rows, groups);
  for (int q = 0; q < n_pairs; ++q) group_of_job[q] = -1;
  *rows_total = 0;
  for (size_t k = 0; k < groups.size(); ++k) {
    const PairGroup& gr = groups[k];
    if (gr.n_jobs < 1 || gr.n_jobs > max_jobs || gr.n_i + gr.n_j > max_rows) return -1;
    *rows_total += gr.n_i + gr.n_j;
    for (int w = 0; w < gr.n_jobs; ++w) {
      const int q = gr.job[w];
      if (q < 0 || q >= n_pairs || group_of_job[q] != -1) return -2;
      if (gr.spec_i[gr.ji[w]] != jobs[q].spec_i || gr.spec_j[gr.jj[w]] != jobs[q].spec_j) return -3;
      group_of_job[q] = (int)k;
    }
  }
  return (int)groups.size();
}

// dtype: 0 = f32, 1 = f64.  x: [n_signals][x_stride]; pairs: [n_pairs][2]; out: [n_pairs][i1-i0]
int emu_cc(int dtype, int64_t n, const int* dims, int nd, const void* x, int64_t x_stride, int n_signals,
           const int* pairs, int n_pairs, int64_t i0, int64_t i1, int normalize, int unbiased, void* out,
           double* sumsq_out, void* spec_out, int n_stack, int n_blocks, int64_t block_len) {
  if (n_stack < 1) n_stack = 1;
  if (dtype == 0)
    return emu_cc_t<float>(n, dims, nd, (const float*)x, x_stride, n_signals, pairs, n_pairs, i0, i1, normalize,
                           unbiased, (float*)out, sumsq_out, (float*)spec_out, n_stack, n_blocks, block_len);
  return emu_cc_t<double>(n, dims, nd, (const double*)x, x_stride, n_signals, pairs, n_pairs, i0, i1, normalize,
                          unbiased, (double*)out, sumsq_out, (double*)spec_out, n_stack, n_blocks, block_len);
}
}
