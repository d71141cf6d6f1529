// shared declarations: tuning knobs, the Dev argument block, interpolation helpers, the DMMA tile instruction
// (part of bmpc_kernels.cuh: include that header, not this file)
#pragma once

namespace bmpc {


constexpr double WEAK_EPS = 1e-6;   // [UPSTREAM] numeric_traits::weakEpsilon
// per-tick counters (ints) read back by the host after the tick: total / maximum line-search trials, failed instances, OR of all status bits
constexpr int CNT_TRIALS = 0, CNT_MAXTRIALS = 1, CNT_FAIL = 2, CNT_STATUS = 3, CNT_N = 8;
#ifndef LQ_PAIR_BLOCKS
#define LQ_PAIR_BLOCKS 2
#endif
#ifndef LS2_BLOCKS
#define LS2_BLOCKS 2   // CTAs (128 threads) per SM of k_linesearch
#endif
#ifndef PROJ_BLOCKS
#define PROJ_BLOCKS 4
#endif
#ifndef EXP_BLOCKS
#define EXP_BLOCKS 5    // CTAs (4 warps) per SM of k_policy_expand: 92 registers without spills, 20 warps keep more loads in flight (1.31 -> 1.12 ms)
#endif
#ifndef RIC_BLOCKS
#define RIC_BLOCKS 3
#endif
#ifndef RIC_WPC
#define RIC_WPC 4   // independent instances (warps) per CTA of k_riccati_warp
#endif

#ifndef FWD_WPC
#define FWD_WPC 7      // independent instances (warps) per CTA of k_forward
#endif
#ifndef FWD_BLOCKS
#define FWD_BLOCKS 2
#endif

#ifndef PF_AHEAD
#define PF_AHEAD 296   // warp-per-stage kernels first ask L2 for the whole input record of the stage PF_AHEAD warps ahead (same wave): the dependent rounds of cold DRAM loads at the start of a warp then hit L2
#endif

template <int NJ> struct RDims;
template <int NJ> struct SDims;

// L2 prefetch of a contiguous record (16-byte aligned address, size a multiple of 16 bytes), one lane issues it.  The warp-per-stage kernels are
// latency bound by dependent rounds of cold DRAM loads at the start of every warp; CTAs are dispatched in blockIdx order, so the warp that starts
// one wave later finds its records in the 126 MB L2 instead.
__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

struct Dev {
  int B, NS, ME, TP, npts;   // NS: node slots per instance = nominal grid + 1 + max_event_nodes
  double dt_nom, horizon;
  const double* t0; const double* x0;
  const double* tgt_t; const double* tgt_x;
  const int* n_ev; const double* ev_t; const int* ev_mode;
  int* n_nodes; double* node_t; int* node_ev; double* st_t; double* st_dt; int* st_mode;
  double* xref; double* zref;
  const int* p_n; const double* p_t; const double* p_x; const double* p_u;   // previous primal solution (warm start)
  const int* p_ev; const double* p_uff; const double* p_K;                   // rest of the previous policy (kept by an instance whose solve fails)
  double* s_x; double* s_u; double* s_uff; double* s_K;                      // new primal solution / linearisation point
  double* lq; double* proj; double* stage; double* ric; const double* jc;
  double* dx; double* du;
  double* perf; double* alpha; double* norms; int* status; int* counters;
};

// ------------------------------------------------------------------------------------------------ helpers
// FP64 tensor-core tile: D(8x8) += A(8x4) B(4x8); lane (g, q) = (lane >> 2, lane & 3) supplies A[g][q], B[q][g] and owns D[g][2q], D[g][2q+1]  (SASS: DMMA)
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ int lower_bound_d(const double* a, int n, double t) {  // first index with a[i] >= t
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < t) lo = mid + 1; else hi = mid; }
  return lo;
}
// [UPSTREAM] LinearInterpolation::timeSegment
__device__ __forceinline__ void time_segment(const double* ta, int n, double t, int& index, double& alpha) {
  int idx = lower_bound_d(ta, n, t);
  int iv = (idx == 0 && n > 0 && t == ta[0]) ? 0 : idx - 1;
  const int last = n - 1;
  if (iv >= 0) {
    if (iv < last) {
      const double len = ta[iv + 1] - ta[iv], till = ta[iv + 1] - t;
      if (len > 2.0 * 2.220446049250313e-16) { index = iv; alpha = till / len; }
      else { index = iv; alpha = (till < 0.5 * len) ? 0.0 : 1.0; }
    } else { index = max(last - 1, 0); alpha = 0.0; }
  } else { index = 0; alpha = 1.0; }
}
__device__ __forceinline__ void interp_vec(const double* ta, const double* data, int n, int dim, double t, double* out) {
  if (n <= 1) { for (int i = 0; i < dim; ++i) out[i] = data[i]; return; }
  int idx; double al; time_segment(ta, n, t, idx, al);
  const double* a = data + (size_t)idx * dim; const double* b = a + dim;
  for (int i = 0; i < dim; ++i) out[i] = al * a[i] + (1.0 - al) * b[i];
}

}  // namespace bmpc
