// K1.5: constraint projection (Householder QR) + change of input variables on the FP64 tensor cores; projected stage record layout
// (part of bmpc_kernels.cuh: include that header, not this file)
#pragma once

namespace bmpc {

// ------------------------------------------------------------------------------------------------ K1.5: constraint projection + change of input variables
// One warp per (instance, stage).
//  (1) Dv (r x NJ, full row rank after the per-foot compression) -> Householder QR of Dv^T = Q [R; 0]:
//      Dv^+ = Q1 R^-T (Moore-Penrose), null(Dv) = span(Q2):  Pxj = -Dv^+ Cv, Pej = -Dv^+ ev, N = Q2
//      (replaces LinearAlgebra::luConstraintProjection [UPSTREAM]).
//  (2) changeOfInputVariables [UPSTREAM] with du = Pe + Px dx + Pu dut, exploiting the block structure
//      (forces of closed contacts stay free, forces of open contacts are fixed to -F, joint velocities = Pej + Pxj dx + N dut_null):
//      writes the projected stage record (SDims) that the sequential Riccati kernel consumes.
// Projected stage record (k_project -> k_riccati, k_policy_expand), padded to NXP = 24 states / MP = 16 reduced inputs:
//   [AB | bt | qt | rt | meta]  one contiguous block that the Riccati kernel stages with a single TMA bulk copy:
//       AB = [At | Bt] (24 x 42, row major; columns 0..23 = At incl. identity, 24..39 = Bt, 40..41 pad).  The leading dimension 42 = 2 mod 4
//       makes the k-permuted transposed DMMA fragment loads of k_riccati bank-conflict free.
//   QF  = Qt (24 x 24, full, diagonal included) in DMMA accumulator-fragment order: [tile 3x3][lane][2]
//   PRF = [Pt | Rt] (16 x 40) in accumulator-fragment order: [tile 2x5][lane][2]   (Rt padded with the identity beyond m)
// Entries that never change (identity rows 0..2 / columns 6..8 of At, padding) are written once by k_stage_static at bmpc_create.
template <int NJ>
struct SDims {
  static constexpr int NX = Dims<NJ>::NX, NXA = Dims<NJ>::NXA, NXR = NX - 3, MP = 16, NXP = 24, LDA = 42;
  static constexpr int S_AB = 0, S_B = S_AB + NXP * LDA, S_Q = S_B + NXP, S_R = S_Q + NXP, S_META = S_R + MP, TMA_DOUBLES = S_META + 8,
                       S_QF = TMA_DOUBLES, S_PRF = S_QF + 9 * 64, SREC = S_PRF + 10 * 64;
  static_assert((TMA_DOUBLES * 8) % 16 == 0 && (SREC * 8) % 16 == 0, "TMA bulk copies need 16-byte multiples");
  // meta slots
  static constexpr int T_TYPE = 0, T_MODE = 1, T_M = 2, T_MJ = 3, T_NCLOSED = 4, T_DT = 5;
  // offset of element (r, c) of a matrix with ntn column tiles stored in accumulator-fragment order (mma.m8n8k4 C layout)
  __host__ __device__ static constexpr int frag(int ntn, int r, int c) { return ((r >> 3) * ntn + (c >> 3)) * 64 + (((r & 7) << 2) + ((c & 7) >> 1)) * 2 + (c & 1); }
  __host__ __device__ static constexpr int qf(int r, int c) { return S_QF + frag(3, r, c); }
  __host__ __device__ static constexpr int prf(int r, int c) { return S_PRF + frag(5, r, c); }   // c < 24: Pt, c >= 24: Rt column c - 24
};

// one-time initialisation of the static entries of every stage record (the buffer is zeroed before)
template <int NJ>
__global__ void k_stage_static(double* stage, size_t nrec) {
  using S = SDims<NJ>;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrec) return;
  double* so = stage + i * S::SREC;
  for (int r = 0; r < 3; ++r) { so[S::S_AB + r * S::LDA + r] = 1.0; so[S::S_AB + (6 + r) * S::LDA + 6 + r] = 1.0; }
}

// ------------------------------------------------------------------------------------------------ upstream's projection: Eigen::FullPivLU
// LinearAlgebra::luConstraintProjection [UPSTREAM] (selected by projectStateInputEqualityConstraints, task.info:76):
//   lu = FullPivLU(D);  Pu = lu.kernel();  Px = -lu.solve(C);  Pe = -lu.solve(e)
// i.e. Gaussian elimination with complete pivoting, rank = #pivots above eps * min(rows, cols) * |largest pivot|, solve() satisfies the
// `rank` pivot rows exactly, ignores the dependent ones (the sixth row of every stance foot) and sets the free unknowns to zero, kernel()
// is [-U11^-1 U12; I] in the pivot ordering.  Restated here for the structure of this problem: the zero-force rows of open contacts are
// identity rows on force columns that no other row touches, so they are pivots whose elimination changes nothing; the closed-contact force
// columns are zero and stay free; what remains is the complete-pivoting elimination of the joint block Dv (NR x NJ raw velocity rows), with
// the identity pivots only entering the rank threshold (largest pivot >= 1, min(rows, cols) of the full matrix).
// (tests/test_oracle_kat.py::test_complete_pivoting_on_joint_block_equals_fullpivlu_on_stacked_rows checks this reduction on the CPU.)
// Implementation: Gauss-Jordan on the augmented columns [Dv | Cv | ev], one column per lane (G1: 34 columns, lanes 0 and 1 carry a second
// one), rows in registers and never swapped physically.  Per pivot: lane-local maximum over the unused rows, exact warp arg-max of the
// 64-bit pattern with two REDUX steps + ballot (first maximum in column order, as Eigen's column-major visitor), the multipliers of the pivot
// column are broadcast through a double-buffered shared-memory vector, every lane eliminates its own column (above AND below the pivot, so
// no back substitution is needed: afterwards row r of column c holds U11^-1 L^-1 P c up to the pivot scale).  The result columns
// -col[row(k)] / pivot(k) are scattered into W (shared memory) at the pivot column's joint index.
// c[r] for a warp-uniform r
template <int NR>
__device__ __forceinline__ void pick_row(const double (&c)[NR], int r, double& out) {
  switch (r) {
    case 0: out = c[0]; break;
    case 1: out = c[1]; break;
    case 2: out = c[2]; break;
    case 3: out = c[3]; break;
    case 4: if constexpr (NR > 4) out = c[4]; break;
    case 5: if constexpr (NR > 5) out = c[5]; break;
    case 6: if constexpr (NR > 6) out = c[6]; break;
    case 7: if constexpr (NR > 7) out = c[7]; break;
    case 8: if constexpr (NR > 8) out = c[8]; break;
    case 9: if constexpr (NR > 9) out = c[9]; break;
    case 10: if constexpr (NR > 10) out = c[10]; break;
    default: if constexpr (NR > 11) out = c[11]; break;
  }
}

// c[r] = v for a warp-uniform r
template <int NR>
__device__ __forceinline__ void put_row(double (&c)[NR], int r, double v) {
  switch (r) {
    case 0: c[0] = v; break;
    case 1: c[1] = v; break;
    case 2: c[2] = v; break;
    case 3: c[3] = v; break;
    case 4: if constexpr (NR > 4) c[4] = v; break;
    case 5: if constexpr (NR > 5) c[5] = v; break;
    case 6: if constexpr (NR > 6) c[6] = v; break;
    case 7: if constexpr (NR > 7) c[7] = v; break;
    case 8: if constexpr (NR > 8) c[8] = v; break;
    case 9: if constexpr (NR > 9) c[9] = v; break;
    case 10: if constexpr (NR > 10) c[10] = v; break;
    default: if constexpr (NR > 11) c[11] = v; break;
  }
}

template <int NJ, int NR>
__device__ __forceinline__ int lu_project(const double* __restrict__ rec, double* __restrict__ W, int ldw, double* __restrict__ sF, int* __restrict__ sPc, double* __restrict__ sIp,
                                          int lane, int n_open, unsigned zrows, double gain) {
  using D = Dims<NJ>;
  // columns: Dv (NJ) | Cv (NXA active state columns) | ev | C[:, base height] = gain on the z rows (zero, and skipped, without positionErrorGain)
  constexpr int NXA = D::NXA, NU = D::NU, NCOLS = NJ + NXA + 2, SD = NR < NJ ? NR : NJ;
  constexpr bool TWO = NCOLS > 32;
  static_assert(NCOLS <= 64 && NR % 2 == 0 && NR <= D::MAXROWS, "column / row capacity of the in-warp elimination");
  double c0[NR], c1[TWO ? NR : 1];
  {
    const int j = lane;
    const double* src = j < NJ ? rec + D::R_DV + j : (j < NJ + NXA ? rec + D::R_CV + (j - NJ) : rec + D::R_EV);
    const int stride = j < NJ ? NJ : (j < NJ + NXA ? NXA : 1);
    const bool on = j <= NJ + NXA, z8 = j == NJ + NXA + 1;
#pragma unroll
    for (int i = 0; i < NR; ++i) c0[i] = on ? src[i * stride] : ((z8 && ((zrows >> i) & 1u)) ? gain : 0.0);
    if constexpr (TWO) {
      const int j2 = lane + 32;
      const double* src2 = j2 < NJ + NXA ? rec + D::R_CV + (j2 - NJ) : rec + D::R_EV;
      const int stride2 = j2 < NJ + NXA ? NXA : 1;
      const bool on2 = j2 <= NJ + NXA, z82 = j2 == NJ + NXA + 1;
#pragma unroll
      for (int i = 0; i < NR; ++i) c1[i] = on2 ? src2[i * stride2] : ((z82 && ((zrows >> i) & 1u)) ? gain : 0.0);
    }
  }
  if (lane < NR) sPc[lane] = -1;
  // Rows are swapped physically (as Eigen does): after step kk the pivot rows sit in positions 0..kk, so the search runs over the static
  // range i >= kk.  perm holds the ORIGINAL row number of every position (4 bits each, warp uniform): the tie rule is stated in original order.
  unsigned long long perm = 0xBA9876543210ull;
  bool used = false, stop = false; int rank = 0;
  double maxpiv = n_open > 0 ? 1.0 : 0.0;
  const int sdfull = (NR + 3 * n_open) < NU ? (NR + 3 * n_open) : NU;
  const double epsd = 2.220446049250313e-16 * (double)sdfull;
#pragma unroll
  for (int kk = 0; kk < SD; ++kk) {
    if (!stop) {   // warp uniform
      double best = -1.0;
#pragma unroll
      for (int i = kk; i < NR; ++i) { const double a_ = fabs(c0[i]); if (a_ > best) best = a_; }
      const bool cand = lane < NJ && !used;
      const unsigned long long key = cand ? (unsigned long long)__double_as_longlong(best) : 0ull;
      const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
      const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
      const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
      if ((mhi | mlo) == 0u) stop = true;   // the remaining block is exactly zero (Eigen: m_nonzero_pivots = k)
      else {
        // coefficients within PIVOT_TIE of the maximum are tied (the last pivot of a stance foot's block is an exact tie between two rows that
        // rounding noise would decide): the first one in (original column, original row) order wins -- same rule as the oracle's emulation
        const double tie = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo)) * (1.0 - 1e-10);
        int bo = 16, bi = NR;   // smallest original row number among the tied rows of the own column, and its position
#pragma unroll
        for (int i = kk; i < NR; ++i) { const int o_ = (int)((perm >> (4 * i)) & 15ull); if (fabs(c0[i]) >= tie && o_ < bo) { bo = o_; bi = i; } }
        const unsigned win = __ballot_sync(0xffffffffu, cand && bi < NR);
        const int bl = __ffs(win) - 1;
        const int prow = __shfl_sync(0xffffffffu, bi, bl);   // position of the pivot row (>= kk)
        // own elements of the pivot row: prow is warp uniform, so a switch (one uniform branch) replaces a chain of selects
        double cp0 = 0.0, cp1 = 0.0;
        pick_row<NR>(c0, prow, cp0);
        if constexpr (TWO) pick_row<NR>(c1, prow, cp1);
        const double p = __shfl_sync(0xffffffffu, cp0, bl);
        const double ap = fabs(p), mp = fmax(maxpiv, ap);
        if (!(ap > epsd * mp)) stop = true;   // below Eigen's rank threshold: with complete pivoting everything that follows is, too
        else {
          maxpiv = mp;
          // swap positions kk <-> prow
          put_row<NR>(c0, prow, c0[kk]); c0[kk] = cp0;
          if constexpr (TWO) { put_row<NR>(c1, prow, c1[kk]); c1[kk] = cp1; }
          {
            const unsigned long long ok_ = (perm >> (4 * kk)) & 15ull, op_ = (perm >> (4 * prow)) & 15ull;
            perm = (perm & ~((15ull << (4 * kk)) | (15ull << (4 * prow)))) | (op_ << (4 * kk)) | (ok_ << (4 * prow));
          }
          const double ip = 1.0 / p;
          double* F = sF + (kk & 1) * 16;
          if (lane == bl) {   // multipliers of the pivot column; the pivot row itself is not eliminated (multiplier 0)
#pragma unroll
            for (int i = 0; i < NR; ++i) F[i] = (i == kk) ? 0.0 : c0[i] * ip;
            used = true;
          }
          if (lane == 0) { sPc[kk] = bl; sIp[kk] = ip; }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < NR; i += 2) {
            const double2 f = reinterpret_cast<const double2*>(F)[i >> 1];
            c0[i] = fma(-f.x, cp0, c0[i]); c0[i + 1] = fma(-f.y, cp0, c0[i + 1]);
            if constexpr (TWO) { c1[i] = fma(-f.x, cp1, c1[i]); c1[i + 1] = fma(-f.y, cp1, c1[i + 1]); }
          }
          ++rank;
        }
      }
    }
  }
  __syncwarp();
  // scatter: W column of this lane's column (full-state order, affine column 6, kernel columns 24 + t in the order of the free joint columns)
  const unsigned free_mask = __ballot_sync(0xffffffffu, lane < NJ && !used);
  auto wcol = [&](int j) {
    if (j < NJ) { const int w = 24 + __popc(free_mask & ((1u << j) - 1u)); return (used || w >= 32) ? -1 : w; }
    if (j < NJ + NXA) { const int c = j - NJ; return c < 6 ? c : c + 3; }
    return j == NJ + NXA ? 6 : (j == NJ + NXA + 1 ? 8 : -1);
  };
  const int wl0 = wcol(lane);
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int pc = sPc[i];
    if (pc >= 0) {
      const double s_ = -sIp[i];
      if (wl0 >= 0) W[pc * ldw + wl0] = c0[i] * s_;
      if constexpr (TWO) { const int wl1 = (lane + 32 < NCOLS) ? wcol(lane + 32) : -1; if (wl1 >= 0) W[pc * ldw + wl1] = c1[i] * s_; }
    }
  }
  if (lane < NJ && !used && wl0 >= 0) W[lane * ldw + wl0] = 1.0;
  return rank;
}

// Lane roles after the projection (one column of W = [Px | Pe | N] per lane, in FULL-STATE column order so that the tensor-core tiles line up with
// the stage record): lane L < 24 = state column L (L = 6 carries the affine column Pe: base-position columns 6..8 of Px are structurally
// zero; 7, 8 and the padding lanes stay zero), lane 24 + t = null-space column t.  The reduced input is ordered [null-space (mj) | closed-contact
// forces (3 nclosed)], so the null rows / columns are tile aligned as well.
// The change of input variables runs on the FP64 tensor cores:
//   M  = W^T (Rj_eff W) (32 x 32)  : tiles (a, b < 3) are Qt in accumulator-fragment order (stored with one 16-byte store per lane and tile),
//                                    row 6 / column 6 hold the qt / rt corrections, tiles (3, b < 3) are Pt, tile (3, 3) the null block of Rt;
//   AJ = B_d[:, joints] W (16 x 32): joint part of At rows 3..11 (stored as row-major pairs), bt, null-space columns of Bt.
// Z^T = W^T Rj is formed first and reused from registers as the B operand of M (same register-chaining trick as k_riccati_warp).
template <int NJ, bool LU>
__global__ void __launch_bounds__(128, PROJ_BLOCKS) k_project(Dev d) {
  using D = Dims<NJ>; using S = SDims<NJ>;
  constexpr int NX = D::NX, NU = D::NU, NXA = D::NXA, MP = S::MP, LDA = S::LDA;
  constexpr int WPB = 4;
  constexpr int LDW = 34, LDR = 18, LDJ = 20;   // leading dimensions = 2 mod 4: k-permuted fragment loads are conflict free
  __shared__ double sM[LU ? 1 : WPB][NJ][12];     // QR: Dv^T  (NJ x r), r <= 10
  __shared__ double sV[LU ? 1 : WPB][10][NJ];     // QR: Householder vectors (zero padded)
  __shared__ __align__(16) double sBeta[WPB][32]; // QR: beta (10) | 1 / R[k][k] (10);  LU: multipliers of the pivot column, double buffered (2 x 16)
  __shared__ double sIpv[WPB][12];                // LU: reciprocal pivot of the step that used row i
  __shared__ int sPcl[WPB][12];                   // LU: joint column of the pivot that used row i (-1: dependent row, ignored as by FullPivLU::solve)
  __shared__ double sG[WPB][10][NXA + 1];   // QR: [Cv | ev]; afterwards (both variants): the padded joint block of B_d
  __shared__ double sBd[WPB][9 * (12 + NJ)];  // B_d rows 3..11
  __shared__ __align__(16) double sW[WPB][16][LDW];        // W, rows >= NJ zero
  __shared__ double sMisc[WPB][32];          // r_j (16) | open-contact correction of bt rows 3..11 (16)
  __shared__ double sRjP[16][LDR];           // joint block of R (model constant), zero padded
  __shared__ double sQd[24];
  for (int rr_ = threadIdx.x >> 5; rr_ < 16; rr_ += 4) { const int cc_ = threadIdx.x & 31; if (cc_ < LDR) sRjP[rr_][cc_] = (rr_ < NJ && cc_ < NJ) ? c_model.Rjoint[rr_ * NJ + cc_] : 0.0; }
  if (threadIdx.x < 24) sQd[threadIdx.x] = threadIdx.x < NX ? c_model.Qdiag[threadIdx.x] : 0.0;
  __syncthreads();
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // broadcast: lets the compiler treat the warp index as warp-uniform
  const int gw = blockIdx.x * WPB + warp;
  const int b = gw / d.NS, k = gw % d.NS;
  if (b >= d.B) return;
  if (PF_AHEAD >= 0 && lane == 0 && (size_t)gw + PF_AHEAD < (size_t)d.B * d.NS) prefetch_l2(d.lq + ((size_t)gw + PF_AHEAD) * D::REC, D::REC * sizeof(double));
  const size_t nb = (size_t)b * d.NS;
  const double* rec = d.lq + (nb + k) * D::REC;
  // one round of loads instead of a chain (n_nodes -> node_ev -> row count): the stage type / row count / mode of the LQ record are requested together
  // with n_nodes (a slot beyond the horizon holds stale but valid memory and is dropped below)
  const int Nn = d.n_nodes[b];
  const double m_type = rec[D::R_MISC + D::M_TYPE], m_rows = rec[D::R_MISC + D::M_NROWS], m_mode = rec[D::R_MISC + D::M_MODE];
  const int N = Nn - 1;
  if (k >= N) return;
  double* out = d.proj + (nb + k) * D::PREC;
  double* so = d.stage + (nb + k) * S::SREC;
  if (m_type != 0.0) {   // event stage (k_lq_pack marks it in the record): only b is needed
    for (int i = lane; i < NX; i += 32) so[S::S_B + i] = rec[D::R_B + i];
    if (lane == 0) { so[S::S_META + S::T_TYPE] = 1.0; so[S::S_META + S::T_M] = 0.0; so[S::S_META + S::T_MJ] = 0.0; so[S::S_META + S::T_NCLOSED] = 0.0; so[S::S_META + S::T_DT] = 0.0; so[S::S_META + S::T_MODE] = -1.0; }
    return;
  }
  const DevModel& M = c_model;
  const int r = (int)m_rows;
  const int mode = (int)m_mode;
  const bool st0 = leg_in_stance(mode, 0), st1 = leg_in_stance(mode, 1);
  double (*G)[NXA + 1] = sG[warp];
  double (*W)[LDW] = sW[warp];
  const double* Bd = sBd[warp];
  bool anomaly = false;
  int mj = 0;
  // lane roles (see the header comment)
  const bool is_x = lane < 6 || (lane >= 9 && lane < NX), is_aff = lane == 6, is_rhs = is_x || is_aff;
  const int gc = is_aff ? NXA : (lane < 6 ? lane : lane - 3);   // column of [Cv | ev] / compressed column index of this lane
  double y[NJ];
  if constexpr (LU) {
    {   // stage B_d rows 3..11 and r_j; zero W
      constexpr int N3 = (9 * NU + 31) / 32;
      double t3[N3];
#pragma unroll
      for (int i = 0; i < N3; ++i) { const int e = lane + 32 * i; t3[i] = e < 9 * NU ? rec[D::R_BD + e] : 0.0; }
      const double trj = lane < NJ ? rec[D::R_R + 12 + lane] : 0.0;
      for (int i = lane; i < 16 * LDW / 2; i += 32) reinterpret_cast<double2*>(&W[0][0])[i] = make_double2(0.0, 0.0);
#pragma unroll
      for (int i = 0; i < N3; ++i) { const int e = lane + 32 * i; if (e < 9 * NU) sBd[warp][e] = t3[i]; }
      if (lane < 16) sMisc[warp][lane] = trj;
    }
    const int n_open = 4 - 2 * (int(st0) + int(st1));
    // z rows of the raw stack (contact 0..3: closed -> third of its three rows, open -> its single row): where positionErrorGain enters C[:, base height]
    unsigned zrows; { const unsigned z0 = st0 ? 0x24u : 0x3u; const int n0 = st0 ? 6 : 2; zrows = z0 | ((st1 ? 0x24u : 0x3u) << n0); }
    int rank;
    switch (r) {   // raw velocity rows: 3 per closed contact, 1 per open contact
      case 4: rank = lu_project<NJ, 4>(rec, &W[0][0], LDW, sBeta[warp], sPcl[warp], sIpv[warp], lane, n_open, zrows, M.gain); break;
      case 8: rank = lu_project<NJ, 8>(rec, &W[0][0], LDW, sBeta[warp], sPcl[warp], sIpv[warp], lane, n_open, zrows, M.gain); break;
      default: rank = lu_project<NJ, 12>(rec, &W[0][0], LDW, sBeta[warp], sPcl[warp], sIpv[warp], lane, n_open, zrows, M.gain); break;
    }
    const int expect = (st0 ? 5 : 2) + (st1 ? 5 : 2);
    anomaly = rank < (expect < NJ ? expect : NJ);
    mj = NJ - rank;
    if (mj > 8) { mj = 8; anomaly = true; }   // the null-space lanes / record hold 8 directions (H1 FLY: 6, G1 FLY: 8)
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NJ; ++i) y[i] = W[i][lane];
    {   // projection record: one destination pointer / stride per lane (Pxj column | Pej | null-space column (zero beyond mj) | Pxj[:, base height]), one store loop
      const bool g8 = lane == 8 && M.gain != 0.0;   // read only with positionErrorGain
      double* pd = is_x ? out + D::P_PX + gc : (is_aff ? out + D::P_PE : (lane >= 24 ? out + D::P_N + (lane - 24) : out + D::P_PX8));
      const int ps = is_x ? NXA : (lane >= 24 ? 8 : 1);
      if (is_x || is_aff || lane >= 24 || g8) {
#pragma unroll
        for (int i = 0; i < NJ; ++i) pd[i * ps] = y[i];
      }
    }
  } else {
  double (*Mt)[12] = sM[warp]; double (*V)[NJ] = sV[warp]; double* beta = sBeta[warp]; double* rinv = sBeta[warp] + 10;
  {   // stage Dv^T, [Cv | ev] and B_d rows 3..11: all global loads are issued before the first shared-memory store (fixed trip counts;
      // rows >= r hold stale but finite data and are never used)
    constexpr int N1 = (10 * NJ + 31) / 32, N2 = (10 * NXA + 31) / 32, N3 = (9 * NU + 31) / 32;
    double t1[N1], t2[N2], t3[N3];
#pragma unroll
    for (int i = 0; i < N1; ++i) { const int e = lane + 32 * i; t1[i] = e < 10 * NJ ? rec[D::R_DV + e] : 0.0; }
#pragma unroll
    for (int i = 0; i < N2; ++i) { const int e = lane + 32 * i; t2[i] = e < 10 * NXA ? rec[D::R_CV + e] : 0.0; }
#pragma unroll
    for (int i = 0; i < N3; ++i) { const int e = lane + 32 * i; t3[i] = e < 9 * NU ? rec[D::R_BD + e] : 0.0; }
    const double tev = lane < 10 ? rec[D::R_EV + lane] : 0.0;
    const double trj = lane < NJ ? rec[D::R_R + 12 + lane] : 0.0;
#pragma unroll
    for (int i = 0; i < N1; ++i) { const int e = lane + 32 * i; if (e < 10 * NJ) Mt[e % NJ][e / NJ] = t1[i]; }
#pragma unroll
    for (int i = 0; i < N2; ++i) { const int e = lane + 32 * i; if (e < 10 * NXA) G[e / NXA][e % NXA] = t2[i]; }
#pragma unroll
    for (int i = 0; i < N3; ++i) { const int e = lane + 32 * i; if (e < 9 * NU) sBd[warp][e] = t3[i]; }
    if (lane < 10) G[lane][NXA] = tev;
    if (lane < 16) sMisc[warp][lane] = trj;
  }
  for (int i = lane; i < 10 * NJ; i += 32) V[i / NJ][i % NJ] = 0.0;
  __syncwarp();
  double rmax = 0.0;
  // Householder QR with compile-time trip counts (rows beyond r are skipped by the warp-uniform test kk < r).  Lane c < r keeps its
  // column of Dv^T in registers; the reflector of column kk is broadcast from lane kk with shuffles (no shared-memory round trips).
  double colv[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) colv[i] = (lane < r) ? Mt[i][lane] : 0.0;
#pragma unroll
  for (int kk = 0; kk < 10; ++kk) {
    if (kk < r) {
      double vk[NJ];
      double nrm2 = 0.0;
#pragma unroll
      for (int i = 0; i < NJ; ++i) { vk[i] = (i >= kk) ? __shfl_sync(0xffffffffu, colv[i], kk) : 0.0; nrm2 += vk[i] * vk[i]; }
      const double x0 = vk[kk];
      const double nrm = nrm2 > 0.0 ? nrm2 * rsqrt(nrm2) : 0.0;
      const double alpha = x0 >= 0.0 ? -nrm : nrm;
      const double v0 = x0 - alpha;
      const double vtv = nrm2 - x0 * x0 + v0 * v0;
      const double bta = vtv > 0.0 ? 2.0 * __drcp_rn(vtv) : 0.0;
      rmax = fmax(rmax, nrm);
      if (!(nrm > 1e-9 * rmax)) anomaly = true;
      vk[kk] = v0;
      if (lane > kk && lane < r) {   // apply the reflector to the own column
        double sdot = 0.0;
#pragma unroll
        for (int i = 0; i < NJ; ++i) if (i >= kk) sdot += vk[i] * colv[i];
        sdot *= bta;
#pragma unroll
        for (int i = 0; i < NJ; ++i) if (i >= kk) colv[i] -= sdot * vk[i];
      }
      if (lane == kk) {
#pragma unroll
        for (int i = 0; i < NJ; ++i) { colv[i] = (i == kk) ? alpha : ((i > kk) ? 0.0 : colv[i]); V[kk][i] = vk[i]; }
        beta[kk] = bta; rinv[kk] = __drcp_rn(alpha);   // 1 / R[kk][kk] for the triangular solves (inf on a rank anomaly, which is flagged)
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NJ; ++i) if (lane < r) Mt[i][lane] = colv[i];   // R (upper triangle) for the triangular solve below
  __syncwarp();
  mj = NJ - r;
  const bool is_null = lane >= 24 && lane - 24 < mj;
#pragma unroll
  for (int i = 0; i < NJ; ++i) y[i] = 0.0;
  if (is_rhs) {   // z = R^-T g  (R^T lower triangular: R[l][i] = Mt[l][i] for l <= i)
#pragma unroll
    for (int i = 0; i < NJ; ++i) if (i < r) {
      double s_ = G[i][gc];
#pragma unroll
      for (int l = 0; l < NJ; ++l) if (l < i) s_ -= Mt[l][i] * y[l];
      y[i] = s_ * rinv[i];
    }
  } else if (is_null) {
    const int t = lane - 24;
#pragma unroll
    for (int i = 0; i < NJ; ++i) if (i == r + t) y[i] = 1.0;
  }
  if (is_rhs || is_null) {
    for (int kk = r - 1; kk >= 0; --kk) {   // y <- H_kk y
      double s_ = 0.0;
#pragma unroll
      for (int i = 0; i < NJ; ++i) s_ += V[kk][i] * y[i];
      s_ *= beta[kk];
#pragma unroll
      for (int i = 0; i < NJ; ++i) y[i] -= s_ * V[kk][i];
    }
    if (is_rhs) {
#pragma unroll
      for (int i = 0; i < NJ; ++i) y[i] = -y[i];
      if (is_x) { for (int i = 0; i < NJ; ++i) out[D::P_PX + i * NXA + gc] = y[i]; }
      else { for (int i = 0; i < NJ; ++i) out[D::P_PE + i] = y[i]; }
    } else {
      const int t = lane - 24;
#pragma unroll
      for (int i = 0; i < NJ; ++i) out[D::P_N + i * 8 + t] = y[i];
    }
  } else if (lane >= 24) {
#pragma unroll
    for (int i = 0; i < NJ; ++i) out[D::P_N + i * 8 + lane - 24] = 0.0;   // unused null-space columns: never leave stale data behind
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) W[i][lane] = (i < NJ) ? y[i < NJ ? i : 0] : 0.0;   // idle lanes hold y = 0
  }
  const double dt = rec[D::R_MISC + D::M_DT], dq = rec[D::R_MISC + D::M_DQ], dr = rec[D::R_MISC + D::M_DR];
  const int nclosed = 2 * (int(st0) + int(st1));
  const int m = 3 * nclosed + mj;
  if (lane < 12) out[D::P_FO + lane] = rec[D::R_FO + lane];
  if (lane == 0) {
    out[D::P_META] = (double)mj; out[D::P_META + 1] = anomaly ? 1.0 : 0.0; out[D::P_META + 2] = (double)mode; if (anomaly) atomicOr(&d.status[b], 2);
    double* mt_ = so + S::S_META;
    mt_[S::T_TYPE] = 0.0; mt_[S::T_MODE] = (double)mode; mt_[S::T_M] = (double)m; mt_[S::T_MJ] = (double)mj; mt_[S::T_NCLOSED] = (double)nclosed; mt_[S::T_DT] = dt;
  }
  __syncwarp();   // [Cv | ev] is dead from here on
  // ---------------- change of input variables
  const int g = lane >> 2, q = lane & 3;
  // original force column of reduced force index cf (closed contacts only)
  auto force_col = [&](int cf) { return st0 ? cf : 6 + cf; };
  // joint block of B_d rows 3..11 (columns zero padded to 16; fragment rows beyond 8 re-read row 8, results discarded) in the storage of [Cv | ev]
  static_assert(9 * LDJ <= 10 * (NXA + 1), "Bj must fit into the [Cv | ev] buffer");
  double (*Bj)[LDJ] = reinterpret_cast<double (*)[LDJ]>(&G[0][0]);
  if (lane < LDJ) {
#pragma unroll
    for (int rr_ = 0; rr_ < 9; ++rr_) Bj[rr_][lane] = (lane < NJ) ? Bd[rr_ * NU + 12 + lane] : 0.0;
  }
  // contribution of the fixed open-contact forces (du_F = -F) to rows 3..11 of bt, one row per lane 0..8
  if (lane < 16) {
    double open_corr = 0.0;
    if (lane < 9) {
      for (int cn = 0; cn < NCON; ++cn) if (!(cn / 2 == 0 ? st0 : st1))
        for (int qq = 0; qq < 3; ++qq) open_corr -= Bd[lane * NU + 3 * cn + qq] * rec[D::R_FO + 3 * cn + qq];
    }
    sMisc[warp][16 + lane] = open_corr;
  }
  // ---- element-wise parts (lane = column of W, values in y[]; done first so that y[] is dead during the tile products)
  {   // At rows 12.. = I + dt Pxj (lane = state column; lane 8, the base-height column, only with positionErrorGain), Bt rows 12.. = dt N (lanes 24..: reduced
      // columns 0..7, zero beyond mj) and bt rows 12.. (affine lane): the same row of [At | Bt] for the first two, so one destination pointer / stride per lane
    const bool rowst = is_x || (lane == 8 && M.gain != 0.0) || lane >= 24;
    double* pd = is_aff ? so + S::S_B + 12 : so + S::S_AB + 12 * LDA + lane;
    const int ps = is_aff ? 1 : LDA;
#pragma unroll
    for (int l = 0; l < NJ; ++l) {
      const double bl = rec[D::R_B + 12 + l];
      const double v = dt * y[l] + (is_aff ? bl : ((12 + l == lane) ? 1.0 : 0.0));
      if (rowst || is_aff) pd[l * ps] = v;
    }
  }
  if (is_aff) {
    const double f = dt / M.total_mass;
    for (int qq = 0; qq < 3; ++qq) {   // rows 0..2 of bt: B_d rows 0..2 = dt/m on the force columns
      double bb = rec[D::R_B + qq];
      for (int cn = 0; cn < NCON; ++cn) if (!(cn / 2 == 0 ? st0 : st1)) bb -= f * rec[D::R_FO + 3 * cn + qq];
      so[S::S_B + qq] = bb;
    }
  }
  // Bt rows 0..2 (dt/m on the closed-contact force columns) and rows 3..11 of the reduced columns 8..15 (force columns or zero)
  for (int i = lane; i < 3 * MP + 9 * 8; i += 32) {
    int r_, c; double v = 0.0;
    if (i < 3 * MP) { r_ = i / MP; c = i % MP; const int cf = c - mj; if (cf >= 0 && cf < 3 * nclosed && cf % 3 == r_) v = dt / M.total_mass; }
    else { const int e = i - 3 * MP; r_ = 3 + e / 8; c = 8 + e % 8; const int cf = c - mj; if (cf >= 0 && cf < 3 * nclosed) v = Bd[(r_ - 3) * NU + force_col(cf)]; }
    so[S::S_AB + r_ * LDA + 24 + c] = v;
  }
  // rt: closed-contact force entries, zero padding (the null-space entries come from the tile products)
  if (lane < MP) { const int cf = lane - mj; if (cf >= 0) so[S::S_R + lane] = (cf < 3 * nclosed) ? rec[D::R_R + force_col(cf)] : 0.0; }
  __syncwarp();
  // accumulator initialisers of AJ (issued early): A_d - I rows 3..11 / b rows 3..11 + open-contact correction
  double ad[2][4][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const int rr = 8 * mt + g, C = 8 * nt + 2 * q + sl;
        double v = 0.0;
        if (rr < 9) {
          if (C < 6 || (C >= 9 && C < NX)) v = rec[D::R_AD + rr * NXA + (C < 6 ? C : C - 3)];
          else if (C == 6) v = rec[D::R_B + 3 + rr] + sMisc[warp][16 + rr];
        }
        ad[mt][nt][sl] = v;
      }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) { ad[mt][3][0] = 0.0; ad[mt][3][1] = 0.0; }
  // ---- step 1: Z^T = W^T Rj (32 x 16), then Z = dt Z + dr W (+ r_j on the affine column): Z[mt][nt] holds (Rj_eff W)[8 nt + 2q + s][8 mt + g]
  double a[2][2][4], Z[4][2][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) { Z[mt][nt][0] = 0.0; Z[mt][nt][1] = 0.0; }
#pragma unroll
  for (int kb = 0; kb < 2; ++kb)
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) a[kb][sl][mt] = W[8 * kb + 2 * q + sl][8 * mt + g];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const double bR = sRjP[8 * kb + 2 * q + sl][8 * nt + g];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) dmma884(Z[mt][nt][0], Z[mt][nt][1], a[kb][sl][mt], bR);
      }
    }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        double v = dt * Z[mt][nt][sl] + dr * a[nt][sl][mt];
        if (mt == 0 && g == 6) v += sMisc[warp][8 * nt + 2 * q + sl];   // affine column 6: t1 = r_j + Rj_eff Pe
        Z[mt][nt][sl] = v;
      }
  // ---- step 2: M = W^T (Rj_eff W), tile by tile, stored straight from the accumulator fragments
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (mt < 3 && nt == 3) continue;   // N-columns of the state rows: the transpose of Pt, not needed
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) dmma884(c0, c1, a[kb][sl][mt], Z[nt][kb][sl]);
      const int R = 8 * mt + g, C0 = 8 * nt + 2 * q;
      if (mt < 3 && nt < 3) {          // Qt tile (R, C0 .. C0+1): row / column 6 carry the affine terms, the diagonal gets dt Q + dq
        if (nt == 0 && q == 3 && R < NX && R != 6) so[S::S_Q + R] = rec[D::R_Q + R] + ((R == 7) ? 0.0 : c0);   // qt = q + Px^T t1 (row 8: non-zero only with positionErrorGain)
        double v0 = (R == 6 || C0 == 6) ? 0.0 : c0, v1 = (R == 6) ? 0.0 : c1;
        if (R == C0 && R < NX) v0 += dt * sQd[R] + dq;
        if (R == C0 + 1 && R < NX) v1 += dt * sQd[R] + dq;
        // Qt is symmetric: tiles below the diagonal are not stored (k_riccati_warp computes the upper tiles of S' only and mirrors them)
        if (mt <= nt) *reinterpret_cast<double2*>(so + S::S_QF + (mt * 3 + nt) * 64 + 2 * lane) = make_double2(v0, v1);
      } else if (nt < 3) {             // mt == 3: Pt rows t = g (zero beyond mj); column 6 is the rt correction of the null-space inputs
        if (nt == 0 && q == 3 && g < mj) so[S::S_R + g] = c0;
        *reinterpret_cast<double2*>(so + S::S_PRF + (0 * 5 + nt) * 64 + 2 * lane) = make_double2((C0 == 6) ? 0.0 : c0, c1);
      } else {                         // mt == nt == 3: null block of Rt = rows / columns 0..7 of Rt; the force / identity part F is added
        double v[2] = {c0, c1};
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int r_ = g, c = 2 * q + sl;
          if (r_ >= mj || c >= mj) {
            double f = 0.0;
            if (r_ >= m || c >= m) f = (r_ == c) ? 1.0 : 0.0;
            else if (r_ >= mj && c >= mj && (r_ - mj) / 3 == (c - mj) / 3) {
              const int cn = (st0 ? 0 : 2) + (r_ - mj) / 3, p_ = (r_ - mj) % 3, q_ = (c - mj) % 3;
              const int lo = p_ < q_ ? p_ : q_, hi = p_ < q_ ? q_ : p_;
              f = rec[D::R_HB + 6 * cn + (lo == 0 ? hi : (lo == 1 ? 2 + hi : 5))];
              if (p_ == q_) f += dt * M.Rforce[3 * cn + p_] + dr;
            }
            v[sl] = f;
          }
        }
        *reinterpret_cast<double2*>(so + S::S_PRF + (0 * 5 + 3) * 64 + 2 * lane) = make_double2(v[0], v[1]);
      }
    }
  // the other three tiles of Rt (rows or columns 8..15): force blocks / identity only
#pragma unroll
  for (int tt = 1; tt < 4; ++tt) {
    const int ta = tt >> 1, tb = tt & 1;
    double v[2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const int r_ = 8 * ta + g, c = 8 * tb + 2 * q + sl;
      double f = 0.0;
      if (r_ >= m || c >= m) f = (r_ == c) ? 1.0 : 0.0;
      else if (r_ >= mj && c >= mj && (r_ - mj) / 3 == (c - mj) / 3) {
        const int cn = (st0 ? 0 : 2) + (r_ - mj) / 3, p_ = (r_ - mj) % 3, q_ = (c - mj) % 3;
        const int lo = p_ < q_ ? p_ : q_, hi = p_ < q_ ? q_ : p_;
        f = rec[D::R_HB + 6 * cn + (lo == 0 ? hi : (lo == 1 ? 2 + hi : 5))];
        if (p_ == q_) f += dt * M.Rforce[3 * cn + p_] + dr;
      }
      v[sl] = f;
    }
    *reinterpret_cast<double2*>(so + S::S_PRF + (ta * 5 + 3 + tb) * 64 + 2 * lane) = make_double2(v[0], v[1]);
  }
  // ---- step 3: AJ = B_d[:, joints] W (16 x 32): joint part of At rows 3..11, bt rows 3..11, null-space columns of Bt rows 3..11
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    double bj[2][2];
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) bj[kb][sl] = Bj[(8 * mt + g) < 9 ? 8 * mt + g : 8][8 * kb + 2 * q + sl];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      double c0 = ad[mt][nt][0], c1 = ad[mt][nt][1];
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) dmma884(c0, c1, bj[kb][sl], a[kb][sl][nt]);
      const int rr = 8 * mt + g, C0 = 8 * nt + 2 * q;
      if (rr < 9) {
        const int sr_ = 3 + rr;
        if (nt < 3) {     // At row 3 + rr, state columns C0, C0 + 1 (columns 6..8: identity entries; column 6 of the product is bt)
          if (C0 == 6) so[S::S_B + sr_] = c0;
          const double v0 = ((C0 == 6) ? 0.0 : c0) + ((sr_ == C0) ? 1.0 : 0.0);   // column 6 carries bt; column 8 (base height) is zero without positionErrorGain
          const double v1 = ((C0 + 1 == 7) ? 0.0 : c1) + ((sr_ == C0 + 1) ? 1.0 : 0.0);
          *reinterpret_cast<double2*>(so + S::S_AB + sr_ * LDA + C0) = make_double2(v0, v1);
        } else {          // Bt row 3 + rr, reduced columns 2q, 2q + 1: null-space columns, then closed-contact force columns, then zero
          double v[2] = {c0, c1};
#pragma unroll
          for (int sl = 0; sl < 2; ++sl) { const int cf = 2 * q + sl - mj; if (cf >= 0) v[sl] = (cf < 3 * nclosed) ? Bd[rr * NU + force_col(cf)] : 0.0; }
          *reinterpret_cast<double2*>(so + S::S_AB + sr_ * LDA + 24 + 2 * q) = make_double2(v[0], v[1]);
        }
      }
    }
  }
  // qt of the base-position rows and the rows the tiles do not reach
  if (lane == 6) so[S::S_Q + 6] = rec[D::R_Q + 6];
}

}  // namespace bmpc
