// K2: backward Riccati recursion (warp per instance: default; CTA per instance: cross-check), TMA / mbarrier helpers, shuffle-only Cholesky
// (part of bmpc_kernels.cuh: include that header, not this file)
#pragma once

namespace bmpc {

// ------------------------------------------------------------------------------------------------ TMA bulk copy + mbarrier helpers (sm_90+/sm_100a PTX)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one thread: arm the barrier with the byte count and launch the bulk copy global -> shared (UBLKCP in SASS)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Right-looking Cholesky of the (symmetric, fully stored) M x M matrix G fused with the forward substitution of [H | g], one warp,
// shuffles only.  Lane l < MP holds column l of G in gc[], lane c holds column c of [H | g] in hc[].  On return lane j holds
// column j of L in gc[] (rows > j; the diagonal entry holds 1 / L[j][j]) and hc[] holds Y = L^-1 H (yg in the g lane).  Serial chain per pivot: shuffle -> rsqrt -> FMA.
template <int M, int MP>
__device__ __forceinline__ bool chol_forward(double (&gc)[MP], double (&hc)[MP], int lane) {
  bool not_pd = false;
#pragma unroll
  for (int j = 0; j < M; ++j) {
    double dj = __shfl_sync(0xffffffffu, gc[j], j);
    if (!(dj > 0.0)) { not_pd = true; dj = 1.0; }
    const double inv = rsqrt(dj);
    const double yj = hc[j] * inv;
    const double gj = (lane > j) ? gc[j] * inv : 0.0;   // L[lane][j] by symmetry of the fully stored G (own column, row j); finished columns stay untouched
    hc[j] = yj;
#pragma unroll
    for (int i = j + 1; i < M; ++i) {
      const double li = __shfl_sync(0xffffffffu, gc[i], j) * inv;   // L[i][j]
      hc[i] -= li * yj;
      gc[i] -= li * gj;
    }
    if (lane == j) {
      gc[j] = inv;   // the reciprocal of the pivot is what the back substitution in k_policy_expand needs (no divisions there)
#pragma unroll
      for (int i = j + 1; i < M; ++i) gc[i] *= inv;
    }
  }
  return not_pd;
}

template <int NJ>
struct RDims {
  static constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU, MP = 16;
  // written by k_riccati: Y[MP][NX], yg[MP], L[MP][MP];  written by k_policy_expand: kappa, ghat, misc (armijo term, event flag)
  static constexpr int K_Y = 0, K_YG = K_Y + MP * NX, K_L = K_YG + MP, K_KAP = K_L + MP * MP, K_G = K_KAP + NU,
                       K_MISC = K_G + NX, KREC = ((K_MISC + 2 + 3) / 4) * 4;
};

// Step C of a Riccati stage for a compile-time reduced input dimension M: [H | G | g] (rows < M) from the staged fragments -> one column per lane,
// chol_forward, the record of the stage for k_policy_expand (Y (M x NX), yg (M), L (M x MP, reciprocal pivots on the diagonal); rows >= M are
// padding: never written, never read) and [Y | yg] back into shared memory for the S' update (rows M .. 4 ceil(M / 4) - 1 zero).
// Every lane runs the same straight-line code: one clamped column index per lane and predicated stores, no divergent branches.
// HG layout per row: H (24) | G (16) | g (column 40).  Returns s' correction sum_i Y[i][lane] yg[i] for lane < 24.
template <int M, int MP, int NX, int LDH>
__device__ __forceinline__ bool chol_stage(double* __restrict__ HG, int lane, double* __restrict__ ric, double& ytyg) {
  using R = RDims<NX - 12>;
  constexpr int M4 = ((M + 3) / 4) * 4;
  const int hcol = lane < 24 ? lane : 40;          // lane 24 carries g / yg
  const int gcol = 24 + (lane & (MP - 1));
  const bool hon = lane <= 24, gon = lane < MP;
  double gc[MP], hc[MP];
#pragma unroll
  for (int i = 0; i < M; ++i) { const double gv_ = HG[i * LDH + gcol], hv_ = HG[i * LDH + hcol]; gc[i] = gon ? gv_ : 0.0; hc[i] = hon ? hv_ : 0.0; }
#pragma unroll
  for (int i = M; i < MP; ++i) { gc[i] = 0.0; hc[i] = 0.0; }
  __syncwarp();
  const bool not_pd = chol_forward<M, MP>(gc, hc, lane);
  const bool yst = lane < NX || lane == 24;
  const int yoff = lane < NX ? R::K_Y + lane : R::K_YG, ystr = lane < NX ? NX : 1;
#pragma unroll
  for (int i = 0; i < M; ++i) {
    if (yst) ric[yoff + i * ystr] = hc[i];
    if (gon) ric[R::K_L + i * MP + lane] = (i >= lane) ? gc[i] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < M4; ++i) if (hon) HG[i * LDH + hcol] = (i < M) ? hc[i] : 0.0;   // Y | yg
  __syncwarp();
  double a = 0.0;
#pragma unroll
  for (int i = 0; i < M; ++i) a += hc[i] * HG[i * LDH + 40];
  ytyg = a;
  return not_pd;
}

// ------------------------------------------------------------------------------------------------ K2 (default): backward Riccati recursion, ONE WARP PER INSTANCE
// No block-level barrier anywhere: the 4 warps of a CTA run 4 independent instances.  The value function S (24 x 24) never leaves the
// warp's registers: it is held as the 3 x 3 accumulator fragments of mma.sync.m8n8k4.f64 (lane (g, q) = (lane >> 2, lane & 3) owns
// S[8a + g][8b + 2q + {0,1}]).  Two observations make every product chain register-to-register:
//   (1) S is symmetric, so the accumulator fragment of tile (a, b) is at the same time the B-operand fragment of rows 8b + {2q, 2q+1},
//       columns 8a + g, provided the k index of the A operand is permuted the same way (k = 2q + slot; two DMMAs cover 8 rows of k);
//   (2) computing the TRANSPOSED products Z^T = [At | Bt]^T S leaves Z = S [At | Bt] in exactly that B-operand form again.
//   Z^T  = AB^T S                       (A operand: AB from the TMA-staged record, k-permuted transposed loads, ld 42 -> conflict free)
//   [H | G] = [Pt | Rt] + Bt^T Z        (accumulators initialised straight from the fragment-ordered record in global memory)
//   S'   = Qt + At^T Z[:, :24] - Y^T Y  (Y = L^-1 H from the in-warp Cholesky; Y staged in shared memory for the last product)
// The staged record is single buffered: the TMA for stage k-1 is issued as soon as the last AB fragment of stage k has been read, and
// lands while the warp runs the Cholesky chain.
template <int NJ>
struct RicWarpSmem {
#ifndef RIC_LDH
#define RIC_LDH 44
#endif
  static constexpr int LDH = RIC_LDH;   // row of the staged [H (24) | G (16) | g]: 41 used; 42 lets 16 warps per SM fit (measured slower)
  alignas(16) double rec[SDims<NJ>::TMA_DOUBLES];   // AB | bt | qt | rt | meta (TMA destination)
  alignas(16) double HG[16 * LDH];                   // [H | G] fragments -> column layout for the Cholesky; afterwards [Y | L]
  double sb[24];
  alignas(16) unsigned long long bar;
};

template <int NJ>
__global__ void __launch_bounds__(32 * RIC_WPC, RIC_BLOCKS) k_riccati_warp(Dev d) {
  using D = Dims<NJ>; using R = RDims<NJ>; using S = SDims<NJ>; using SM = RicWarpSmem<NJ>;
  constexpr int NX = D::NX, MP = S::MP, LDA = S::LDA, LDH = SM::LDH;
  // the staged record arrives as two bulk copies on one barrier: rows 0..NX-1 of AB and the tail [bt | qt | rt | meta]; the zero padding rows NX..23 of
  // AB are written once here and never fetched
  constexpr unsigned AB_BYTES = NX * LDA * sizeof(double), TAIL_BYTES = (S::TMA_DOUBLES - S::S_B) * sizeof(double);
  static_assert(AB_BYTES % 16 == 0 && TAIL_BYTES % 16 == 0 && (S::S_B * sizeof(double)) % 16 == 0, "bulk copies need 16-byte multiples");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // broadcast: lets the compiler treat the warp index as warp-uniform
  const int b = blockIdx.x * RIC_WPC + warp;
  if (b >= d.B) return;
  SM& sm = reinterpret_cast<SM*>(smem_raw)[warp];
  const int N = d.n_nodes[b] - 1;
  const size_t nb = (size_t)b * d.NS;
  const int g = lane >> 2, q = lane & 3;
  double Sf[3][3][2];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) { Sf[a][c][0] = 0.0; Sf[a][c][1] = 0.0; }   // terminal value function: zero (no terminal cost installed)
  double s_l = 0.0;   // lane r < 24 holds s[r]
  if (lane == 0) { mbar_init(&sm.bar, 1); fence_mbar_init(); }
  for (int i = lane; i < (S::NXP - NX) * LDA; i += 32) sm.rec[S::S_AB + NX * LDA + i] = 0.0;
  __syncwarp();
  auto stage_fetch = [&](const double* src) {   // one lane
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm.bar)), "r"(AB_BYTES + TAIL_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm.rec)), "l"(src), "r"(AB_BYTES), "r"(smem_u32(&sm.bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm.rec + S::S_B)), "l"(src + S::S_B), "r"(TAIL_BYTES), "r"(smem_u32(&sm.bar)) : "memory");
  };
  if (lane == 0 && N >= 1) { fence_proxy_async(); stage_fetch(d.stage + (nb + N - 1) * S::SREC); }
  unsigned phase_bit = 0;
  const double* sr = sm.rec;
  const double* AB = sm.rec + S::S_AB;
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    mbar_wait(&sm.bar, phase_bit);
    phase_bit ^= 1u;
    const double* __restrict__ grec = d.stage + (nb + k) * S::SREC;
    double* __restrict__ ric = d.ric + (nb + k) * R::KREC;
    const bool is_event = sr[S::S_META + S::T_TYPE] != 0.0;
    const int m = (int)sr[S::S_META + S::T_M];
    // ---- sb = s + S bt  (lane-level on the fragments: partial row sums, reduced over the 4 lanes of a quad)
    {
      double p[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) acc += Sf[a][c][0] * sr[S::S_B + 8 * c + 2 * q] + Sf[a][c][1] * sr[S::S_B + 8 * c + 2 * q + 1];
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        p[a] = acc;
      }
      if (q == 0) { sm.sb[g] = p[0]; sm.sb[8 + g] = p[1]; sm.sb[16 + g] = p[2]; }
      __syncwarp();
      const double sbv = (lane < 24) ? s_l + sm.sb[lane] : 0.0;
      __syncwarp();
      if (lane < 24) sm.sb[lane] = sbv;
      if (is_event) {   // A = I, Q = 0, no input: S unchanged, s <- s + S b
        s_l = sbv;
        __syncwarp();
        if (lane == 0 && k >= 1) { fence_proxy_async(); stage_fetch(d.stage + (nb + k - 1) * S::SREC); }
        continue;
      }
    }
    // accumulator initialisers, fragment ordered in global memory: [Pt | Rt] now, Qt below (in flight during the products)
    // (rows 8..15 of Pt belong to closed-contact forces: the state-input cross Hessian vanishes there, k_project never writes those three tiles)
    double HGf[2][5][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        if (a == 1 && c < 3) { HGf[a][c][0] = 0.0; HGf[a][c][1] = 0.0; continue; }
        const double2 v = *reinterpret_cast<const double2*>(grec + S::S_PRF + (a * 5 + c) * 64 + 2 * lane); HGf[a][c][0] = v.x; HGf[a][c][1] = v.y;
      }
    // ---- step A: Z^T = AB^T S   (Z[mt][nt] holds (S AB)[8 nt + 2q + slot][8 mt + g])
    double Z[5][3][2];
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) { Z[a][c][0] = 0.0; Z[a][c][1] = 0.0; }
#pragma unroll
    for (int kb = 0; kb < 3; ++kb)
#pragma unroll
      for (int mt = 0; mt < 5; ++mt) {
        const double a0 = AB[(8 * kb + 2 * q) * LDA + 8 * mt + g], a1 = AB[(8 * kb + 2 * q + 1) * LDA + 8 * mt + g];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) { dmma884(Z[mt][nt][0], Z[mt][nt][1], a0, Sf[nt][kb][0]); dmma884(Z[mt][nt][0], Z[mt][nt][1], a1, Sf[nt][kb][1]); }
      }
    double Sn[3][3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // S' is symmetric: only its tiles on and above the diagonal are computed (Qt is stored that way too); the end of the stage mirrors the
        // off-diagonal tiles and symmetrises the diagonal ones
        if (c >= a) { const double2 v = *reinterpret_cast<const double2*>(grec + S::S_QF + (a * 3 + c) * 64 + 2 * lane); Sn[a][c][0] = v.x; Sn[a][c][1] = v.y; }
        else { Sn[a][c][0] = 0.0; Sn[a][c][1] = 0.0; }
      }
    __syncwarp();   // sb visible to every lane
    // ---- step B: [H | G] += Bt^T Z ; g = rt + Bt^T sb
#pragma unroll
    for (int kb = 0; kb < 3; ++kb)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const double a0 = AB[(8 * kb + 2 * q) * LDA + 24 + 8 * mt + g], a1 = AB[(8 * kb + 2 * q + 1) * LDA + 24 + 8 * mt + g];
#pragma unroll
        for (int nt = 0; nt < 5; ++nt) { dmma884(HGf[mt][nt][0], HGf[mt][nt][1], a0, Z[nt][kb][0]); dmma884(HGf[mt][nt][0], HGf[mt][nt][1], a1, Z[nt][kb][1]); }
      }
    {
      double gval = 0.0, sn = 0.0;
      if (lane < MP) { gval = sr[S::S_R + lane]; for (int r = 0; r < 24; ++r) gval += AB[r * LDA + 24 + lane] * sm.sb[r]; }
      // ---- step D1: S' = Qt + At^T Z[:, :24] ; s' = qt + At^T sb
      if (lane < 24) { sn = sr[S::S_Q + lane]; for (int r = 0; r < 24; ++r) sn += AB[r * LDA + lane] * sm.sb[r]; }
      s_l = sn;
      if (lane < MP) sm.HG[lane * LDH + 40] = gval;   // g: column 40 of the staged [H | G]
    }
#pragma unroll
    for (int kb = 0; kb < 3; ++kb)
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        const double a0 = AB[(8 * kb + 2 * q) * LDA + 8 * mt + g], a1 = AB[(8 * kb + 2 * q + 1) * LDA + 8 * mt + g];
#pragma unroll
        for (int nt = mt; nt < 3; ++nt) { dmma884(Sn[mt][nt][0], Sn[mt][nt][1], a0, Z[nt][kb][0]); dmma884(Sn[mt][nt][0], Sn[mt][nt][1], a1, Z[nt][kb][1]); }
      }
    // ---- the staged record is free: prefetch the next stage while the Cholesky chain runs
    __syncwarp();
    if (lane == 0 && k >= 1) { fence_proxy_async(); stage_fetch(d.stage + (nb + k - 1) * S::SREC); }
    // ---- step C: [H | G] fragments -> shared memory -> one column per lane ; Cholesky of G fused with the forward substitution of [H | g]
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 5; ++c) *reinterpret_cast<double2*>(&sm.HG[(8 * a + g) * LDH + 8 * c + 2 * q]) = make_double2(HGf[a][c][0], HGf[a][c][1]);
    __syncwarp();
    {
      bool not_pd; double ytyg;
      switch (m) {   // reduced input dimensions that occur: H1 6 / 9 / 12 (FLY / single stance / double stance), G1 8 / 11 / 14
        case 6: not_pd = chol_stage<6, MP, NX, LDH>(sm.HG, lane, ric, ytyg); break;
        case 9: not_pd = chol_stage<9, MP, NX, LDH>(sm.HG, lane, ric, ytyg); break;
        case 12: not_pd = chol_stage<12, MP, NX, LDH>(sm.HG, lane, ric, ytyg); break;
        case 8: not_pd = chol_stage<8, MP, NX, LDH>(sm.HG, lane, ric, ytyg); break;
        case 11: not_pd = chol_stage<11, MP, NX, LDH>(sm.HG, lane, ric, ytyg); break;
        case 14: not_pd = chol_stage<14, MP, NX, LDH>(sm.HG, lane, ric, ytyg); break;
        default: not_pd = chol_stage<MP, MP, NX, LDH>(sm.HG, lane, ric, ytyg); break;   // padded pivots are identity rows
      }
      if (not_pd && lane == 0) atomicOr(&d.status[b], 1);
      if (lane < 24) s_l -= ytyg;   // s' -= Y^T yg
    }
    // ---- step E: S' -= Y^T Y  (natural k order: both operands come from the staged Y)
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      if (4 * kb < m) {
        double y[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) y[t] = sm.HG[(4 * kb + q) * LDH + 8 * t + g];
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
          for (int nt = mt; nt < 3; ++nt) dmma884(Sn[mt][nt][0], Sn[mt][nt][1], -y[mt], y[nt]);
      }
    }
    // ---- S from the upper tiles of S'.  Element (8 nt + 2q + s, 8 mt + g) of tile (nt, mt) lives in lane (2q + s) * 4 + g / 2, slot g & 1:
    //      diagonal tiles are symmetrised, (S' + S'^T) / 2; a tile below the diagonal is the transpose of its mirror image
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = mt; nt < 3; ++nt)
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int src = (2 * q + sl) * 4 + (g >> 1);
          const double t0 = __shfl_sync(0xffffffffu, Sn[mt][nt][0], src), t1 = __shfl_sync(0xffffffffu, Sn[mt][nt][1], src);
          const double tr = (g & 1) ? t1 : t0;   // transposed element
          if (mt == nt) Sf[mt][mt][sl] = 0.5 * (Sn[mt][mt][sl] + tr);
          else { Sf[mt][nt][sl] = Sn[mt][nt][sl]; Sf[nt][mt][sl] = tr; }
        }
    __syncwarp();   // HG / gv / sb are rewritten by the next stage
  }
}

}  // namespace bmpc
