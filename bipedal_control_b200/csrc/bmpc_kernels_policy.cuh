// K2b / K3: gains and closed-loop stage maps, forward substitution
// (part of bmpc_kernels.cuh: include that header, not this file)
#pragma once

namespace bmpc {

// ------------------------------------------------------------------------------------------------ K2b: gains and closed-loop stage maps, one warp per (instance, stage)
//   Kt = -L^-T Y, kt = -L^-T yg;  K = Px + Pu Kt, kappa = Pe + Pu kt, uff0 = u - K x   ([UPSTREAM] remapProjectedGain / toPrimalSolution)
//   ghat = qt + Kt^T rt, misc = rt^T kt (armijoDescentMetric).  The closed-loop stage map is NOT materialised: k_forward applies the
//   original dynamics dx+ = A_d dx + B_d (K dx + kappa) + b from the compact LQ record, which equals (At + Bt Kt) dx + bt + Bt kt.
template <int NJ>
struct PolSmem {
  static constexpr int NX = Dims<NJ>::NX, MP = 16;
  static constexpr int NT = (NX + 1 + 7) / 8, LDK = NT * 8 + 4;   // column tiles of [Kt | kt] (H1: 3, G1: 4); ld = 4 or 12 mod 16
  double L[MP][MP + 1];
  double Kt[MP * LDK];           // [Kt | kt | 0]
  double P[(12 + NJ) * 25];      // K[r][c] * x[c] (row sums give K x)
  double rt[MP], xk[24], Nn[NJ * 8];
};

// z <- L^-T z for the leading M x M block (the record stores 1 / L[i][i] on the diagonal); rows >= M are left as they are (zero)
template <int M, int MP>
__device__ __forceinline__ void back_substitute(double (&z)[MP], const double (&L)[MP][MP + 1]) {
#pragma unroll
  for (int i = M - 1; i >= 0; --i) {
    double a = z[i];
#pragma unroll
    for (int l = i + 1; l < M; ++l) a -= L[l][i] * z[l];
    z[i] = a * L[i][i];
  }
}

template <int NJ>
__global__ void __launch_bounds__(128, EXP_BLOCKS) k_policy_expand(Dev d) {
  using D = Dims<NJ>; using R = RDims<NJ>; using S = SDims<NJ>; using PS = PolSmem<NJ>;
  constexpr int NX = D::NX, NU = D::NU, NXA = D::NXA, MP = S::MP, WPB = 4, LDK = PS::LDK;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // broadcast: lets the compiler treat the warp index as warp-uniform
  PS& sm = reinterpret_cast<PS*>(smem_raw)[warp];
  const int gw = blockIdx.x * WPB + warp;
  const int b = gw / d.NS, k = gw % d.NS;
  if (b >= d.B) return;
  const size_t nb = (size_t)b * d.NS;
  const double* __restrict__ sr = d.stage + (nb + k) * S::SREC;
  const double* meta = sr + S::S_META;
  // one round of loads instead of a chain (n_nodes -> stage meta): a slot beyond the horizon holds stale but valid memory and is dropped below
  const int Nn = d.n_nodes[b];
  const double m_type = meta[S::T_TYPE], m_m = meta[S::T_M], m_mj = meta[S::T_MJ], m_nc = meta[S::T_NCLOSED], m_mode = meta[S::T_MODE];
  const int N = Nn - 1;
  if (k >= N) return;
  double* __restrict__ ric = d.ric + (nb + k) * R::KREC;
  double* __restrict__ Kg = d.s_K + (nb + k) * (size_t)(NU * NX);
  double* __restrict__ uffg = d.s_uff + (nb + k) * NU;
  if (m_type != 0.0) {   // event stage: K = 0 (k_forward: dx+ = dx + b, du = 0)
    for (int i = lane; i < NU; i += 32) { ric[R::K_KAP + i] = 0.0; uffg[i] = 0.0; }
    for (int i = lane; i < NX; i += 32) ric[R::K_G + i] = 0.0;
    for (int i = lane; i < NU * NX; i += 32) Kg[i] = 0.0;
    if (lane == 0) { ric[R::K_MISC] = 0.0; ric[R::K_MISC + 1] = 1.0; }
    return;
  }
  const int m = (int)m_m, mj = (int)m_mj, nclosed = (int)m_nc, mode = (int)m_mode;
  if (PF_AHEAD >= 0 && (size_t)gw + PF_AHEAD < (size_t)d.B * d.NS) {   // inputs of the stage one wave ahead -> L2 (its m is guessed to be this stage's)
    const size_t ga = (size_t)gw + PF_AHEAD;
    const int mg = m > 0 ? m : 9;
    if (lane == 0) prefetch_l2(d.ric + ga * R::KREC + R::K_Y, (unsigned)(mg * NX * sizeof(double)));
    if (lane == 1) prefetch_l2(d.ric + ga * R::KREC + R::K_YG, (unsigned)((MP + mg * MP) * sizeof(double)));
    if (lane == 2) prefetch_l2(d.proj + ga * D::PREC, (unsigned)(D::PREC * sizeof(double)));
    if (lane == 3) prefetch_l2(d.stage + ga * S::SREC + S::S_B, (unsigned)((S::TMA_DOUBLES - S::S_B) * sizeof(double)));
  }
  const bool st0 = leg_in_stance(mode, 0), st1 = leg_in_stance(mode, 1);
  const double* __restrict__ prj = d.proj + (nb + k) * D::PREC;
  // ---- issue every global load up front (independent: their latency overlaps with the back substitution below)
  const bool active = lane <= NX;
  double z[MP];
#pragma unroll
  for (int i = 0; i < MP; ++i) z[i] = (active && i < m) ? (lane < NX ? ric[R::K_Y + i * NX + lane] : ric[R::K_YG + i]) : 0.0;   // rows >= m: padding, not stored
  double lreg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) lreg[i] = (((lane + 32 * i) >> 4) < m) ? ric[R::K_L + lane + 32 * i] : 0.0;
  const double rt_l = (lane < MP) ? sr[S::S_R + lane] : 0.0;
  const double xk_l = (lane < NX) ? d.s_x[(nb + k) * NX + lane] : 0.0;
  const double qt_l = (lane < NX) ? sr[S::S_Q + lane] : 0.0;
  double nn[(NJ * 8 + 31) / 32];
#pragma unroll
  for (int i = 0; i < (NJ * 8 + 31) / 32; ++i) nn[i] = (lane + 32 * i < NJ * 8) ? prj[D::P_N + lane + 32 * i] : 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const int e = lane + 32 * i; sm.L[e / MP][e % MP] = lreg[i]; }
  if (lane < MP) sm.rt[lane] = rt_l;
  if (lane < 24) sm.xk[lane] = xk_l;
#pragma unroll
  for (int i = 0; i < (NJ * 8 + 31) / 32; ++i) if (lane + 32 * i < NJ * 8) sm.Nn[lane + 32 * i] = nn[i];
  __syncwarp();
  // ---- back substitution Kt = -L^-T Y (lane c < NX: column c; lane NX: kt from yg); trip counts fixed by the reduced input dimension m
  switch (m) {   // H1: 6 / 9 / 12 (FLY / single stance / double stance), G1: 8 / 11 / 14
    case 6: back_substitute<6, MP>(z, sm.L); break;
    case 9: back_substitute<9, MP>(z, sm.L); break;
    case 12: back_substitute<12, MP>(z, sm.L); break;
    case 8: back_substitute<8, MP>(z, sm.L); break;
    case 11: back_substitute<11, MP>(z, sm.L); break;
    case 14: back_substitute<14, MP>(z, sm.L); break;
    default: back_substitute<MP, MP>(z, sm.L); break;   // rows >= m of z and of L (incl. the reciprocal pivots) were loaded as zero
  }
#pragma unroll
  for (int i = 0; i < MP; ++i) { z[i] = -z[i]; if (lane < LDK) sm.Kt[i * LDK + lane] = z[i]; }
  __syncwarp();
  // ---- ghat = qt + Kt^T rt ; misc = rt^T kt
  if (active) {
    double gh = qt_l;
#pragma unroll
    for (int j = 0; j < MP; ++j) gh += sm.rt[j] * z[j];
    if (lane < NX) ric[R::K_G + lane] = gh; else { ric[R::K_MISC] = gh; ric[R::K_MISC + 1] = 0.0; }
  }
  // ---- K = Px + Pu Kt, kappa = Pe + Pu kt ; products K[r][c] x[c] staged in shared memory for uff0 = u - K x
  if (active) {
#pragma unroll
    for (int r = 0; r < 12; ++r) {
      const int cn = r / 3; const bool cl = (cn / 2 == 0) ? st0 : st1;
      double a = 0.0;
      if (cl) a = sm.Kt[(mj + (st0 ? cn : cn - 2) * 3 + r % 3) * LDK + lane];   // reduced inputs: [null space (mj) | closed-contact forces]
      else if (lane == NX) a = -prj[D::P_FO + r];
      if (lane < NX) { Kg[r * NX + lane] = a; sm.P[r * 25 + lane] = a * xk_l; } else ric[R::K_KAP + r] = a;
    }
    double zn[8];   // null-space part of the own column
#pragma unroll
    for (int t = 0; t < 8; ++t) zn[t] = (t < mj) ? sm.Kt[t * LDK + lane] : 0.0;
    const bool xact = lane < 6 || (lane >= 9 && lane < NX);
    const int xc = xcol(lane);
    double pxv[NJ];
#pragma unroll
    for (int l = 0; l < NJ; ++l) pxv[l] = (lane < NX) ? (xact ? prj[D::P_PX + l * NXA + xc] : 0.0) : prj[D::P_PE + l];
    if (c_model.gain != 0.0 && lane == 8) {   // positionErrorGain: the base-height column of Pxj is not structurally zero
#pragma unroll
      for (int l = 0; l < NJ; ++l) pxv[l] = prj[D::P_PX8 + l];
    }
#pragma unroll
    for (int l = 0; l < NJ; ++l) {
      double a = pxv[l];
#pragma unroll
      for (int t = 0; t < 8; ++t) a += sm.Nn[l * 8 + t] * zn[t];
      const int r = 12 + l;
      if (lane < NX) { Kg[r * NX + lane] = a; sm.P[r * 25 + lane] = a * xk_l; } else ric[R::K_KAP + r] = a;
    }
  }
  __syncwarp();
  if (lane < NU) {
    double kx = 0.0;
#pragma unroll
    for (int c = 0; c < NX; ++c) kx += sm.P[lane * 25 + c];
    uffg[lane] = d.s_u[(nb + k) * NU + lane] - kx;
  }
}

// ------------------------------------------------------------------------------------------------ K3: forward substitution, one warp per instance
//   du_k = K_k dx_k + kappa_k,  dx_{k+1} = A_d dx_k + B_d du_k + b   (the ORIGINAL dynamics of the compact LQ record; identical to the closed-loop map
//   (At + Bt Kt) dx + bt + Bt kt of the projected problem because K = Px + Pu Kt, kappa = Pe + Pu kt -- so no Phi is ever stored).
//   Structure of the RK2 sensitivities: rows 0..2 of A_d - I vanish and B_d rows 0..2 are dt/m [I I I I]; rows 12.. of A_d - I
//   vanish and B_d rows 12.. are dt I; only rows 3..11 are dense: (A_d - I)[9 x NXA] on the active state columns and B_d[9 x NU].
//   Lane roles per stage: phase 1: lane L < NU -> du[L]; lanes NU..NU+8 -> the A_d part of state rows 3..11 (independent of du, same time);
//   phase 2 (after du is visible): lanes NU..NU+8 finish rows 3..11 with B_d du, lanes 0..2 and 12..NX-1 their own (structured) rows.
template <int NJ>
struct FwdSmem {
  static constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU, NXA = Dims<NJ>::NXA, LDP = NX + 1, LDA = NXA | 1, LDB = NU | 1;
  // single buffered: stage k+1 waits in registers (fetch) while stage k is applied from shared memory, and is stashed afterwards
  double K[NU * LDP], AD[9 * LDA], BD[9 * LDB], v[2 * NX + NU + 4];   // v: b | ghat | kappa | armijo term, event flag, dt
  double dx[NX], du[NU];
};

template <int NJ>
__global__ void __launch_bounds__(32 * FWD_WPC, FWD_BLOCKS) k_forward(Dev d) {
  using D = Dims<NJ>; using R = RDims<NJ>; using FS = FwdSmem<NJ>;
  constexpr int NX = D::NX, NU = D::NU, NXA = D::NXA, WPB = FWD_WPC, LDP = FS::LDP, LDA = FS::LDA, LDB = FS::LDB;
  constexpr int NK = (NU * NX + 31) / 32, NA = (9 * NXA + 31) / 32, NB = (9 * NU + 31) / 32;
  constexpr int FIT = (32 - NU) < 9 ? (32 - NU) : 9;   // dense rows whose A_d part fits on the lanes beyond the du rows (H1: all 9, G1: 8)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // broadcast: lets the compiler treat the warp index as warp-uniform
  FS& sm = reinterpret_cast<FS*>(smem_raw)[warp];
  const int b = blockIdx.x * WPB + warp;
  if (b >= d.B) return;
  const int N = d.n_nodes[b] - 1;
  const size_t nb = (size_t)b * d.NS;
  double* dx = sm.dx; double* du = sm.du;
  const double imass = 1.0 / c_model.total_mass;
  // dx_0 = x0 - x[0]
  double s0 = 0.0;
  if (lane < NX) { const double e = d.x0[(size_t)b * NX + lane] - d.s_x[nb * NX + lane]; dx[lane] = e; d.dx[nb * NX + lane] = e; s0 = e * e; }
  double armijo = 0.0, dxn = s0, dun = 0.0, pc = 0.0, pd = 0.0, pe = 0.0;
  double rk[NK], ra[NA], rb[NB], rv[5], rperf[3] = {0.0, 0.0, 0.0};
  auto fetch = [&](int k) {   // global -> registers (all loads independent, in flight while the previous stage is applied)
    const double* ric = d.ric + (nb + k) * R::KREC;
    const double* lq = d.lq + (nb + k) * D::REC;
    const double* Kg = d.s_K + (nb + k) * (size_t)(NU * NX);
#pragma unroll
    for (int i = 0; i < NK; ++i) { const int e = lane + 32 * i; rk[i] = e < NU * NX ? Kg[e] : 0.0; }
#pragma unroll
    for (int i = 0; i < NA; ++i) { const int e = lane + 32 * i; ra[i] = e < 9 * NXA ? lq[D::R_AD + e] : 0.0; }
#pragma unroll
    for (int i = 0; i < NB; ++i) { const int e = lane + 32 * i; rb[i] = e < 9 * NU ? lq[D::R_BD + e] : 0.0; }
    rv[0] = lane < NX ? lq[D::R_B + lane] : 0.0; rv[1] = lane < NX ? ric[R::K_G + lane] : 0.0; rv[2] = lane < NU ? ric[R::K_KAP + lane] : 0.0;
    rv[3] = lane < 2 ? ric[R::K_MISC + lane] : 0.0; rv[4] = lq[D::R_MISC + D::M_DT];
    if (lane < 3) rperf[lane] = lq[D::R_MISC + D::M_PCOST + lane];
  };
  auto stash = [&]() {   // registers -> shared memory
#pragma unroll
    for (int i = 0; i < NK; ++i) { const int e = lane + 32 * i; if (e < NU * NX) sm.K[(e / NX) * LDP + e % NX] = rk[i]; }
#pragma unroll
    for (int i = 0; i < NA; ++i) { const int e = lane + 32 * i; if (e < 9 * NXA) sm.AD[(e / NXA) * LDA + e % NXA] = ra[i]; }
#pragma unroll
    for (int i = 0; i < NB; ++i) { const int e = lane + 32 * i; if (e < 9 * NU) sm.BD[(e / NU) * LDB + e % NU] = rb[i]; }
    double* v = sm.v;
    if (lane < NX) { v[lane] = rv[0]; v[NX + lane] = rv[1]; }
    if (lane < NU) v[2 * NX + lane] = rv[2];
    if (lane < 2) v[2 * NX + NU + lane] = rv[3];
    if (lane == 2) v[2 * NX + NU + 2] = rv[4];
  };
  if (N > 0) { fetch(0); stash(); }
  __syncwarp();
  // state row this lane completes in phase 2 (-1: none): lanes NU..NU+FIT-1 -> dense rows 3.., lanes 0..2 and 12..NX-1 -> their own row;
  // dense rows that do not fit beyond the du lanes (G1: row 11) go to the otherwise idle lane of the same number and do their A_d part in phase 2
  const int row2 = (lane >= NU && lane < NU + FIT) ? lane - NU + 3 : ((lane < 3 || (lane >= 12 && lane < NX) || (lane >= 3 + FIT && lane < 12)) ? lane : -1);
  auto ad_part = [&](int i) {   // (A_d - I) dx of dense row i (state row 3 + i)
    const double* Ar = sm.AD + i * LDA;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int c = 0; c < 6; c += 3) { a0 += Ar[c] * dx[c]; a1 += Ar[c + 1] * dx[c + 1]; a2 += Ar[c + 2] * dx[c + 2]; }
#pragma unroll
    for (int c = 6; c + 2 < NXA; c += 3) { a0 += Ar[c] * dx[c + 3]; a1 += Ar[c + 1] * dx[c + 4]; a2 += Ar[c + 2] * dx[c + 5]; }
#pragma unroll
    for (int c = 6 + ((NXA - 6) / 3) * 3; c < NXA; ++c) a0 += Ar[c] * dx[c + 3];
    return a0 + a1 + a2;
  };
  for (int k = 0; k < N; ++k) {
    if (lane == 0) { pc += rperf[0]; } if (lane == 1) pd += rperf[1]; if (lane == 2) pe += rperf[2];
    if (k + 1 < N) fetch(k + 1);
    const double* Kk = sm.K; const double* v = sm.v;
    const double misc = v[2 * NX + NU];
    const bool is_event = v[2 * NX + NU + 1] != 0.0;
    const double dt = v[2 * NX + NU + 2];
    double du_ = 0.0, ga = 0.0, part = 0.0;
    if (lane < NU) {   // du = kappa + K dx, four independent partial sums (the dependent-FMA chain is what this kernel waits for)
      double a0 = v[2 * NX + lane], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int c = 0; c + 3 < NX; c += 4) { a0 += Kk[lane * LDP + c] * dx[c]; a1 += Kk[lane * LDP + c + 1] * dx[c + 1]; a2 += Kk[lane * LDP + c + 2] * dx[c + 2]; a3 += Kk[lane * LDP + c + 3] * dx[c + 3]; }
#pragma unroll
      for (int c = NX & ~3; c < NX; ++c) a0 += Kk[lane * LDP + c] * dx[c];
      du_ = is_event ? 0.0 : (a0 + a1) + (a2 + a3);
      du[lane] = du_; d.du[(nb + k) * NU + lane] = du_;
      ga = v[NX + lane] * dx[lane];   // NX == NU
    } else if (lane < NU + FIT && !is_event) part = ad_part(lane - NU);   // state row 3 + (lane - NU)
    __syncwarp();   // du visible
    double nx_ = 0.0;
    if (row2 >= 0) {
      double a = dx[row2] + v[row2];
      if (!is_event) {
        if (row2 < 3) a += dt * imass * ((du[row2] + du[3 + row2]) + (du[6 + row2] + du[9 + row2]));
        else if (row2 >= 12) a += dt * du[row2];
        else {
          const double* Br = sm.BD + (row2 - 3) * LDB;
          if (FIT < 9 && lane < NU) part = ad_part(row2 - 3);
          double a0 = part, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
          for (int c = 0; c + 3 < NU; c += 4) { a0 += Br[c] * du[c]; a1 += Br[c + 1] * du[c + 1]; a2 += Br[c + 2] * du[c + 2]; a3 += Br[c + 3] * du[c + 3]; }
#pragma unroll
          for (int c = NU & ~3; c < NU; ++c) a0 += Br[c] * du[c];
          a += (a0 + a1) + (a2 + a3);
        }
      }
      nx_ = a;
    }
    __syncwarp();   // every lane has read dx / du / the stage operands
    if (row2 >= 0) { dx[row2] = nx_; d.dx[(nb + k + 1) * NX + row2] = nx_; }
    armijo += ga + (lane == 0 ? misc : 0.0);
    dxn += nx_ * nx_; dun += du_ * du_;
    if (k + 1 < N) stash();
    __syncwarp();
  }
  pc = __shfl_sync(0xffffffffu, pc, 0); pd = __shfl_sync(0xffffffffu, pd, 1); pe = __shfl_sync(0xffffffffu, pe, 2);
  for (int o = 16; o > 0; o >>= 1) { armijo += __shfl_xor_sync(0xffffffffu, armijo, o); dxn += __shfl_xor_sync(0xffffffffu, dxn, o); dun += __shfl_xor_sync(0xffffffffu, dun, o); s0 += __shfl_xor_sync(0xffffffffu, s0, o); }
  if (lane == 0) {
    double* pf = d.perf + (size_t)b * 8;
    pf[0] = pc; pf[1] = pd + s0; pf[2] = pe; pf[7] = armijo;
    d.norms[2 * b] = sqrt(dxn); d.norms[2 * b + 1] = sqrt(dun);
  }
}

}  // namespace bmpc
