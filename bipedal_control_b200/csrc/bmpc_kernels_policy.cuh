// K2b / K3: gains and closed-loop stage maps, forward substitution
// (part of bmpc_kernels.cuh: include that header, not this file)
#pragma once

namespace bmpc {

// ------------------------------------------------------------------------------------------------ K2b: gains and closed-loop stage maps, one warp per (instance, stage)
//   Kt = -L^-T Y, kt = -L^-T yg;  K = Px + Pu Kt, kappa = Pe + Pu kt, uff0 = u - K x   ([UPSTREAM] remapProjectedGain / toPrimalSolution)
//   Phi = At + Bt Kt, phi = bt + Bt kt (forward substitution), ghat = qt + Kt^T rt, misc = rt^T kt (armijoDescentMetric)
template <int NJ>
struct PolSmem {
  static constexpr int NX = Dims<NJ>::NX, MP = 16;
  static constexpr int NT = (NX + 1 + 7) / 8, LDK = NT * 8 + 4;   // column tiles of [Kt | kt] (H1: 3, G1: 4); ld = 4 or 12 mod 16
  double L[MP][MP + 1];
  double Kt[MP * LDK];           // [Kt | kt | 0]
  double P[(12 + NJ) * 25];      // K[r][c] * x[c] (row sums give K x)
  double rt[MP], xk[24], Nn[NJ * 8];
};

// value of the padded operand [At | bt | 0] (24 x 24) at (r, c), read from the stage record (rows >= NX of AB and bt are zero padding)
template <int NJ>
__device__ __forceinline__ double stage_At_aug(const double* __restrict__ sr, int r, int c) {
  using S = SDims<NJ>; constexpr int NX = Dims<NJ>::NX;
  if (c < NX) return sr[S::S_AB + r * S::LDA + c];
  return c == NX ? sr[S::S_B + r] : 0.0;
}

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
__global__ void __launch_bounds__(128, 4) k_policy_expand(Dev d) {
  using D = Dims<NJ>; using R = RDims<NJ>; using S = SDims<NJ>; using PS = PolSmem<NJ>;
  constexpr int NX = D::NX, NU = D::NU, NXA = D::NXA, MP = S::MP, WPB = 4, NT = PS::NT, LDK = PS::LDK, NTILES = 3 * NT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // broadcast: lets the compiler treat the warp index as warp-uniform
  PS& sm = reinterpret_cast<PS*>(smem_raw)[warp];
  const int gw = blockIdx.x * WPB + warp;
  const int b = gw / d.NS, k = gw % d.NS;
  if (b >= d.B) return;
  const int N = d.n_nodes[b] - 1;
  if (k >= N) return;
  const size_t nb = (size_t)b * d.NS;
  const double* __restrict__ sr = d.stage + (nb + k) * S::SREC;
  double* __restrict__ ric = d.ric + (nb + k) * R::KREC;
  double* __restrict__ Kg = d.s_K + (nb + k) * (size_t)(NU * NX);
  double* __restrict__ uffg = d.s_uff + (nb + k) * NU;
  const double* meta = sr + S::S_META;
  if (meta[S::T_TYPE] != 0.0) {   // event stage: K = 0, Phi = I, phi = b
    for (int i = lane; i < NU; i += 32) { ric[R::K_KAP + i] = 0.0; uffg[i] = 0.0; }
    for (int i = lane; i < NX; i += 32) { ric[R::K_SPHI + i] = sr[S::S_B + i]; ric[R::K_G + i] = 0.0; }
    for (int i = lane; i < NX * NX; i += 32) ric[R::K_PHI + i] = (i / NX == i % NX) ? 1.0 : 0.0;
    for (int i = lane; i < NU * NX; i += 32) Kg[i] = 0.0;
    if (lane == 0) { ric[R::K_MISC] = 0.0; ric[R::K_MISC + 1] = 1.0; }
    return;
  }
  const int m = (int)meta[S::T_M], mj = (int)meta[S::T_MJ], nclosed = (int)meta[S::T_NCLOSED], mode = (int)meta[S::T_MODE];
  const double dt = meta[S::T_DT];
  const bool st0 = leg_in_stance(mode, 0), st1 = leg_in_stance(mode, 1);
  const double* __restrict__ prj = d.proj + (nb + k) * D::PREC;
  const int lr = lane >> 2, lc = lane & 3;
  // ---- issue every global load up front (independent: their latency overlaps with the back substitution below)
  const bool active = lane <= NX;
  double z[MP];
#pragma unroll
  for (int i = 0; i < MP; ++i) z[i] = (active && i < m) ? (lane < NX ? ric[R::K_Y + i * NX + lane] : ric[R::K_YG + i]) : 0.0;   // rows >= m: padding, not stored
  double lreg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) lreg[i] = (((lane + 32 * i) >> 4) < m) ? ric[R::K_L + lane + 32 * i] : 0.0;
  // accumulators of the 9 output tiles initialised with [At | bt | 0]; A fragments of Bt (24 x 16)
  double c0[NTILES], c1[NTILES], af[3][4];
#pragma unroll
  for (int t = 0; t < NTILES; ++t) { const int r = 8 * (t / NT) + lr, c = 8 * (t % NT) + 2 * lc; c0[t] = stage_At_aug<NJ>(sr, r, c); c1[t] = stage_At_aug<NJ>(sr, r, c + 1); }
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int r = 8 * mt + lr, c = 4 * kk + lc;
      af[mt][kk] = sr[S::S_AB + r * S::LDA + 24 + c];
    }
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
  // ---- [Phi | phi] = [At | bt] + Bt [Kt | kt] on the FP64 tensor cores: 9 tiles x 4 k-steps, results stored straight from the fragments
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    if (4 * kk < m) {   // rows >= m of [Kt | kt] are zero (warp-uniform skip)
#pragma unroll
      for (int t = 0; t < NTILES; ++t) dmma884(c0[t], c1[t], af[t / NT][kk], sm.Kt[(4 * kk + lc) * LDK + 8 * (t % NT) + lr]);
    }
#pragma unroll
  for (int t = 0; t < NTILES; ++t) {
    const int r = 8 * (t / NT) + lr, c = 8 * (t % NT) + 2 * lc;
    if (r < NX) {
      if (c < NX) ric[R::K_PHI + r * NX + c] = c0[t]; else if (c == NX) ric[R::K_SPHI + r] = c0[t];
      if (c + 1 < NX) ric[R::K_PHI + r * NX + c + 1] = c1[t]; else if (c + 1 == NX) ric[R::K_SPHI + r] = c1[t];
    }
  }
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
template <int NJ>
struct FwdSmem {
  static constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU, LDP = NX + 1;
  double Phi[2][NX * LDP], K[2][NU * LDP], v[2][3 * NX + NU + 2];   // double buffered: stage k+1 is fetched while stage k is applied
  double dx[NX];
};

template <int NJ>
__global__ void __launch_bounds__(128) k_forward(Dev d) {
  using D = Dims<NJ>; using R = RDims<NJ>; using FS = FwdSmem<NJ>;
  constexpr int NX = D::NX, NU = D::NU, WPB = 4, LDP = FS::LDP;
  constexpr int NPH = (NX * NX + 31) / 32, NK = (NU * NX + 31) / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // broadcast: lets the compiler treat the warp index as warp-uniform
  FS& sm = reinterpret_cast<FS*>(smem_raw)[warp];
  const int b = blockIdx.x * WPB + warp;
  if (b >= d.B) return;
  const int N = d.n_nodes[b] - 1;
  const size_t nb = (size_t)b * d.NS;
  double* dx = sm.dx;
  // dx_0 = x0 - x[0]
  double s0 = 0.0;
  if (lane < NX) { const double e = d.x0[(size_t)b * NX + lane] - d.s_x[nb * NX + lane]; dx[lane] = e; d.dx[nb * NX + lane] = e; s0 = e * e; }
  double armijo = 0.0, dxn = s0, dun = 0.0, pc = 0.0, pd = 0.0, pe = 0.0;
  double rphi[NPH], rk[NK], rv[4], rperf[3] = {0.0, 0.0, 0.0};
  auto fetch = [&](int k) {   // global -> registers (all loads independent, in flight while the previous stage is applied)
    const double* ric = d.ric + (nb + k) * R::KREC;
    const double* Kg = d.s_K + (nb + k) * (size_t)(NU * NX);
#pragma unroll
    for (int i = 0; i < NPH; ++i) { const int e = lane + 32 * i; rphi[i] = e < NX * NX ? ric[R::K_PHI + e] : 0.0; }
#pragma unroll
    for (int i = 0; i < NK; ++i) { const int e = lane + 32 * i; rk[i] = e < NU * NX ? Kg[e] : 0.0; }
    rv[0] = lane < NX ? ric[R::K_SPHI + lane] : 0.0; rv[1] = lane < NX ? ric[R::K_G + lane] : 0.0; rv[2] = lane < NU ? ric[R::K_KAP + lane] : 0.0;
    rv[3] = lane < 2 ? ric[R::K_MISC + lane] : 0.0;
    if (lane < 3) rperf[lane] = d.lq[(nb + k) * D::REC + D::R_MISC + D::M_PCOST + lane];
  };
  auto stash = [&](int buf) {   // registers -> shared memory buffer
#pragma unroll
    for (int i = 0; i < NPH; ++i) { const int e = lane + 32 * i; if (e < NX * NX) sm.Phi[buf][(e / NX) * LDP + e % NX] = rphi[i]; }
#pragma unroll
    for (int i = 0; i < NK; ++i) { const int e = lane + 32 * i; if (e < NU * NX) sm.K[buf][(e / NX) * LDP + e % NX] = rk[i]; }
    double* v = sm.v[buf];
    if (lane < NX) { v[lane] = rv[0]; v[NX + lane] = rv[1]; }
    if (lane < NU) v[2 * NX + lane] = rv[2];
    if (lane < 2) v[3 * NX + NU + lane] = rv[3];
  };
  if (N > 0) { fetch(0); stash(0); }
  __syncwarp();
  for (int k = 0; k < N; ++k) {
    const int buf = k & 1;
    if (lane == 0) { pc += rperf[0]; } if (lane == 1) pd += rperf[1]; if (lane == 2) pe += rperf[2];
    if (k + 1 < N) fetch(k + 1);
    const double* Phi = sm.Phi[buf]; const double* Kk = sm.K[buf]; const double* v = sm.v[buf];
    const double misc = v[3 * NX + NU];
    const bool is_event = v[3 * NX + NU + 1] != 0.0;
    double nx_ = 0.0, du_ = 0.0, ga = 0.0;
    if (lane < NX) {
      double a = v[lane];
#pragma unroll
      for (int c = 0; c < NX; ++c) a += Phi[lane * LDP + c] * dx[c];
      nx_ = a;
      ga = v[NX + lane] * dx[lane];
    }
    if (lane < NU) {
      double a = v[2 * NX + lane];
#pragma unroll
      for (int c = 0; c < NX; ++c) a += Kk[lane * LDP + c] * dx[c];
      du_ = is_event ? 0.0 : a; d.du[(nb + k) * NU + lane] = du_;
    }
    __syncwarp();
    if (lane < NX) { dx[lane] = nx_; d.dx[(nb + k + 1) * NX + lane] = nx_; }
    armijo += ga + (lane == 0 ? misc : 0.0);
    dxn += nx_ * nx_; dun += du_ * du_;
    if (k + 1 < N) stash(buf ^ 1);
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
