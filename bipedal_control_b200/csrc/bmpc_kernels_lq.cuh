// K1: LQ approximation (packed warp-cooperative kernel + the earlier variants kept for cross-checking)
// (part of bmpc_kernels.cuh: include that header, not this file)
#pragma once

namespace bmpc {

// one column (X index c >= 6) of d f / d x from the base record: rows 3..5 -> col[0..2], rows 6..8 -> col[3..5], rows 9..11 -> col[6..8]
template <int NJ>
__device__ __forceinline__ void lq_dq_column(const double* __restrict__ bs, const double* __restrict__ u, int c, double* col) {
  using BD = BaseDims<NJ>; constexpr int NL = Dims<NJ>::NL;
  const double imass = 1.0 / c_model.total_mass;
  // operands through one pointer set per lane (base Euler-angle columns 6..8: the whole tree moves; joint columns: the joint's subtree):
  // the same straight-line code for every lane, no lane-dependent branch
  const bool isb = c < 9;
  const int kq = isb ? c - 6 : 0, j = isb ? 0 : c - 9;
  const double* J = bs + BD::B_J + BD::JS * j;
  const bool root = (j % NL) == 0;
  const v3 ak = ld3(isb ? bs + BD::B_BAX + 3 * kq : J + BD::J_A), ok = ld3(isb ? bs + BD::B_PB : J + BD::J_O);
  const SI sub = ld_si(isb ? bs + BD::B_TOT : J + BD::J_SI);
  Mom hsub; { const double* ph = isb ? bs + BD::B_HTOT : J + BD::J_HN; hsub.n = ld3(ph); hsub.p = ld3(ph + 3); }
  static_assert(BD::J_HP == BD::J_HN + 3, "subtree momentum: moment then linear part");
  const v3 wp = ld3(isb ? bs + BD::B_WE + 3 * kq : (root ? bs + BD::B_WE + 9 : J - BD::JS + BD::J_W));
  const v3 vp = ld3(isb ? bs + BD::B_VE + 3 * kq : (root ? bs + BD::B_VE + 9 : J - BD::JS + BD::J_V));
  const v3 AlinK = ld3(isb ? bs + BD::B_ALE + 3 * kq : J + BD::J_AL);
  const int leg_first = isb ? 0 : j / NL, leg_last = isb ? 1 : j / NL;
  const v3 com = ld3(bs + BD::B_COM), ptot = ld3(bs + BD::B_HTOT + 3), Ftot = ld3(bs + BD::B_FTOT);
  const double* A22i = bs + BD::B_A22I; const double* A12 = bs + BD::B_A12;
  const v3 s = cross(ok, ak);
  const v3 mom1 = cross(ak, hsub.n) + cross(s, hsub.p);
  const v3 frc1 = cross(ak, hsub.p);
  const v3 w1 = cross(ak, wp);
  const v3 v1 = cross(ak, vp) + cross(s, wp);
  const Mom m2 = si_apply(sub, w1, v1);
  const v3 dlin = frc1 - m2.p;
  const v3 dnO = mom1 - m2.n;
  const v3 dcom = imass * AlinK;
  const v3 dang = dnO - cross(dcom, ptot) - cross(com, dlin);
  const v3 e = mk(A22i[0] * dang.x + A22i[1] * dang.y + A22i[2] * dang.z, A22i[3] * dang.x + A22i[4] * dang.y + A22i[5] * dang.z, A22i[6] * dang.x + A22i[7] * dang.y + A22i[8] * dang.z);
  const v3 l = imass * (dlin - mk(A12[0] * e.x + A12[1] * e.y + A12[2] * e.z, A12[3] * e.x + A12[4] * e.y + A12[5] * e.z, A12[6] * e.x + A12[7] * e.y + A12[8] * e.z));
  col[3] = -l.x; col[4] = -l.y; col[5] = -l.z; col[6] = -e.x; col[7] = -e.y; col[8] = -e.z;
  v3 t = mk(0.0, 0.0, 0.0);
#pragma unroll
  for (int cc = 0; cc < NCON; ++cc) {
    const v3 e = cross(cross(ak, ld3(bs + BD::B_PC + 3 * cc) - ok), mk(u[3 * cc], u[3 * cc + 1], u[3 * cc + 2]));
    if (cc / 2 >= leg_first && cc / 2 <= leg_last) t = t + e;
  }
  t = imass * (t - cross(dcom, Ftot));
  col[0] = t.x; col[1] = t.y; col[2] = t.z;
}
template <int NJ>
__device__ __forceinline__ void lq_x_column(const double* __restrict__ bs, const double* __restrict__ u, int c, double* col) {
  using BD = BaseDims<NJ>;
#pragma unroll
  for (int i = 0; i < 9; ++i) col[i] = 0.0;
  if (c < 3) {   // (no dynamic register-array index: it would push col[] into local memory)
#pragma unroll
    for (int i = 0; i < 3; ++i) col[3 + i] = (c == i) ? 1.0 : 0.0;
  } else if (c < 6) {
    const double* A22i = bs + BD::B_A22I; const double* A12 = bs + BD::B_A12; const int cc = c - 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) { col[3 + r] = -(A12[3 * r] * A22i[cc] + A12[3 * r + 1] * A22i[3 + cc] + A12[3 * r + 2] * A22i[6 + cc]); col[6 + r] = c_model.total_mass * A22i[3 * r + cc]; }
  } else lq_dq_column<NJ>(bs, u, c, col);
}
// column l of d f / d qd_j (rows 6..11)
template <int NJ>
__device__ __forceinline__ void lq_bj_column(const double* __restrict__ bs, int l, double* col) {
  using BD = BaseDims<NJ>;
  const double imass = 1.0 / c_model.total_mass;
  const double* J = bs + BD::B_J + BD::JS * l; const double* A22i = bs + BD::B_A22I; const double* A12 = bs + BD::B_A12;
  const v3 n = ld3(J + BD::J_AA), p = ld3(J + BD::J_AL);
  const v3 e = mk(A22i[0] * n.x + A22i[1] * n.y + A22i[2] * n.z, A22i[3] * n.x + A22i[4] * n.y + A22i[5] * n.z, A22i[6] * n.x + A22i[7] * n.y + A22i[8] * n.z);
  const v3 lv = imass * (p - mk(A12[0] * e.x + A12[1] * e.y + A12[2] * e.z, A12[3] * e.x + A12[4] * e.y + A12[5] * e.z, A12[6] * e.x + A12[7] * e.y + A12[8] * e.z));
  col[0] = -lv.x; col[1] = -lv.y; col[2] = -lv.z; col[3] = -e.x; col[4] = -e.y; col[5] = -e.z;
}
// column c (force component) of d f / d F (rows 3..5): column (c % 3) of skew(p_i - com) / m
template <int NJ>
__device__ __forceinline__ void lq_bf_column(const double* __restrict__ bs, int c, double* col) {
  using BD = BaseDims<NJ>;
  const double imass = 1.0 / c_model.total_mass;
  const v3 r = imass * (ld3(bs + BD::B_PC + 3 * (c / 3)) - ld3(bs + BD::B_COM));
  const int a = c % 3;
  col[0] = a == 0 ? 0.0 : (a == 1 ? -r.z : r.y);
  col[1] = a == 0 ? r.z : (a == 1 ? 0.0 : -r.x);
  col[2] = a == 0 ? -r.y : (a == 1 ? r.x : 0.0);
}

// Column pass of one stage (lane = column): analytic d f / d x, d f / d u from the two base records b1, b2 (Heun evaluations), RK2 sensitivities,
// cost gradient, soft friction-cone barrier, compressed constraint rows -> compact LQ record `rec`.  xs/us/xns/xrs: x_k, u_k, x_{k+1}, x_ref.
template <int NJ, bool RAW>
__device__ __forceinline__ void lq_stage_columns(const Dev& d, size_t nb, int k, double* __restrict__ rec, const double* __restrict__ b1, const double* __restrict__ b2,
                                                 const double* xs, const double* us, const double* xns, const double* xrs, double (*sA2w)[Dims<NJ>::NXA + 1], const double* __restrict__ mt, int lane) {
  using D = Dims<NJ>; using BD = BaseDims<NJ>;
  constexpr int NX = D::NX, NU = D::NU, NXA = D::NXA, NL = D::NL;
  const DevModel& M = c_model;
  // per-stage scalars staged by the kernel prologue (mt = [zref 4 | dt | mode | event flag]): no global load inside the column pass
  const double dt = mt[4];
  const int mode = (int)mt[5];
  const double hdt = 0.5 * dt, imass = 1.0 / M.total_mass;
  // ---- Jacobian columns.  Lanes 0..15 work on the first Heun evaluation (b1), lanes 16..31 on the second (b2), with the same instructions:
  //   heavy pass : half-lane cl < NH = NXA - 6 -> state column 6 + cl (base Euler angles, leg joints), lq_dq_column
  //   cheap pass : the angular-momentum columns 3..5 (lanes NH..NH+2 for b1, lanes 24..26 for b2); columns 0..2 are constant
  // Everything below addresses the state columns through xcl = xc_of(lane): lanes 0..NH-1 -> columns 6.., NH..NH+2 -> 3..5, NH+3..NXA-1 -> 0..2,
  // so the columns of the first evaluation (a1) are already in the lane that consumes them; the second evaluation's go through sA2w.
  constexpr int NH = NXA - 6;
  static_assert(NH <= 16 && NH + 6 <= 32, "half-warp split of the Jacobian columns");
  const int half = lane >> 4, cl = lane & 15;
  const double* __restrict__ bsel = half ? b2 : b1;
  const int xcl = lane < NH ? 6 + lane : (lane < NH + 3 ? 3 + (lane - NH) : lane - (NH + 3));   // state column (active-x index) of this lane, lane < NXA
  // All of this is straight-line code executed by every lane (idle lanes work on a clamped column and their results are masked): one large basic
  // block lets the compiler interleave the independent chains, which is what hides latency at two warps per scheduler.
  double a1[9], bj1[6], bj2[6], bf1[3], bf2[3];
  {
    double col[9];
    lq_dq_column<NJ>(bsel, us, cl < NH ? 6 + cl : 6, col);
    const bool own1 = half == 0 && cl < NH, own2 = half == 1 && cl < NH;
#pragma unroll
    for (int r = 0; r < 9; ++r) a1[r] = own1 ? col[r] : 0.0;
    if (own2) {
#pragma unroll
      for (int r = 0; r < 9; ++r) sA2w[r][6 + cl] = col[r];
    }
  }
  {
    // cheap columns: angular momentum 3..5 (lanes NH..NH+2 for b1, lanes 24..26 for b2): -A12 A22i e_c / m A22i e_c; linear momentum 0..2: identity
    const bool c1 = lane >= NH && lane < NH + 3, c2 = lane >= 24 && lane < 27, c3 = lane >= NH + 3 && lane < NXA;
    const int cc = c1 ? lane - NH : (c2 ? lane - 24 : 0), ci = c3 ? lane - (NH + 3) : 0;
    const double* bq = c2 ? b2 : b1;
    const double* A22i = bq + BD::B_A22I; const double* A12 = bq + BD::B_A12;
    const double i0 = A22i[cc], i1 = A22i[3 + cc], i2 = A22i[6 + cc];
    double col[6];
#pragma unroll
    for (int r = 0; r < 3; ++r) { col[r] = -(A12[3 * r] * i0 + A12[3 * r + 1] * i1 + A12[3 * r + 2] * i2); col[3 + r] = M.total_mass * (r == 0 ? i0 : (r == 1 ? i1 : i2)); }
#pragma unroll
    for (int r = 0; r < 6; ++r) a1[3 + r] = c1 ? col[r] : a1[3 + r];
#pragma unroll
    for (int i = 0; i < 3; ++i) a1[3 + i] = (c3 && ci == i) ? 1.0 : a1[3 + i];
    if (c2) {
#pragma unroll
      for (int r = 0; r < 9; ++r) sA2w[r][3 + cc] = r < 3 ? 0.0 : col[r - 3];
    }
    if (c3) {
#pragma unroll
      for (int r = 0; r < 9; ++r) sA2w[r][ci] = (r == 3 + ci) ? 1.0 : 0.0;
    }
  }
  {
    double bjv[6], bfv[3];
    lq_bj_column<NJ>(bsel, cl < NJ ? cl : 0, bjv);
    lq_bf_column<NJ>(bsel, cl < 12 ? cl : 0, bfv);
#pragma unroll
    for (int i = 0; i < 6; ++i) { bjv[i] = cl < NJ ? bjv[i] : 0.0; bj1[i] = bjv[i]; bj2[i] = __shfl_sync(0xffffffffu, bjv[i], (lane + 16) & 31); }
#pragma unroll
    for (int i = 0; i < 3; ++i) { bfv[i] = cl < 12 ? bfv[i] : 0.0; bf1[i] = bfv[i]; bf2[i] = __shfl_sync(0xffffffffu, bfv[i], (lane + 16) & 31); }
  }
  __syncwarp();
  // ---- dynamics: b, (A_d - I), B_d   [UPSTREAM SensitivityIntegrator RK2]
  double pdyn;
  {
    const int li = lane < NX ? lane : 0;
    const double bi = xs[li] + hdt * (b1[BD::B_F + li] + b2[BD::B_F + li]) - xns[li];
    if (lane < NX) rec[D::R_B + lane] = bi;
    pdyn = lane < NX ? bi * bi : 0.0;
  }
  for (int o = 16; o > 0; o >>= 1) pdyn += __shfl_xor_sync(0xffffffffu, pdyn, o);
  {
    const int xc_ = lane < NXA ? xcl : 0, fl = lane < 12 ? lane : 0, jl = lane < NJ ? lane : 0, a = fl % 3;
    double vA[9], vF[9], vJ[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const double a3 = sA2w[r][3], a4 = sA2w[r][4], a5 = sA2w[r][5], a6 = sA2w[r][6], a7 = sA2w[r][7], a8 = sA2w[r][8];
      const double sA = (a3 * a1[0] + a6 * a1[6]) + (a4 * a1[1] + a7 * a1[7]) + (a5 * a1[2] + a8 * a1[8]);
      vA[r] = hdt * (a1[r] + sA2w[r][xc_] + dt * sA);
      const double sF = sA2w[r][a] * imass + a3 * bf1[0] + a4 * bf1[1] + a5 * bf1[2];
      vF[r] = hdt * ((r < 3 ? (bf1[r < 3 ? r : 0] + bf2[r < 3 ? r : 0]) : 0.0) + dt * sF);
      const double sJ = sA2w[r][9 + jl] + a6 * bj1[3] + a7 * bj1[4] + a8 * bj1[5];
      vJ[r] = hdt * ((r >= 3 ? (bj1[r >= 3 ? r - 3 : 0] + bj2[r >= 3 ? r - 3 : 0]) : 0.0) + dt * sJ);
    }
    if (lane < NXA) {
#pragma unroll
      for (int r = 0; r < 9; ++r) rec[D::R_AD + r * NXA + xcl] = vA[r];
    }
    if (lane < 12) {
#pragma unroll
      for (int r = 0; r < 9; ++r) rec[D::R_BD + r * NU + lane] = vF[r];
    }
    if (lane < NJ) {
#pragma unroll
      for (int r = 0; r < 9; ++r) rec[D::R_BD + r * NU + 12 + lane] = vJ[r];
    }
  }
  // ---- cost gradient / barrier blocks (one contact per lane 0..3, one joint per lane for the joint part)
  const bool st0 = leg_in_stance(mode, 0), st1 = leg_in_stance(mode, 1);
  const int nst = 2 * (int(st0) + int(st1));
  const double fznom = nst > 0 ? M.total_mass * 9.81 * (nst == 2 ? 0.5 : 0.25) : 0.0;
  double cpart = 0.0;   // this lane's share of the stage cost value (tracking cost + barrier), summed over the warp below
  if (lane < NX) { const double dq_ = xs[lane] - xrs[lane]; rec[D::R_Q + lane] = dt * M.Qdiag[lane] * dq_; cpart = 0.5 * M.Qdiag[lane] * dq_ * dq_; }
  double shift = 0.0;
  if (lane < NCON) {
    const int c = lane;
    const bool st = (c / 2 == 0) ? st0 : st1;
    double r3[3] = {M.Rforce[3 * c] * us[3 * c], M.Rforce[3 * c + 1] * us[3 * c + 1], M.Rforce[3 * c + 2] * (us[3 * c + 2] - (st ? fznom : 0.0))};
    cpart += 0.5 * (r3[0] * us[3 * c] + r3[1] * us[3 * c + 1] + r3[2] * (us[3 * c + 2] - (st ? fznom : 0.0)));
    double hb[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (st) {
      const double fx = us[3 * c], fy = us[3 * c + 1], fz = us[3 * c + 2];
      const double ts = fx * fx + fy * fy + M.fr_reg, itn = rsqrt(ts), tn = ts * itn, it32 = itn * itn * itn;
      const double h = M.mu_f * (fz + M.fr_grip) - tn;
      double p, dp, ddp; barrier_penalty(h, p, dp, ddp);
      const double g0 = -fx * itn, g1 = -fy * itn, g2 = M.mu_f;
      const double H00 = -(fy * fy + M.fr_reg) * it32, H01 = fx * fy * it32, H11 = -(fx * fx + M.fr_reg) * it32;
      r3[0] += dp * g0; r3[1] += dp * g1; r3[2] += dp * g2;
      hb[0] = ddp * g0 * g0 + dp * H00; hb[1] = ddp * g0 * g1 + dp * H01; hb[2] = ddp * g0 * g2;
      hb[3] = ddp * g1 * g1 + dp * H11; hb[4] = ddp * g1 * g2; hb[5] = ddp * g2 * g2;
      shift = -dp * M.fr_shift;
      cpart += p;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { rec[D::R_R + 3 * c + a] = dt * r3[a]; rec[D::R_FO + 3 * c + a] = us[3 * c + a]; }
#pragma unroll
    for (int a = 0; a < 6; ++a) rec[D::R_HB + 6 * c + a] = dt * hb[a];
  }
  for (int o = 2; o > 0; o >>= 1) shift += __shfl_xor_sync(0xffffffffu, shift, o);   // lanes 0..3
  shift = __shfl_sync(0xffffffffu, shift, 0);
  if (lane < NJ) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) s += M.Rjoint[lane * NJ + j] * us[12 + j];
    rec[D::R_R + 12 + lane] = dt * s;
    cpart += 0.5 * us[12 + lane] * s;
  }
  const double pcost = dt * warp_sum(cpart);
  // ---- contact velocity Jacobians of the first evaluation and the compressed constraint rows
  v3 jx[NCON], ju[NCON];
  {
    // The "direct" term of d v_c / d q_k (the column's own joint moves the contact) is the same straight-line code for every lane: its operands
    // come through one clamped pointer set per lane (base Euler-angle columns 6..8: base axes / partial base twists; joint columns: the joint's
    // record), chosen once per stage, and the result is selected in -- no lane-dependent branches inside the contact loop.
    const v3 pb = ld3(b1 + BD::B_PB);
    const bool xb_ = xcl >= 6 && xcl < 9 && lane < NXA, xj_ = xcl >= 9 && lane < NXA;
    const int jq = xj_ ? xcl - 9 : 0, kq = xb_ ? xcl - 6 : 0;
    const double* Jq = b1 + BD::B_J + BD::JS * jq;
    const v3 ak = ld3(xb_ ? b1 + BD::B_BAX + 3 * kq : Jq + BD::J_A), ok = ld3(xb_ ? b1 + BD::B_PB : Jq + BD::J_O);
    const v3 wk = ld3(xb_ ? b1 + BD::B_WE + 3 * (kq + 1) : Jq + BD::J_W), vk = ld3(xb_ ? b1 + BD::B_VE + 3 * (kq + 1) : Jq + BD::J_V);
    const int xleg = jq / NL;
    const int ju_ = lane < NJ ? lane : 0, uleg = ju_ / NL;
    const double* Ju = b1 + BD::B_J + BD::JS * ju_;
    const v3 aj = ld3(Ju + BD::J_A), oj = ld3(Ju + BD::J_O);
    v3 bax3[3];
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) bax3[kk] = ld3(b1 + BD::B_BAX + 3 * kk);
#pragma unroll
    for (int c = 0; c < NCON; ++c) {
      const int leg = c / 2;
      const v3 p = ld3(b1 + BD::B_PC + 3 * c), vcp = ld3(b1 + BD::B_VC + 3 * c);
      v3 Jb[3];
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) Jb[kk] = cross(bax3[kk], p - pb);
      {
        v3 t = mk(a1[3], a1[4], a1[5]) + a1[6] * Jb[0] + a1[7] * Jb[1] + a1[8] * Jb[2];   // a1 = 0 on lanes >= NXA
        const v3 uw = vcp - (cross(wk, p) + vk), apo = cross(ak, p - ok);   // apo = d p / d q_k
        v3 e = cross(ak, uw) + cross(wk, apo);
        e.z += M.gain * apo.z;   // positionErrorGain (BipedalRobotInterface.cpp:350-359, BipedalRobotPreComputation.cpp:71-80): the z rows also see gain * p_z
        const bool direct = xb_ || (xj_ && xleg == leg);
        jx[c] = direct ? t + e : t;
      }
      {
        v3 t = mk(bj1[0], bj1[1], bj1[2]) + bj1[3] * Jb[0] + bj1[4] * Jb[1] + bj1[5] * Jb[2];
        const v3 e = cross(aj, p - oj);
        ju[c] = (lane < NJ && uleg == leg) ? t + e : t;
      }
    }
  }
  int nrows = 0; double peq = 0.0;
  if constexpr (RAW) {
    // raw rows in upstream's stacking order (BipedalRobotInterface.cpp:187-191): contact 0..3; closed: the 3 zero-velocity rows
    // (ZeroVelocityConstraintCppAd.cpp:58-77), open: the normal-velocity row (NormalVelocityConstraintCppAd.cpp:59-84; its 3 zero-force rows
    // are identity rows on the force columns and handled structurally by k_project)
#pragma unroll
    for (int c = 0; c < NCON; ++c) {
      const bool st = (c / 2 == 0) ? st0 : st1;
      const v3 vcc = ld3(b1 + BD::B_VC + 3 * c);
      const double pz = b1[BD::B_PC + 3 * c + 2];   // contact height (terrain height 0, SwitchedModelReferenceManager.cpp:67)
      if (st) {
        const double ez = vcc.z + M.gain * pz;
        peq += vcc.x * vcc.x + vcc.y * vcc.y + ez * ez;
        if (lane < NXA) { rec[D::R_CV + (nrows + 0) * NXA + xcl] = jx[c].x; rec[D::R_CV + (nrows + 1) * NXA + xcl] = jx[c].y; rec[D::R_CV + (nrows + 2) * NXA + xcl] = jx[c].z; }
        if (lane < NJ) { rec[D::R_DV + (nrows + 0) * NJ + lane] = ju[c].x; rec[D::R_DV + (nrows + 1) * NJ + lane] = ju[c].y; rec[D::R_DV + (nrows + 2) * NJ + lane] = ju[c].z; }
        if (lane == 0) { rec[D::R_EV + nrows] = vcc.x; rec[D::R_EV + nrows + 1] = vcc.y; rec[D::R_EV + nrows + 2] = ez; }
        nrows += 3;
      } else {
        const double ev = vcc.z - mt[c / 2] + M.gain * (pz - mt[2 + c / 2]);
        if (lane < NXA) rec[D::R_CV + nrows * NXA + xcl] = jx[c].z;
        if (lane < NJ) rec[D::R_DV + nrows * NJ + lane] = ju[c].z;
        if (lane == 0) rec[D::R_EV + nrows] = ev;
        peq += ev * ev + us[3 * c] * us[3 * c] + us[3 * c + 1] * us[3 * c + 1] + us[3 * c + 2] * us[3 * c + 2];
        ++nrows;
      }
    }
  } else {
  const double is2 = 0.7071067811865476;
#pragma unroll
  for (int leg = 0; leg < 2; ++leg) {
    const int ca = 2 * leg, cb = 2 * leg + 1;
    const bool st = leg == 0 ? st0 : st1;
    const v3 va = ld3(b1 + BD::B_VC + 3 * ca), vb = ld3(b1 + BD::B_VC + 3 * cb);
    if (st) {
      peq += dot(va, va) + dot(vb, vb);
      v3 r = ld3(b1 + BD::B_PC + 3 * ca) - ld3(b1 + BD::B_PC + 3 * cb);
      r = rsqrt(dot(r, r)) * r;
      const double ax = fabs(r.x), ay = fabs(r.y), az = fabs(r.z);
      const v3 e = (ax <= ay && ax <= az) ? mk(1.0, 0.0, 0.0) : ((ay <= az) ? mk(0.0, 1.0, 0.0) : mk(0.0, 0.0, 1.0));
      v3 n1 = cross(r, e); n1 = rsqrt(dot(n1, n1)) * n1;
      const v3 n2 = cross(r, n1);
      const v3 sx_ = is2 * (jx[ca] + jx[cb]), dx_ = is2 * (jx[ca] - jx[cb]), su_ = is2 * (ju[ca] + ju[cb]), du_ = is2 * (ju[ca] - ju[cb]);
      if (lane < NXA) {
        rec[D::R_CV + (nrows + 0) * NXA + xcl] = sx_.x; rec[D::R_CV + (nrows + 1) * NXA + xcl] = sx_.y; rec[D::R_CV + (nrows + 2) * NXA + xcl] = sx_.z;
        rec[D::R_CV + (nrows + 3) * NXA + xcl] = dot(n1, dx_); rec[D::R_CV + (nrows + 4) * NXA + xcl] = dot(n2, dx_);
      }
      if (lane < NJ) {
        rec[D::R_DV + (nrows + 0) * NJ + lane] = su_.x; rec[D::R_DV + (nrows + 1) * NJ + lane] = su_.y; rec[D::R_DV + (nrows + 2) * NJ + lane] = su_.z;
        rec[D::R_DV + (nrows + 3) * NJ + lane] = dot(n1, du_); rec[D::R_DV + (nrows + 4) * NJ + lane] = dot(n2, du_);
      }
      if (lane == 0) {
        const v3 sv = is2 * (va + vb), dv = is2 * (va - vb);
        rec[D::R_EV + nrows] = sv.x; rec[D::R_EV + nrows + 1] = sv.y; rec[D::R_EV + nrows + 2] = sv.z; rec[D::R_EV + nrows + 3] = dot(n1, dv); rec[D::R_EV + nrows + 4] = dot(n2, dv);
      }
      nrows += 5;
    } else {
      const double zr = mt[leg];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int c0 = t == 0 ? ca : cb;
        if (lane < NXA) rec[D::R_CV + nrows * NXA + xcl] = jx[c0].z;
        if (lane < NJ) rec[D::R_DV + nrows * NJ + lane] = ju[c0].z;
        const double ev = (t == 0 ? va.z : vb.z) - zr;
        if (lane == 0) rec[D::R_EV + nrows] = ev;
        peq += ev * ev + us[3 * c0] * us[3 * c0] + us[3 * c0 + 1] * us[3 * c0 + 1] + us[3 * c0 + 2] * us[3 * c0 + 2];
        ++nrows;
      }
    }
  }
  }
  if (lane == 0) {
    double* misc = rec + D::R_MISC;
    misc[D::M_DT] = dt; misc[D::M_DQ] = dt * shift; misc[D::M_DR] = dt * shift; misc[D::M_MODE] = (double)mode; misc[D::M_NROWS] = (double)nrows;
    misc[D::M_TYPE] = 0.0; misc[D::M_PCOST] = pcost; misc[D::M_PDYN] = dt * pdyn; misc[D::M_PEQ] = dt * peq;
  }
}

// Packed LQ kernel (default): one warp per G consecutive stages of an instance (H1: G = 3 segments of 10 lanes, G1: G = 2 segments of 16).
// The base pass (lane inside the segment = leg joint) runs for the G stages at once, so G NJ of 32 lanes are busy instead of NJ; the
// column pass (lane = column) then handles the stages one after the other.  A segment whose stage is an event node or beyond the horizon
// mirrors the inputs of a stage that needs the model (results discarded).
template <int NJ>
struct LqPackSmem {
  static constexpr int BASE = BaseDims<NJ>::BASE, NXA = Dims<NJ>::NXA, WPB = 4;
  static constexpr int SEG = (NJ <= 10) ? 10 : 16, G = 32 / SEG;
  double jc[NJ][28];
  double base[WPB][G][2 * BASE];
  double A2[WPB][9][NXA + 1];
  double xu[WPB][G][4 * 24];   // per stage: x, u, xnext, xref
  double mt[WPB][G][8];        // per stage: zref (4), dt, mode, event flag
};
template <int NJ, bool RAW>
__global__ void __launch_bounds__(128, LQ_PAIR_BLOCKS) k_lq_pack(Dev d) {
  using D = Dims<NJ>; using BD = BaseDims<NJ>; using SM = LqPackSmem<NJ>;
  constexpr int NX = D::NX, NU = D::NU, WPB = SM::WPB, BASE = BD::BASE, SEG = SM::SEG, G = SM::G;
  static_assert(G * 24 <= 9 * (D::NXA + 1), "the second RK2 evaluation points are staged in the A2 buffer");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // broadcast: lets the compiler treat the warp index as warp-uniform
  for (int i = threadIdx.x; i < NJ * 28; i += 128) (&sm.jc[0][0])[i] = d.jc[i];
  __syncthreads();
  const int NP = (d.NS + G - 1) / G;
  const int gw = blockIdx.x * WPB + warp;
  const int b = gw / NP, k0 = G * (gw % NP);
  if (b >= d.B) return;
  // ---- one round of global loads: everything the G stages need is requested before the first dependent use (rows are clamped into the
  //      instance's node slots, so no load waits for n_nodes / node_ev); the per-stage scalars travel through shared memory as well
  const size_t nb = (size_t)b * d.NS;
  const int Nn = d.n_nodes[b];
  int evv[G];
#pragma unroll
  for (int s = 0; s < G; ++s) {
    const int kc = (k0 + s < d.NS - 1) ? k0 + s : d.NS - 2;
    double* xs = sm.xu[warp][s];
    evv[s] = d.node_ev[nb + kc];
    const double v0 = lane < NX ? d.s_x[(nb + kc) * NX + lane] : 0.0, v1 = lane < NX ? d.s_x[(nb + kc + 1) * NX + lane] : 0.0;
    const double v2 = lane < NX ? d.xref[(nb + kc) * NX + lane] : 0.0, v3_ = lane < NU ? d.s_u[(nb + kc) * NU + lane] : 0.0;
    double v4 = 0.0;
    if (lane < 4) v4 = d.zref[(nb + kc) * 4 + lane]; else if (lane == 4) v4 = d.st_dt[nb + kc]; else if (lane == 5) v4 = (double)d.st_mode[nb + kc];
    if (lane < NX) { xs[lane] = v0; xs[48 + lane] = v1; xs[72 + lane] = v2; }
    if (lane < NU) xs[24 + lane] = v3_;
    if (lane < 6) sm.mt[warp][s][lane] = v4;
  }
  const int N = Nn - 1;
  if (k0 >= N) return;
  bool has[G], ev[G], comp[G];   // stage exists / is an event node / needs the model
  int first_comp = -1;
#pragma unroll
  for (int s = 0; s < G; ++s) {
    has[s] = k0 + s < N;
    ev[s] = has[s] && evv[s] == 1;
    comp[s] = has[s] && !ev[s];
    if (comp[s] && first_comp < 0) first_comp = s;
    if (lane == 6) sm.mt[warp][s][6] = ev[s] ? 1.0 : 0.0;
  }
  __syncwarp();
  if (first_comp >= 0) {
    const int h = lane / SEG;          // segment of this lane (lanes beyond the last complete segment tag along with segment 0's data)
    int ms = first_comp;               // stage whose inputs this segment evaluates
#pragma unroll
    for (int s = 0; s < G; ++s) if (h == s && comp[s]) ms = s;
    const int hs_ = h < G ? h : 0;
    const double* xh = sm.xu[warp][ms]; const double* uh = xh + 24;
    double* bh = sm.base[warp][hs_];
    const int jl = lane % SEG;
    const double* jc = sm.jc[jl < NJ ? jl : 0];
    double* x2 = &sm.A2[warp][0][0];   // scratch for the second RK2 evaluation points (A2 is filled later): G x 24 doubles
    // the two Heun evaluations share one copy of the (large) base-pass code: the kernel is instruction-cache bound otherwise
#pragma unroll 1
    for (int ev_ = 0; ev_ < 2; ++ev_) {
      warp_model_base<NJ, SEG>(ev_ == 0 ? xh : x2 + 24 * ms, uh, bh + ev_ * BASE, lane, jc);
      __syncwarp();
      if (ev_ == 0) {
#pragma unroll
        for (int s = 0; s < G; ++s)
          if (lane < NX) x2[24 * s + lane] = comp[s] ? sm.xu[warp][s][lane] + sm.mt[warp][s][4] * sm.base[warp][s][BD::B_F + lane] : 0.0;
        __syncwarp();
      }
    }
  }
#pragma unroll 1
  for (int s = 0; s < G; ++s) {
    if (k0 + s >= N) break;
    const int k = k0 + s;
    double* __restrict__ rec = d.lq + (nb + k) * D::REC;
    const double* xs = sm.xu[warp][s];
    if (sm.mt[warp][s][6] != 0.0) {   // [UPSTREAM] setupEventNode
      double sq = 0.0;
      if (lane < NX) { const double bi = xs[lane] - xs[48 + lane]; rec[D::R_B + lane] = bi; sq = bi * bi; }
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0) {
        rec[D::R_MISC + D::M_TYPE] = 1.0; rec[D::R_MISC + D::M_DT] = 0.0; rec[D::R_MISC + D::M_MODE] = -1.0;
        rec[D::R_MISC + D::M_PCOST] = 0.0; rec[D::R_MISC + D::M_PDYN] = sq; rec[D::R_MISC + D::M_PEQ] = 0.0;
      }
      continue;
    }
    lq_stage_columns<NJ, RAW>(d, nb, k, rec, sm.base[warp][s], sm.base[warp][s] + BASE, xs, xs + 24, xs + 48, xs + 72, sm.A2[warp], sm.mt[warp][s], lane);
    __syncwarp();   // A2 is reused by the next stage
  }
}

}  // namespace bmpc
