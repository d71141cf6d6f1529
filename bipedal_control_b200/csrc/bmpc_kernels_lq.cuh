// K1: LQ approximation (packed warp-cooperative kernel + the earlier variants kept for cross-checking)
// (part of bmpc_kernels.cuh: include that header, not this file)
#pragma once

namespace bmpc {

// ------------------------------------------------------------------------------------------------ K1: LQ approximation, one thread per (instance, stage)
template <int NJ>
__global__ void __launch_bounds__(64, LQ_MIN_BLOCKS) k_lq(Dev d) {
  using D = Dims<NJ>;
  constexpr int NX = D::NX, NU = D::NU, NXA = D::NXA;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = gid / d.NS, k = gid % d.NS;
  if (b >= d.B) return;
  const int N = d.n_nodes[b] - 1;
  if (k >= N) return;
  const size_t nb = (size_t)b * d.NS;
  const double* xg = d.s_x + (nb + k) * NX; const double* ug = d.s_u + (nb + k) * NU; const double* xng = xg + NX;
  double* rec = d.lq + (nb + k) * D::REC;
  double x[NX], u[NU];
#pragma unroll 1
  for (int i = 0; i < NX; ++i) x[i] = xg[i];
  if (d.node_ev[nb + k] == 1) {   // [UPSTREAM] setupEventNode: identity jump map, no input
    double s = 0.0;
    for (int i = 0; i < NX; ++i) { const double bi = x[i] - xng[i]; rec[D::R_B + i] = bi; s += bi * bi; }
    rec[D::R_MISC + D::M_TYPE] = 1.0; rec[D::R_MISC + D::M_DT] = 0.0; rec[D::R_MISC + D::M_MODE] = -1.0;
    rec[D::R_MISC + D::M_PCOST] = 0.0; rec[D::R_MISC + D::M_PDYN] = s; rec[D::R_MISC + D::M_PEQ] = 0.0;
    return;
  }
#pragma unroll 1
  for (int i = 0; i < NU; ++i) u[i] = ug[i];
  const double dt = d.st_dt[nb + k];
  const int mode = d.st_mode[nb + k];
  const DevModel& M = c_model;
  // ---- dynamics: Heun / RK2 with sensitivities  [UPSTREAM SensitivityIntegrator RK2]
  ModelEval<NJ> E1; ContactJac<NJ> CJ;
  model_eval<NJ, 2>(x, u, E1, &CJ);
  double x2[NX];
#pragma unroll 1
  for (int i = 0; i < NX; ++i) x2[i] = x[i] + dt * E1.f[i];
  ModelEval<NJ> E2;
  model_eval<NJ, 1>(x2, u, E2, nullptr);
  const double hdt = 0.5 * dt, imass = 1.0 / M.total_mass;
  double pdyn = 0.0;
#pragma unroll 1
  for (int i = 0; i < NX; ++i) { const double bi = x[i] + hdt * (E1.f[i] + E2.f[i]) - xng[i]; rec[D::R_B + i] = bi; pdyn += bi * bi; }
  // (A_d - I) rows 3..11, active columns: dt/2 (A1 + A2 + dt A2 A1); A1 rows that matter: states 3,4,5 (block rows 0..2) and 9,10,11 (block rows 6..8)
  for (int r = 0; r < 9; ++r)
    for (int c = 0; c < NXA; ++c) {
      double s = 0.0;
#pragma unroll 1
      for (int t = 0; t < 3; ++t) s += E2.Ac[r][3 + t] * E1.Ac[t][c] + E2.Ac[r][6 + t] * E1.Ac[6 + t][c];
      rec[D::R_AD + r * NXA + c] = hdt * (E1.Ac[r][c] + E2.Ac[r][c] + dt * s);
    }
  // B_d rows 3..11
  for (int r = 0; r < 9; ++r) {
    for (int c = 0; c < 12; ++c) {   // force columns
      const int a = c % 3;
      double s = E2.Ac[r][a] * imass;
#pragma unroll 1
      for (int t = 0; t < 3; ++t) s += E2.Ac[r][3 + t] * E1.Bf[t][c];
      const double b12 = r < 3 ? (E1.Bf[r][c] + E2.Bf[r][c]) : 0.0;
      rec[D::R_BD + r * NU + c] = hdt * (b12 + dt * s);
    }
    for (int l = 0; l < NJ; ++l) {   // joint-velocity columns
      double s = E2.Ac[r][9 + l];
#pragma unroll 1
      for (int t = 0; t < 3; ++t) s += E2.Ac[r][6 + t] * E1.Bj[3 + t][l];
      const double b12 = r >= 3 ? (E1.Bj[r - 3][l] + E2.Bj[r - 3][l]) : 0.0;
      rec[D::R_BD + r * NU + 12 + l] = hdt * (b12 + dt * s);
    }
  }
  // ---- cost (x dt): tracking cost + soft friction cones
  const double* xr = d.xref + (nb + k) * NX;
  const bool st0 = leg_in_stance(mode, 0), st1 = leg_in_stance(mode, 1);
  const int nst = 2 * (int(st0) + int(st1));
  const double fznom = nst > 0 ? M.total_mass * 9.81 / nst : 0.0;
  for (int i = 0; i < NX; ++i) rec[D::R_Q + i] = dt * M.Qdiag[i] * (x[i] - xr[i]);
  double shift = 0.0;
  for (int c = 0; c < NCON; ++c) {
    const bool st = (c / 2 == 0) ? st0 : st1;
    double r3[3] = {M.Rforce[3 * c] * u[3 * c], M.Rforce[3 * c + 1] * u[3 * c + 1], M.Rforce[3 * c + 2] * (u[3 * c + 2] - (st ? fznom : 0.0))};
    double hb[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (st) {  // constraint/FrictionConeConstraint.cpp:96-166 wrapped by StateInputSoftConstraint + RelaxedBarrierPenalty
      const double fx = u[3 * c], fy = u[3 * c + 1], fz = u[3 * c + 2];
      const double ts = fx * fx + fy * fy + M.fr_reg, tn = sqrt(ts), t32 = tn * ts;
      const double h = M.mu_f * (fz + M.fr_grip) - tn;
      double p, dp, ddp; barrier_penalty(h, p, dp, ddp);
      const double g0 = -fx / tn, g1 = -fy / tn, g2 = M.mu_f;
      const double H00 = -(fy * fy + M.fr_reg) / t32, H01 = fx * fy / t32, H11 = -(fx * fx + M.fr_reg) / t32;
      r3[0] += dp * g0; r3[1] += dp * g1; r3[2] += dp * g2;
      hb[0] = ddp * g0 * g0 + dp * H00; hb[1] = ddp * g0 * g1 + dp * H01; hb[2] = ddp * g0 * g2;
      hb[3] = ddp * g1 * g1 + dp * H11; hb[4] = ddp * g1 * g2; hb[5] = ddp * g2 * g2;
      shift += -dp * M.fr_shift;   // FrictionConeConstraint.cpp:192-206: whole uu / xx diagonals
    }
    for (int a = 0; a < 3; ++a) rec[D::R_R + 3 * c + a] = dt * r3[a];
    for (int a = 0; a < 6; ++a) rec[D::R_HB + 6 * c + a] = dt * hb[a];
    for (int a = 0; a < 3; ++a) rec[D::R_FO + 3 * c + a] = u[3 * c + a];
  }
  for (int i = 0; i < NJ; ++i) {
    double s = 0.0;
#pragma unroll 1
    for (int j = 0; j < NJ; ++j) s += M.Rjoint[i * NJ + j] * u[12 + j];
    rec[D::R_R + 12 + i] = dt * s;
  }
  const double pcost = dt * stage_cost_value<NJ>(mode, x, u, xr);
  // ---- equality constraints on the contact velocities (rows compressed per foot: the two sole points of a stance foot give
  //      6 rows of rank 5; the sum / difference rotation below is orthogonal, the dropped row has an identically zero D part,
  //      so the Moore-Penrose solution is unchanged)
  int nrows = 0; double peq = 0.0;
  const double is2 = 0.7071067811865476;
  for (int leg = 0; leg < 2; ++leg) {
    const int ca = 2 * leg, cb = 2 * leg + 1;
    const bool st = leg == 0 ? st0 : st1;
    if (st) {
      const v3 va = E1.vc[ca], vb = E1.vc[cb];
      peq += dot(va, va) + dot(vb, vb);
      for (int i = 0; i < 3; ++i) {   // sum rows
        double* Cr = rec + D::R_CV + (nrows + i) * NXA; double* Dr = rec + D::R_DV + (nrows + i) * NJ;
        for (int c = 0; c < NXA; ++c) Cr[c] = is2 * (CJ.Jx[ca][i][c] + CJ.Jx[cb][i][c]);
        for (int c = 0; c < NJ; ++c) Dr[c] = is2 * (CJ.Ju[ca][i][c] + CJ.Ju[cb][i][c]);
        rec[D::R_EV + nrows + i] = is2 * (comp(va, i) + comp(vb, i));
      }
      nrows += 3;
      v3 r = E1.pc[ca] - E1.pc[cb];
      r = (1.0 / sqrt(dot(r, r))) * r;
      const double ax = fabs(r.x), ay = fabs(r.y), az = fabs(r.z);
      const v3 e = (ax <= ay && ax <= az) ? mk(1.0, 0.0, 0.0) : ((ay <= az) ? mk(0.0, 1.0, 0.0) : mk(0.0, 0.0, 1.0));
      v3 n1 = cross(r, e); n1 = (1.0 / sqrt(dot(n1, n1))) * n1;
      const v3 n2 = cross(r, n1);
      for (int t = 0; t < 2; ++t) {   // difference rows projected on the plane normal to the foot axis
        const v3 nn = t == 0 ? n1 : n2;
        double* Cr = rec + D::R_CV + (nrows + t) * NXA; double* Dr = rec + D::R_DV + (nrows + t) * NJ;
        for (int c = 0; c < NXA; ++c) Cr[c] = is2 * (nn.x * (CJ.Jx[ca][0][c] - CJ.Jx[cb][0][c]) + nn.y * (CJ.Jx[ca][1][c] - CJ.Jx[cb][1][c]) + nn.z * (CJ.Jx[ca][2][c] - CJ.Jx[cb][2][c]));
        for (int c = 0; c < NJ; ++c) Dr[c] = is2 * (nn.x * (CJ.Ju[ca][0][c] - CJ.Ju[cb][0][c]) + nn.y * (CJ.Ju[ca][1][c] - CJ.Ju[cb][1][c]) + nn.z * (CJ.Ju[ca][2][c] - CJ.Ju[cb][2][c]));
        rec[D::R_EV + nrows + t] = is2 * dot(nn, va - vb);
      }
      nrows += 2;
    } else {
      const double zr = d.zref[(nb + k) * 2 + leg];
      for (int t = 0; t < 2; ++t) {   // normal velocity rows (NormalVelocityConstraintCppAd.cpp:59-84, BipedalRobotPreComputation.cpp:71-80)
        const int c0 = t == 0 ? ca : cb;
        double* Cr = rec + D::R_CV + nrows * NXA; double* Dr = rec + D::R_DV + nrows * NJ;
        for (int c = 0; c < NXA; ++c) Cr[c] = CJ.Jx[c0][2][c];
        for (int c = 0; c < NJ; ++c) Dr[c] = CJ.Ju[c0][2][c];
        const double ev = E1.vc[c0].z - zr;
        rec[D::R_EV + nrows] = ev;
        peq += ev * ev + u[3 * c0] * u[3 * c0] + u[3 * c0 + 1] * u[3 * c0 + 1] + u[3 * c0 + 2] * u[3 * c0 + 2];   // + zero-force rows
        ++nrows;
      }
    }
  }
  double* misc = rec + D::R_MISC;
  misc[D::M_DT] = dt; misc[D::M_DQ] = dt * shift; misc[D::M_DR] = dt * shift; misc[D::M_MODE] = (double)mode; misc[D::M_NROWS] = (double)nrows;
  misc[D::M_TYPE] = 0.0; misc[D::M_PCOST] = pcost; misc[D::M_PDYN] = dt * pdyn; misc[D::M_PEQ] = dt * peq;
}

// ------------------------------------------------------------------------------------------------ K1a/K1b: LQ approximation split by parallelism
// K1a k_model_base : one THREAD per stage, values only (FK, composite inertias, CMM, twists, subtree momenta) for both RK2 evaluations
// K1b k_lq_assemble: one WARP per stage, lane = column: analytic Jacobian columns, RK2 sensitivities, cost, constraint rows -> compact LQ record
// (same record as k_lq; k_lq is kept as the single-kernel reference implementation for cross-checks).
template <int NJ>
__global__ void __launch_bounds__(64, 6) k_model_base(Dev d) {
  using D = Dims<NJ>; using BD = BaseDims<NJ>;
  constexpr int NX = D::NX, NU = D::NU;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = gid / d.NS, k = gid % d.NS;
  if (b >= d.B) return;
  const int N = d.n_nodes[b] - 1;
  if (k >= N) return;
  const size_t nb = (size_t)b * d.NS;
  if (d.node_ev[nb + k] == 1) return;
  double x[NX], u[NU];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = d.s_x[(nb + k) * NX + i];
#pragma unroll
  for (int i = 0; i < NU; ++i) u[i] = d.s_u[(nb + k) * NU + i];
  double* base0 = d.base + (nb + k) * (size_t)(2 * BD::BASE);
  model_base<NJ>(x, u, base0);
  const double dt = d.st_dt[nb + k];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] += dt * base0[BD::B_F + i];
  model_base<NJ>(x, u, base0 + BD::BASE);
}

// one column (X index c >= 6) of d f / d x from the base record: rows 3..5 -> col[0..2], rows 6..8 -> col[3..5], rows 9..11 -> col[6..8]
template <int NJ>
__device__ __forceinline__ void lq_dq_column(const double* __restrict__ bs, const double* __restrict__ u, int c, double* col) {
  using BD = BaseDims<NJ>; constexpr int NL = Dims<NJ>::NL;
  const double imass = 1.0 / c_model.total_mass;
  v3 ak, ok, wp, vp, AlinK; SI sub; Mom hsub; int leg_first, leg_last;
  if (c < 9) {
    const int k = c - 6;
    ak = ld3(bs + BD::B_BAX + 3 * k); ok = ld3(bs + BD::B_PB); sub = ld_si(bs + BD::B_TOT);
    hsub.n = ld3(bs + BD::B_HTOT); hsub.p = ld3(bs + BD::B_HTOT + 3);
    wp = ld3(bs + BD::B_WE + 3 * k); vp = ld3(bs + BD::B_VE + 3 * k); AlinK = ld3(bs + BD::B_ALE + 3 * k); leg_first = 0; leg_last = 1;
  } else {
    const int j = c - 9; const double* J = bs + BD::B_J + BD::JS * j;
    ak = ld3(J + BD::J_A); ok = ld3(J + BD::J_O); sub = ld_si(J + BD::J_SI); hsub.n = ld3(J + BD::J_HN); hsub.p = ld3(J + BD::J_HP);
    if (j % NL == 0) { wp = ld3(bs + BD::B_WE + 9); vp = ld3(bs + BD::B_VE + 9); } else { wp = ld3(J - BD::JS + BD::J_W); vp = ld3(J - BD::JS + BD::J_V); }
    AlinK = ld3(J + BD::J_AL); leg_first = leg_last = j / NL;
  }
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
  for (int cc = 0; cc < NCON; ++cc)
    if (cc / 2 >= leg_first && cc / 2 <= leg_last) t = t + cross(cross(ak, ld3(bs + BD::B_PC + 3 * cc) - ok), mk(u[3 * cc], u[3 * cc + 1], u[3 * cc + 2]));
  t = imass * (t - cross(dcom, Ftot));
  col[0] = t.x; col[1] = t.y; col[2] = t.z;
}
template <int NJ>
__device__ __forceinline__ void lq_x_column(const double* __restrict__ bs, const double* __restrict__ u, int c, double* col) {
  using BD = BaseDims<NJ>;
#pragma unroll
  for (int i = 0; i < 9; ++i) col[i] = 0.0;
  if (c < 3) col[3 + c] = 1.0;
  else if (c < 6) {
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
template <int NJ>
__device__ __forceinline__ void lq_stage_columns(const Dev& d, size_t nb, int k, double* __restrict__ rec, const double* __restrict__ b1, const double* __restrict__ b2,
                                                 const double* xs, const double* us, const double* xns, const double* xrs, double (*sA2w)[Dims<NJ>::NXA + 1], int lane) {
  using D = Dims<NJ>; using BD = BaseDims<NJ>;
  constexpr int NX = D::NX, NU = D::NU, NXA = D::NXA, NL = D::NL;
  const DevModel& M = c_model;
  const double dt = d.st_dt[nb + k];
  const int mode = d.st_mode[nb + k];
  const double hdt = 0.5 * dt, imass = 1.0 / M.total_mass;
  // ---- Jacobian columns (lane = column)
  double a1[9], a2[9], bj1[6], bj2[6], bf1[3], bf2[3];
  if (lane < NXA) { lq_x_column<NJ>(b1, us, lane, a1); lq_x_column<NJ>(b2, us, lane, a2); }
  if (lane < NJ) { lq_bj_column<NJ>(b1, lane, bj1); lq_bj_column<NJ>(b2, lane, bj2); }
  if (lane < 12) { lq_bf_column<NJ>(b1, lane, bf1); lq_bf_column<NJ>(b2, lane, bf2); }
  if (lane < NXA) {
#pragma unroll
    for (int r = 0; r < 9; ++r) sA2w[r][lane] = a2[r];
  }
  __syncwarp();
  // ---- dynamics: b, (A_d - I), B_d   [UPSTREAM SensitivityIntegrator RK2]
  double pdyn = 0.0;
  if (lane < NX) { const double bi = xs[lane] + hdt * (b1[BD::B_F + lane] + b2[BD::B_F + lane]) - xns[lane]; rec[D::R_B + lane] = bi; pdyn = bi * bi; }
  for (int o = 16; o > 0; o >>= 1) pdyn += __shfl_xor_sync(0xffffffffu, pdyn, o);
  if (lane < NXA) {
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      double s = 0.0;
#pragma unroll
      for (int t = 0; t < 3; ++t) s += sA2w[r][3 + t] * a1[t] + sA2w[r][6 + t] * a1[6 + t];
      rec[D::R_AD + r * NXA + lane] = hdt * (a1[r] + a2[r] + dt * s);
    }
  }
  if (lane < 12) {
    const int a = lane % 3;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      double s = sA2w[r][a] * imass;
#pragma unroll
      for (int t = 0; t < 3; ++t) s += sA2w[r][3 + t] * bf1[t];
      const double b12 = r < 3 ? (bf1[r] + bf2[r]) : 0.0;
      rec[D::R_BD + r * NU + lane] = hdt * (b12 + dt * s);
    }
  }
  if (lane < NJ) {
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      double s = sA2w[r][9 + lane];
#pragma unroll
      for (int t = 0; t < 3; ++t) s += sA2w[r][6 + t] * bj1[3 + t];
      const double b12 = r >= 3 ? (bj1[r - 3] + bj2[r - 3]) : 0.0;
      rec[D::R_BD + r * NU + 12 + lane] = hdt * (b12 + dt * s);
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
    const v3 pb = ld3(b1 + BD::B_PB);
#pragma unroll
    for (int c = 0; c < NCON; ++c) {
      const int leg = c / 2;
      const v3 p = ld3(b1 + BD::B_PC + 3 * c), vcp = ld3(b1 + BD::B_VC + 3 * c);
      v3 Jb[3];
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) Jb[kk] = cross(ld3(b1 + BD::B_BAX + 3 * kk), p - pb);
      jx[c] = mk(0.0, 0.0, 0.0); ju[c] = mk(0.0, 0.0, 0.0);
      if (lane < NXA) {
        v3 t = mk(a1[3], a1[4], a1[5]) + a1[6] * Jb[0] + a1[7] * Jb[1] + a1[8] * Jb[2];
        bool direct = false; v3 ak, ok, wk, vk;
        if (lane >= 6 && lane < 9) { const int kk = lane - 6; direct = true; ak = ld3(b1 + BD::B_BAX + 3 * kk); ok = pb; wk = ld3(b1 + BD::B_WE + 3 * (kk + 1)); vk = ld3(b1 + BD::B_VE + 3 * (kk + 1)); }
        else if (lane >= 9 && (lane - 9) / NL == leg) { const double* J = b1 + BD::B_J + BD::JS * (lane - 9); direct = true; ak = ld3(J + BD::J_A); ok = ld3(J + BD::J_O); wk = ld3(J + BD::J_W); vk = ld3(J + BD::J_V); }
        if (direct) { const v3 uw = vcp - (cross(wk, p) + vk); t = t + cross(ak, uw) + cross(wk, cross(ak, p - ok)); }
        jx[c] = t;
      }
      if (lane < NJ) {
        v3 t = mk(bj1[0], bj1[1], bj1[2]) + bj1[3] * Jb[0] + bj1[4] * Jb[1] + bj1[5] * Jb[2];
        if (lane / NL == leg) { const double* J = b1 + BD::B_J + BD::JS * lane; t = t + cross(ld3(J + BD::J_A), p - ld3(J + BD::J_O)); }
        ju[c] = t;
      }
    }
  }
  int nrows = 0; double peq = 0.0;
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
        rec[D::R_CV + (nrows + 0) * NXA + lane] = sx_.x; rec[D::R_CV + (nrows + 1) * NXA + lane] = sx_.y; rec[D::R_CV + (nrows + 2) * NXA + lane] = sx_.z;
        rec[D::R_CV + (nrows + 3) * NXA + lane] = dot(n1, dx_); rec[D::R_CV + (nrows + 4) * NXA + lane] = dot(n2, dx_);
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
      const double zr = d.zref[(nb + k) * 2 + leg];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int c0 = t == 0 ? ca : cb;
        if (lane < NXA) rec[D::R_CV + nrows * NXA + lane] = jx[c0].z;
        if (lane < NJ) rec[D::R_DV + nrows * NJ + lane] = ju[c0].z;
        const double ev = (t == 0 ? va.z : vb.z) - zr;
        if (lane == 0) rec[D::R_EV + nrows] = ev;
        peq += ev * ev + us[3 * c0] * us[3 * c0] + us[3 * c0 + 1] * us[3 * c0 + 1] + us[3 * c0 + 2] * us[3 * c0 + 2];
        ++nrows;
      }
    }
  }
  if (lane == 0) {
    double* misc = rec + D::R_MISC;
    misc[D::M_DT] = dt; misc[D::M_DQ] = dt * shift; misc[D::M_DR] = dt * shift; misc[D::M_MODE] = (double)mode; misc[D::M_NROWS] = (double)nrows;
    misc[D::M_TYPE] = 0.0; misc[D::M_PCOST] = pcost; misc[D::M_PDYN] = dt * pdyn; misc[D::M_PEQ] = dt * peq;
  }
}

template <int NJ, bool FUSED>
__global__ void __launch_bounds__(128, FUSED ? LQ_FUSED_BLOCKS : 4) k_lq_assemble(Dev d) {
  using D = Dims<NJ>; using BD = BaseDims<NJ>;
  constexpr int NX = D::NX, NU = D::NU, NXA = D::NXA, NL = D::NL, WPB = 4, BASE = BD::BASE;
  __shared__ double sbase[WPB][2 * BASE];
  __shared__ double sA2[WPB][9][NXA + 1];
  __shared__ double sxu[WPB][4 * 24];   // x, u, xnext, xref
  __shared__ double sjc[FUSED ? NJ : 1][28];   // per-joint model constants (lane-indexed reads of __constant__ memory would serialise)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // broadcast: lets the compiler treat the warp index as warp-uniform
  if (FUSED) {
    for (int i = threadIdx.x; i < NJ * 28; i += 128) (&sjc[0][0])[i] = d.jc[i];   // packed [Rj 9 | pj 3 | axis 3 | mass | com 3 | inertia 9] per joint
    __syncthreads();
  }
  const int gw = blockIdx.x * WPB + warp;
  const int b = gw / d.NS, k = gw % d.NS;
  if (b >= d.B) return;
  const int N = d.n_nodes[b] - 1;
  if (k >= N) return;
  const size_t nb = (size_t)b * d.NS;
  double* __restrict__ rec = d.lq + (nb + k) * D::REC;
  const double* xg = d.s_x + (nb + k) * NX; const double* xng = xg + NX;
  if (d.node_ev[nb + k] == 1) {   // [UPSTREAM] setupEventNode
    double s = 0.0;
    if (lane < NX) { const double bi = xg[lane] - xng[lane]; rec[D::R_B + lane] = bi; s = bi * bi; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      rec[D::R_MISC + D::M_TYPE] = 1.0; rec[D::R_MISC + D::M_DT] = 0.0; rec[D::R_MISC + D::M_MODE] = -1.0;
      rec[D::R_MISC + D::M_PCOST] = 0.0; rec[D::R_MISC + D::M_PDYN] = s; rec[D::R_MISC + D::M_PEQ] = 0.0;
    }
    return;
  }
  // ---- the two base records (FUSED: computed here by the warp, lane = joint; otherwise staged from k_model_base's output) and the linearisation point
  double* xs = sxu[warp]; double* us = xs + 24; double* xns = xs + 48; double* xrs = xs + 72;
  if (FUSED) {
    double* x2 = &sA2[warp][0][0];   // scratch for the second RK2 evaluation point (sA2 is filled later)
    if (lane < NX) { xs[lane] = xg[lane]; xns[lane] = xng[lane]; xrs[lane] = d.xref[(nb + k) * NX + lane]; }
    if (lane < NU) us[lane] = d.s_u[(nb + k) * NU + lane];
    __syncwarp();
    const double* jc = sjc[lane < NJ ? lane : 0];
    warp_model_base<NJ>(xs, us, sbase[warp], lane, jc);
    __syncwarp();
    if (lane < NX) x2[lane] = xs[lane] + d.st_dt[nb + k] * sbase[warp][BD::B_F + lane];
    __syncwarp();
    warp_model_base<NJ>(x2, us, sbase[warp] + BASE, lane, jc);
  } else {
    const double* __restrict__ bg = d.base + (nb + k) * (size_t)(2 * BASE);
    constexpr int NIT = (2 * BASE + 31) / 32;
    double tmp[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) tmp[i] = (lane + 32 * i < 2 * BASE) ? bg[lane + 32 * i] : 0.0;
    if (lane < NX) { xs[lane] = xg[lane]; xns[lane] = xng[lane]; xrs[lane] = d.xref[(nb + k) * NX + lane]; }
    if (lane < NU) us[lane] = d.s_u[(nb + k) * NU + lane];
#pragma unroll
    for (int i = 0; i < NIT; ++i) if (lane + 32 * i < 2 * BASE) sbase[warp][lane + 32 * i] = tmp[i];
  }
  __syncwarp();
  lq_stage_columns<NJ>(d, nb, k, rec, sbase[warp], sbase[warp] + BASE, xs, us, xns, xrs, sA2[warp], lane);
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
};
template <int NJ>
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
  const int N = d.n_nodes[b] - 1;
  if (k0 >= N) return;
  const size_t nb = (size_t)b * d.NS;
  bool has[G], ev[G], comp[G];   // stage exists / is an event node / needs the model
  int first_comp = -1;
#pragma unroll
  for (int s = 0; s < G; ++s) {
    has[s] = k0 + s < N;
    ev[s] = has[s] && d.node_ev[nb + k0 + s] == 1;
    comp[s] = has[s] && !ev[s];
    if (comp[s] && first_comp < 0) first_comp = s;
  }
#pragma unroll
  for (int s = 0; s < G; ++s) {
    if (!has[s]) continue;
    const int k = k0 + s;
    double* xs = sm.xu[warp][s];
    if (lane < NX) { xs[lane] = d.s_x[(nb + k) * NX + lane]; xs[48 + lane] = d.s_x[(nb + k + 1) * NX + lane]; xs[72 + lane] = d.xref[(nb + k) * NX + lane]; }
    if (lane < NU) xs[24 + lane] = d.s_u[(nb + k) * NU + lane];
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
          if (lane < NX) x2[24 * s + lane] = comp[s] ? sm.xu[warp][s][lane] + d.st_dt[nb + k0 + s] * sm.base[warp][s][BD::B_F + lane] : 0.0;
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
    if (d.node_ev[nb + k] == 1) {   // [UPSTREAM] setupEventNode
      double sq = 0.0;
      if (lane < NX) { const double bi = xs[lane] - xs[48 + lane]; rec[D::R_B + lane] = bi; sq = bi * bi; }
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0) {
        rec[D::R_MISC + D::M_TYPE] = 1.0; rec[D::R_MISC + D::M_DT] = 0.0; rec[D::R_MISC + D::M_MODE] = -1.0;
        rec[D::R_MISC + D::M_PCOST] = 0.0; rec[D::R_MISC + D::M_PDYN] = sq; rec[D::R_MISC + D::M_PEQ] = 0.0;
      }
      continue;
    }
    lq_stage_columns<NJ>(d, nb, k, rec, sm.base[warp][s], sm.base[warp][s] + BASE, xs, xs + 24, xs + 48, xs + 72, sm.A2[warp], lane);
    __syncwarp();   // A2 is reused by the next stage
  }
}

// Split LQ variant ("lq_mode" 4): k_base_pack evaluates the two base records of every stage (same packed base pass as k_lq_pack) and writes them
// to global memory; k_lq_assemble<NJ, false> then runs the column pass with 128 registers / 16 warps per SM instead of 255 / 8.
template <int NJ>
__global__ void __launch_bounds__(128, BASE_BLOCKS) k_base_pack(Dev d) {
  using D = Dims<NJ>; using BD = BaseDims<NJ>;
  constexpr int NX = D::NX, NU = D::NU, WPB = 4, BASE = BD::BASE, SEG = LqPackSmem<NJ>::SEG, G = LqPackSmem<NJ>::G;
  __shared__ double sjc[NJ][28];
  __shared__ double sxu[WPB][G][3 * 24];   // x, u, x2
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // broadcast: lets the compiler treat the warp index as warp-uniform
  for (int i = threadIdx.x; i < NJ * 28; i += 128) (&sjc[0][0])[i] = d.jc[i];
  __syncthreads();
  const int NP = (d.NS + G - 1) / G;
  const int gw = blockIdx.x * WPB + warp;
  const int b = gw / NP, k0 = G * (gw % NP);
  if (b >= d.B) return;
  const int N = d.n_nodes[b] - 1;
  if (k0 >= N) return;
  const size_t nb = (size_t)b * d.NS;
  bool comp[G];
  int first_comp = -1;
#pragma unroll
  for (int s = 0; s < G; ++s) {
    comp[s] = k0 + s < N && d.node_ev[nb + k0 + s] != 1;
    if (comp[s] && first_comp < 0) first_comp = s;
  }
  if (first_comp < 0) return;
#pragma unroll
  for (int s = 0; s < G; ++s) {
    if (!comp[s]) continue;
    double* xs = sxu[warp][s];
    if (lane < NX) xs[lane] = d.s_x[(nb + k0 + s) * NX + lane];
    if (lane < NU) xs[24 + lane] = d.s_u[(nb + k0 + s) * NU + lane];
  }
  __syncwarp();
  const int h = lane / SEG;
  int ms = first_comp; bool own = false;
#pragma unroll
  for (int s = 0; s < G; ++s) if (h == s && comp[s]) { ms = s; own = true; }
  const double* xh = sxu[warp][ms]; const double* uh = xh + 24;
  double* bh = d.base + (nb + k0 + ms) * (size_t)(2 * BASE);
  const int jl = lane % SEG;
  const double* jc = sjc[jl < NJ ? jl : 0];
  warp_model_base<NJ, SEG>(xh, uh, bh, lane, jc, own);
  __syncwarp();
#pragma unroll
  for (int s = 0; s < G; ++s)
    if (comp[s] && lane < NX) sxu[warp][s][48 + lane] = sxu[warp][s][lane] + d.st_dt[nb + k0 + s] * d.base[(nb + k0 + s) * (size_t)(2 * BASE) + BD::B_F + lane];
  __syncwarp();
  warp_model_base<NJ, SEG>(xh + 48, uh, bh + BASE, lane, jc, own);
}

}  // namespace bmpc
