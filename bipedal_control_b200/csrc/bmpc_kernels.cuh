// CUDA kernels of one SQP tick (sm_100a, FP64).  Kernel <-> reference mapping:
//   k_time_grid / k_node_setup : timeDiscretizationWithEvents, ModeSchedule::modeAtTime, TargetTrajectories::getDesiredState,
//                                SwingTrajectoryPlanner::getZvelocityConstraint (foot_planner/SwingTrajectoryPlanner.cpp:50-118),
//                                multiple_shooting::initializeStateInputTrajectories + BipedalRobotInitializer::compute [UPSTREAM / initializer]
//   k_lq                       : multiple_shooting::setupIntermediateNode / setupEventNode (dynamics RK2 sensitivity, cost, soft friction cone,
//                                zero-force / zero-velocity / normal-velocity constraints)  -> compact LQ record
//   k_project                  : LinearAlgebra::luConstraintProjection replacement (Householder QR, min-norm particular solution)
//   k_riccati                  : changeOfInputVariables + HPIPM backward Riccati + feedback gains (DMMA m8n8k4 tiles in shared memory)
//   k_forward                  : HPIPM forward substitution, armijoDescentMetric, PerformanceIndex reduction
//   k_linesearch_eval/k_accept : SqpSolver::computePerformance + FilterLinesearch::acceptStep
//   k_update / k_policy_fill   : incrementTrajectory + multiple_shooting::toPrimalSolution (LinearController uff, K)
#pragma once
#include "bmpc_device.cuh"

namespace bmpc {

constexpr double WEAK_EPS = 1e-6;   // [UPSTREAM] numeric_traits::weakEpsilon
constexpr int WS_THREADS = 128;     // threads per CTA of the Riccati kernel (one instance per CTA)
#ifndef LQ_MIN_BLOCKS
#define LQ_MIN_BLOCKS 8
#endif
#ifndef LQ_PAIR_BLOCKS
#define LQ_PAIR_BLOCKS 2
#endif
#ifndef LS_BLOCKS
#define LS_BLOCKS 6
#endif
#ifndef LS2_BLOCKS
#define LS2_BLOCKS 4
#endif
#ifndef PROJ_BLOCKS
#define PROJ_BLOCKS 4
#endif
#ifndef BASE_BLOCKS
#define BASE_BLOCKS 4
#endif
#ifndef RIC_BLOCKS
#define RIC_BLOCKS 3
#endif
#ifndef RIC_WPC
#define RIC_WPC 4   // independent instances (warps) per CTA of k_riccati_warp
#endif
#ifndef LQ_FUSED_BLOCKS
#define LQ_FUSED_BLOCKS 2
#endif

template <int NJ> struct RDims;
template <int NJ> struct SDims;

struct Dev {
  int B, NS, ME, TP, npts;
  double dt_nom, horizon;
  const double* t0; const double* x0;
  const double* tgt_t; const double* tgt_x;
  const int* n_ev; const double* ev_t; const int* ev_mode;
  int* n_nodes; double* node_t; int* node_ev; double* st_t; double* st_dt; int* st_mode;
  double* xref; double* zref;
  const int* p_n; const double* p_t; const double* p_x; const double* p_u;   // previous primal solution (warm start)
  double* s_x; double* s_u; double* s_uff; double* s_K;                      // new primal solution / linearisation point
  double* lq; double* proj; double* stage; double* ric; double* base; const double* jc;
  double* dx; double* du;
  double* perf_trial; double* perf; double* alpha; double* norms; int* done; int* status; int* counters;
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

// ------------------------------------------------------------------------------------------------ K0a: time grid
__global__ void k_time_grid(Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  const double t0 = d.t0[b], tf = t0 + d.horizon, dt = d.dt_nom;
  const double* ev = d.ev_t + (size_t)b * d.ME; const int ne = d.n_ev[b];
  double* nt = d.node_t + (size_t)b * d.NS; int* nev = d.node_ev + (size_t)b * d.NS;
  const double dt_min = 10.0 * WEAK_EPS;
  int n = 0; bool overflow = false;
  nt[0] = t0; nev[0] = 0; n = 1;
  int nextEvent = lower_bound_d(ev, ne, t0);
  double nextT = t0; int nextE = 0;
  while (nt[n - 1] < tf) {
    nextT = nextT + dt; nextE = 0;
    if (nextEvent < ne && nextT >= ev[nextEvent]) { nextT = ev[nextEvent]; nextE = 1; ++nextEvent; }
    if (nextT >= tf) { nextT = tf; nextE = 0; }
    if (nextT > nt[n - 1] + dt_min) { if (n >= d.NS) { overflow = true; break; } nt[n] = nextT; nev[n] = nextE; ++n; }
    else { nt[n - 1] = nextT; nev[n - 1] = nextE; }
    if (nextE == 1) { if (n >= d.NS) { overflow = true; break; } nt[n] = nextT; nev[n] = 2; ++n; }
  }
  if (overflow) { atomicOr(&d.status[b], 32); nt[n - 1] = tf; nev[n - 1] = 0; }
  d.n_nodes[b] = n;
  double* stt = d.st_t + (size_t)b * d.NS; double* std_ = d.st_dt + (size_t)b * d.NS;
  for (int i = 0; i + 1 < n; ++i) {
    const double ts = nev[i] == 2 ? nt[i] + WEAK_EPS : nt[i];
    const double te = nev[i + 1] == 1 ? nt[i + 1] - WEAK_EPS : nt[i + 1];
    stt[i] = ts; std_[i] = (nev[i] == 1) ? 0.0 : te - ts;
  }
}

// ------------------------------------------------------------------------------------------------ K0b: per-node references + warm start
// swing height velocity of leg `leg` at time t (foot_planner/SwingTrajectoryPlanner.cpp:50-118, SplineCpg.cpp:38-60, CubicSpline.cpp:38-75)
__device__ inline double swing_zvel(const double* ev, const int* modes, int ne, int leg, double t, int* status) {
  const int np = ne + 1;
  const int p = lower_bound_d(ev, ne, t);
  int start = -1;
  for (int ip = p - 1; ip >= 0; --ip) if (leg_in_stance(modes[ip], leg)) { start = ip; break; }
  int fin = np - 1;
  for (int ip = p + 1; ip < np; ++ip) if (leg_in_stance(modes[ip], leg)) { fin = ip - 1; break; }
  if (start < 0 || fin >= np - 1) { atomicOr(status, 4); return 0.0; }
  const double ts = ev[start], tf = ev[fin];
  const double scaling = fmin(1.0, (tf - ts) / c_model.swing_time_scale);
  const double mid_t = 0.5 * (ts + tf), mid_h = scaling * c_model.swing_height;
  double t_a, p_a, v_a, t_b, p_b, v_b;
  if (t < mid_t) { t_a = ts; p_a = 0.0; v_a = scaling * c_model.liftoff_vel; t_b = mid_t; p_b = mid_h; v_b = 0.0; }
  else { t_a = mid_t; p_a = mid_h; v_a = 0.0; t_b = tf; p_b = 0.0; v_b = scaling * c_model.touchdown_vel; }
  const double dts = t_b - t_a, dp = p_b - p_a, dv = v_b - v_a;
  const double c1 = v_a * dts, c2 = -(3.0 * v_a + dv) * dts + 3.0 * dp, c3 = (2.0 * v_a + dv) * dts - 2.0 * dp;
  const double tn = (t - t_a) / dts;
  return (3.0 * c3 * tn * tn + 2.0 * c2 * tn + c1) / dts;
}

template <int NJ>
__global__ void k_node_setup(Dev d) {
  constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = gid / d.NS, k = gid % d.NS;
  if (b >= d.B) return;
  const int n = d.n_nodes[b];
  if (k >= n) return;
  const int N = n - 1;
  const size_t nb = (size_t)b * d.NS;
  const double* ev = d.ev_t + (size_t)b * d.ME; const int* modes = d.ev_mode + (size_t)b * (d.ME + 1); const int ne = d.n_ev[b];
  const int* nev = d.node_ev + nb;
  const double* nt = d.node_t + nb;
  const double* stt = d.st_t + nb; const double* std_ = d.st_dt + nb;
  // ---- stage references
  if (k < N) {
    int mode = -1;
    if (nev[k] != 1) {
      const double t = stt[k];
      mode = modes[lower_bound_d(ev, ne, t)];
      interp_vec(d.tgt_t + (size_t)b * d.TP, d.tgt_x + (size_t)b * d.TP * NX, d.npts, NX, t, d.xref + (nb + k) * NX);
      for (int leg = 0; leg < 2; ++leg)
        d.zref[(nb + k) * 2 + leg] = leg_in_stance(mode, leg) ? 0.0 : swing_zvel(ev, modes, ne, leg, t, &d.status[b]);
    }
    d.st_mode[nb + k] = mode;
  }
  // ---- initial guess: [UPSTREAM] multiple_shooting::initializeStateInputTrajectories
  const int pn = d.p_n ? d.p_n[b] : 0;
  const double* pt = d.p_t + nb; const double* px = d.p_x + nb * NX; const double* pu = d.p_u + nb * NU;
  double stateTill = nt[0], inputTill = nt[0];
  if (pn >= 2) { stateTill = pt[pn - 1]; inputTill = pt[pn - 2]; }
  auto interval_uses_initializer = [&](int i) {   // interval i = [node i, node i+1]; true also for event nodes (state copied)
    if (nev[i] == 1) return true;
    const double ti = stt[i], tn = stt[i] + std_[i];
    return (ti > inputTill || tn > stateTill);
  };
  // state of node k
  int j = k;
  while (j > 0 && interval_uses_initializer(j - 1)) --j;
  double* xo = d.s_x + (nb + k) * NX;
  if (j == 0) {
    const double tinit = nev[0] == 2 ? nt[0] + WEAK_EPS : nt[0];
    if (tinit < stateTill) interp_vec(pt, px, pn, NX, tinit, xo);
    else for (int i = 0; i < NX; ++i) xo[i] = d.x0[(size_t)b * NX + i];
  } else {
    interp_vec(pt, px, pn, NX, stt[j - 1] + std_[j - 1], xo);
  }
  // input of stage k
  if (k < N) {
    double* uo = d.s_u + (nb + k) * NU;
    if (nev[k] == 1) { for (int i = 0; i < NU; ++i) uo[i] = 0.0; }
    else if (interval_uses_initializer(k)) {   // initialization/BipedalRobotInitializer.cpp:56-63 + common/utils.h:63-77
      const int mode = modes[lower_bound_d(ev, ne, stt[k])];
      const bool s0 = leg_in_stance(mode, 0), s1 = leg_in_stance(mode, 1);
      const int ns = 2 * (int(s0) + int(s1));
      const double fz = ns > 0 ? c_model.total_mass * 9.81 / ns : 0.0;
      for (int i = 0; i < NU; ++i) uo[i] = 0.0;
      if (s0) { uo[2] = fz; uo[5] = fz; }
      if (s1) { uo[8] = fz; uo[11] = fz; }
    } else interp_vec(pt, pu, pn, NU, stt[k], uo);
  }
}

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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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

// Lane roles after the QR (one column of W = [Px | Pe | N] per lane, in FULL-STATE column order so that the tensor-core tiles line up with
// the stage record): lane L < 24 = state column L (L = 6 carries the affine column Pe: base-position columns 6..8 of Px are structurally
// zero; 7, 8 and the padding lanes stay zero), lane 24 + t = null-space column t.  The reduced input is ordered [null-space (mj) | closed-contact
// forces (3 nclosed)], so the null rows / columns are tile aligned as well.
// The change of input variables runs on the FP64 tensor cores:
//   M  = W^T (Rj_eff W) (32 x 32)  : tiles (a, b < 3) are Qt in accumulator-fragment order (stored with one 16-byte store per lane and tile),
//                                    row 6 / column 6 hold the qt / rt corrections, tiles (3, b < 3) are Pt, tile (3, 3) the null block of Rt;
//   AJ = B_d[:, joints] W (16 x 32): joint part of At rows 3..11 (stored as row-major pairs), bt, null-space columns of Bt.
// Z^T = W^T Rj is formed first and reused from registers as the B operand of M (same register-chaining trick as k_riccati_warp).
template <int NJ>
__global__ void __launch_bounds__(128, PROJ_BLOCKS) k_project(Dev d) {
  using D = Dims<NJ>; using S = SDims<NJ>;
  constexpr int NX = D::NX, NU = D::NU, NXA = D::NXA, MP = S::MP, LDA = S::LDA;
  constexpr int WPB = 4;
  constexpr int LDW = 34, LDR = 18, LDJ = 20;   // leading dimensions = 2 mod 4: k-permuted fragment loads are conflict free
  __shared__ double sM[WPB][NJ][12];     // Dv^T  (NJ x r), r <= 10
  __shared__ double sV[WPB][10][NJ];     // Householder vectors (zero padded)
  __shared__ double sBeta[WPB][20];      // beta (10) | 1 / R[k][k] (10)
  __shared__ double sG[WPB][10][NXA + 1];   // [Cv | ev]; after the triangular solves: the padded joint block of B_d
  __shared__ double sBd[WPB][9 * (12 + NJ)];  // B_d rows 3..11
  __shared__ double sW[WPB][16][LDW];        // W, rows >= NJ zero
  __shared__ double sMisc[WPB][32];          // r_j (16) | open-contact correction of bt rows 3..11 (16)
  __shared__ double sRjP[16][LDR];           // joint block of R (model constant), zero padded
  __shared__ double sQd[24];
  for (int i = threadIdx.x; i < 16 * LDR; i += 128) { const int rr_ = i / LDR, cc_ = i % LDR; sRjP[rr_][cc_] = (rr_ < NJ && cc_ < NJ) ? c_model.Rjoint[rr_ * NJ + cc_] : 0.0; }
  if (threadIdx.x < 24) sQd[threadIdx.x] = threadIdx.x < NX ? c_model.Qdiag[threadIdx.x] : 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * WPB + warp;
  const int b = gw / d.NS, k = gw % d.NS;
  if (b >= d.B) return;
  const int N = d.n_nodes[b] - 1;
  if (k >= N) return;
  const size_t nb = (size_t)b * d.NS;
  const double* rec = d.lq + (nb + k) * D::REC;
  double* out = d.proj + (nb + k) * D::PREC;
  double* so = d.stage + (nb + k) * S::SREC;
  if (d.node_ev[nb + k] == 1) {   // event stage: only b is needed
    for (int i = lane; i < NX; i += 32) so[S::S_B + i] = rec[D::R_B + i];
    if (lane == 0) { so[S::S_META + S::T_TYPE] = 1.0; so[S::S_META + S::T_M] = 0.0; so[S::S_META + S::T_MJ] = 0.0; so[S::S_META + S::T_NCLOSED] = 0.0; so[S::S_META + S::T_DT] = 0.0; so[S::S_META + S::T_MODE] = -1.0; }
    return;
  }
  const DevModel& M = c_model;
  const int r = (int)rec[D::R_MISC + D::M_NROWS];
  double (*Mt)[12] = sM[warp]; double (*V)[NJ] = sV[warp]; double* beta = sBeta[warp]; double* rinv = sBeta[warp] + 10; double (*G)[NXA + 1] = sG[warp];
  double (*W)[LDW] = sW[warp];
  const double* Bd = sBd[warp];
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
  bool anomaly = false;
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
  // lane roles (see the header comment)
  const int mj = NJ - r;
  const bool is_x = lane < 6 || (lane >= 9 && lane < NX), is_aff = lane == 6, is_rhs = is_x || is_aff;
  const bool is_null = lane >= 24 && lane - 24 < mj;
  const int gc = is_aff ? NXA : (lane < 6 ? lane : lane - 3);   // column of [Cv | ev] / compressed column index of this lane
  double y[NJ];
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
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) W[i][lane] = (i < NJ) ? y[i < NJ ? i : 0] : 0.0;   // idle lanes hold y = 0
  const double dt = rec[D::R_MISC + D::M_DT], dq = rec[D::R_MISC + D::M_DQ], dr = rec[D::R_MISC + D::M_DR];
  const int mode = (int)rec[D::R_MISC + D::M_MODE];
  const bool st0 = leg_in_stance(mode, 0), st1 = leg_in_stance(mode, 1);
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
  for (int i = lane; i < 9 * LDJ; i += 32) { const int rr_ = i / LDJ, cc_ = i % LDJ; Bj[rr_][cc_] = (cc_ < NJ) ? Bd[rr_ * NU + 12 + cc_] : 0.0; }
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
  if (is_x) {
#pragma unroll
    for (int l = 0; l < NJ; ++l) so[S::S_AB + (12 + l) * LDA + lane] = dt * y[l] + ((12 + l == lane) ? 1.0 : 0.0);   // At rows 12..: I + dt Pxj
  } else if (is_aff) {
#pragma unroll
    for (int l = 0; l < NJ; ++l) so[S::S_B + 12 + l] = rec[D::R_B + 12 + l] + dt * y[l];                              // bt rows 12..
    const double f = dt / M.total_mass;
    for (int qq = 0; qq < 3; ++qq) {   // rows 0..2 of bt: B_d rows 0..2 = dt/m on the force columns
      double bb = rec[D::R_B + qq];
      for (int cn = 0; cn < NCON; ++cn) if (!(cn / 2 == 0 ? st0 : st1)) bb -= f * rec[D::R_FO + 3 * cn + qq];
      so[S::S_B + qq] = bb;
    }
  } else if (lane >= 24) {
#pragma unroll
    for (int l = 0; l < NJ; ++l) so[S::S_AB + (12 + l) * LDA + lane] = dt * y[l];   // Bt rows 12.., reduced columns 0..7: dt N (zero beyond mj)
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
        if (nt == 0 && q == 3 && R < NX && R != 6) so[S::S_Q + R] = rec[D::R_Q + R] + ((R == 7 || R == 8) ? 0.0 : c0);   // qt = q + Px^T t1
        double v0 = (R == 6 || C0 == 6) ? 0.0 : c0, v1 = (R == 6) ? 0.0 : c1;
        if (R == C0 && R < NX) v0 += dt * sQd[R] + dq;
        if (R == C0 + 1 && R < NX) v1 += dt * sQd[R] + dq;
        *reinterpret_cast<double2*>(so + S::S_QF + (mt * 3 + nt) * 64 + 2 * lane) = make_double2(v0, v1);
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
          const double v0 = ((C0 >= 6 && C0 <= 8) ? 0.0 : c0) + ((sr_ == C0) ? 1.0 : 0.0);
          const double v1 = ((C0 + 1 >= 6 && C0 + 1 <= 8) ? 0.0 : c1) + ((sr_ == C0 + 1) ? 1.0 : 0.0);
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
  if (lane >= 6 && lane < 9) so[S::S_Q + lane] = rec[D::R_Q + lane];
}

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

// ------------------------------------------------------------------------------------------------ DMMA tile GEMM in shared memory
// C[8MT x 8NT] = (ACC ? C : 0) + sign * op(A) op(B), K = 4 KT.  TA: A is given transposed (As[k][m]); TB: B is given transposed (Bs[n][k]).
// Output tiles are distributed round-robin over warps [W0, W0 + NW) of the CTA; each warp interleaves the k-loops of its tiles
// (independent accumulator chains).  All leading dimensions are == 4 or 12 (mod 16) doubles: every fragment load is bank-conflict free.
template <int MT, int NT, int KT, bool TA, bool TB, bool ACC, bool NEG, int NW, int W0>
__device__ __forceinline__ void gemm_tiles(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb, double* __restrict__ C, int ldc, int warp, int lane) {
  constexpr int TPW = (MT * NT + NW - 1) / NW;
  const int lr = lane >> 2, lc = lane & 3;
  const int w = warp - W0;
  if (w < 0 || w >= NW) return;
  double c0[TPW], c1[TPW];
  int mt[TPW], nt[TPW];
#pragma unroll
  for (int i = 0; i < TPW; ++i) {
    const int t = w + i * NW;
    mt[i] = t / NT; nt[i] = t % NT;
    c0[i] = 0.0; c1[i] = 0.0;
    if (ACC && t < MT * NT) { const double* cp = C + (8 * mt[i] + lr) * ldc + 8 * nt[i] + 2 * lc; c0[i] = cp[0]; c1[i] = cp[1]; }
  }
#pragma unroll
  for (int kk = 0; kk < KT; ++kk) {
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
      if (w + i * NW < MT * NT) {
        double a = TA ? A[(4 * kk + lc) * lda + 8 * mt[i] + lr] : A[(8 * mt[i] + lr) * lda + 4 * kk + lc];
        const double bb = TB ? B[(8 * nt[i] + lr) * ldb + 4 * kk + lc] : B[(4 * kk + lc) * ldb + 8 * nt[i] + lr];
        if (NEG) a = -a;
        dmma884(c0[i], c1[i], a, bb);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TPW; ++i)
    if (w + i * NW < MT * NT) { double* cp = C + (8 * mt[i] + lr) * ldc + 8 * nt[i] + 2 * lc; cp[0] = c0[i]; cp[1] = c1[i]; }
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

// ------------------------------------------------------------------------------------------------ K2: backward Riccati recursion
// One CTA (4 warps) per instance; S, At, SA, Bt, SB, H, G live in shared memory, padded to NXP = 24 states / MP = 16 reduced inputs.
//   SA = S At, SB = S Bt, sb = s + S bt;  H = Pt + Bt^T SA, G = Rt + Bt^T SB, g = rt + Bt^T sb
//   G = L L^T, Y = L^-1 H, yg = L^-1 g                    (warp 0; warps 1-3 compute At^T SA meanwhile)
//   S' = Qt + At^T SA - Y^T Y,  s' = qt + At^T sb - Y^T yg
// The gains Kt = -L^-T Y are recovered off the critical path by k_policy_expand.
template <int NJ>
struct RicSmem {
  static constexpr int NXP = 24, MP = 16, LD = 28, LDM = 20;
  double S[NXP * LD], At[NXP * LD], SA[NXP * LD];
  double Bt[NXP * LDM], SB[NXP * LDM];
  double H[MP * LD], G[MP * LDM];
  double s[NXP], sb[NXP], bt[NXP], qt[NXP], snew[NXP], qd[NXP];
  double rt[MP], g[MP];
  double lcol[2][MP + 2];
  alignas(16) double stage[SDims<NJ>::SREC];   // TMA-staged stage record (refilled right after the scatter phase)
  alignas(8) unsigned long long bar;
};

template <int NJ>
struct RDims {
  static constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU, MP = 16;
  // written by k_riccati: Y[MP][NX], yg[MP], L[MP][MP];  written by k_policy_expand: kappa, Phi, phi, ghat, misc
  static constexpr int K_Y = 0, K_YG = K_Y + MP * NX, K_L = K_YG + MP, K_KAP = K_L + MP * MP, K_PHI = K_KAP + NU, K_SPHI = K_PHI + NX * NX, K_G = K_SPHI + NX,
                       K_MISC = K_G + NX, KREC = ((K_MISC + 2 + 3) / 4) * 4;
};

template <int NJ>
__global__ void __launch_bounds__(WS_THREADS, 4) k_riccati(Dev d) {
  using D = Dims<NJ>; using R = RDims<NJ>; using SM = RicSmem<NJ>; using S = SDims<NJ>;
  constexpr int NX = D::NX, NXA = D::NXA, NXR = S::NXR, NXP = SM::NXP, MP = SM::MP, LD = SM::LD, LDM = SM::LDM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = d.n_nodes[b] - 1;
  const size_t nb = (size_t)b * d.NS;
  const double imass = 1.0 / c_model.total_mass;
  // warp w runs on SM sub-partition w % 4: rotate the serial roles (Cholesky, mat-vecs) over the CTAs so that co-resident CTAs
  // do not pile their serial FP64 work onto the same sub-partition
  const int cw = blockIdx.x & 3;            // Cholesky warp of this CTA
  const int vw = (warp - cw - 1) & 3;       // 0..2 for the other three warps, 3 for the Cholesky warp
  const int mw = (cw + 2) & 3;              // mat-vec warp
  // terminal value function: zero (no terminal cost installed)
  for (int i = tid; i < NXP * LD; i += WS_THREADS) { sm.S[i] = 0.0; sm.At[i] = 0.0; sm.SA[i] = 0.0; }
  for (int i = tid; i < NXP * LDM; i += WS_THREADS) { sm.Bt[i] = 0.0; sm.SB[i] = 0.0; }
  for (int i = tid; i < MP * LD; i += WS_THREADS) sm.H[i] = 0.0;   // padded columns must stay zero (shared memory is not cleared between CTAs)
  for (int i = tid; i < MP * LDM; i += WS_THREADS) sm.G[i] = 0.0;
  for (int i = tid; i < NXP; i += WS_THREADS) { sm.s[i] = 0.0; sm.sb[i] = 0.0; sm.bt[i] = 0.0; sm.qt[i] = 0.0; sm.snew[i] = 0.0; sm.qd[i] = 0.0; }
  constexpr unsigned REC_BYTES = S::SREC * sizeof(double);
  if (tid == 0) { mbar_init(&sm.bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (tid == 0 && N >= 1) tma_load_1d(sm.stage, d.stage + (nb + N - 1) * S::SREC, REC_BYTES, &sm.bar);
  unsigned phase_bit = 0;
  for (int k = N - 1; k >= 0; --k) {
    mbar_wait(&sm.bar, phase_bit);
    phase_bit ^= 1u;
    const double* sr = sm.stage;
    double* ric = d.ric + (nb + k) * R::KREC;
    const double* meta = sr + S::S_META;
    const bool is_event = meta[S::T_TYPE] != 0.0;
    if (is_event) {   // S unchanged (A = I, Q = 0, no input); s <- s + S b
      if (tid < NX) sm.bt[tid] = sr[S::S_B + tid];
      __syncthreads();
      if (tid == 0 && k >= 1) { fence_proxy_async(); tma_load_1d(sm.stage, d.stage + (nb + k - 1) * S::SREC, REC_BYTES, &sm.bar); }
      if (tid < NX) { double a = sm.s[tid]; for (int c = 0; c < NX; ++c) a += sm.S[tid * LD + c] * sm.bt[c]; sm.snew[tid] = a; }
      __syncthreads();
      if (tid < NX) sm.s[tid] = sm.snew[tid];
      __syncthreads();
      continue;
    }
    const int m = (int)meta[S::T_M], mj = (int)meta[S::T_MJ], nclosed = (int)meta[S::T_NCLOSED], mode = (int)meta[S::T_MODE];
    const double dt = meta[S::T_DT];
    const bool st0 = leg_in_stance(mode, 0);
    // ---- phase 1: copy the projected stage record into the padded operand matrices (lane = column, warp = row stride)
    {
      for (int r = warp; r < NX; r += 4) if (lane < NXP) sm.At[r * LD + lane] = sr[S::S_AB + r * S::LDA + lane];
      const int bc = lane & 15, bh = lane >> 4;
      for (int r = 2 * warp + bh; r < NX; r += 8) sm.Bt[r * LDM + bc] = sr[S::S_AB + r * S::LDA + 24 + bc];
      for (int r = warp; r < MP; r += 4) if (lane < NXP) sm.H[r * LD + lane] = sr[S::prf(r, lane)];           // H <- Pt
      for (int r = 2 * warp + bh; r < MP; r += 8) sm.G[r * LDM + bc] = sr[S::prf(r, 24 + bc)];                // G <- Rt
    }
    if (tid < NX) { sm.bt[tid] = sr[S::S_B + tid]; sm.qt[tid] = sr[S::S_Q + tid]; }
    if (tid >= 32 && tid < 32 + MP) sm.rt[tid - 32] = sr[S::S_R + tid - 32];
    // Qt entries this thread adds in phase 6 (pairs r <= c), held in registers so that the staging buffer can be refilled now
    constexpr int QPT = (NX + 3) / 4;
    double qreg[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
      const int r = warp + 4 * q, c = lane;
      double a = 0.0;
      if (r < NX && c < NX && c >= r) a = sr[S::qf(r, c)];
      qreg[q] = a;
    }
    __syncthreads();
    // the staging buffer is free: prefetch the next stage record while this stage is being processed (TMA, completes on the mbarrier)
    if (tid == 0 && k >= 1) { fence_proxy_async(); tma_load_1d(sm.stage, d.stage + (nb + k - 1) * S::SREC, REC_BYTES, &sm.bar); }
    // ---- phase 2: SA = S At, SB = S Bt, sb = s + S bt
    gemm_tiles<3, 3, 6, false, false, false, false, 4, 0>(sm.S, LD, sm.At, LD, sm.SA, LD, warp, lane);
    gemm_tiles<3, 2, 6, false, false, false, false, 4, 0>(sm.S, LD, sm.Bt, LDM, sm.SB, LDM, warp, lane);
    if (warp == mw && lane < NX) { const int r = lane; double a = sm.s[r]; for (int c = 0; c < NX; ++c) a += sm.S[r * LD + c] * sm.bt[c]; sm.sb[r] = a; }
    __syncthreads();
    // ---- phase 3: H += Bt^T SA, G += Bt^T SB, g = rt + Bt^T sb
    gemm_tiles<2, 3, 6, true, false, true, false, 4, 0>(sm.Bt, LDM, sm.SA, LD, sm.H, LD, warp, lane);
    gemm_tiles<2, 2, 6, true, false, true, false, 4, 0>(sm.Bt, LDM, sm.SB, LDM, sm.G, LDM, warp, lane);
    if (warp == mw && lane < MP) { const int c = lane; double a = sm.rt[c]; for (int r = 0; r < NX; ++r) a += sm.Bt[r * LDM + c] * sm.sb[r]; sm.g[c] = a; }
    __syncthreads();
    // ---- phase 4: warp 0: right-looking Cholesky of G fused with the forward substitution of [H | g];  warps 1-3: S <- At^T SA
    if (warp == cw) {
      // lane l < MP owns column l of G (lower part) ; lane c < NX owns column c of H ; lane NX owns g.
      // Shuffle-only right-looking elimination: the raw pivot column is broadcast from lane j while every lane computes the
      // reciprocal square root of the pivot redundantly, so the serial chain per pivot is shuffle -> rsqrt -> one FMA.
      double gc[MP], hc[MP];
#pragma unroll
      for (int i = 0; i < MP; ++i) { gc[i] = (lane < MP) ? sm.G[i * LDM + lane] : 0.0; hc[i] = (lane < NX) ? sm.H[i * LD + lane] : ((lane == NX) ? sm.g[i] : 0.0); }
      bool not_pd;
      switch (m) {   // reduced input dimensions that occur: H1 6 / 9 / 12 (FLY / single stance / double stance), G1 8 / 11 / 14
        case 6: not_pd = chol_forward<6, MP>(gc, hc, lane); break;
        case 9: not_pd = chol_forward<9, MP>(gc, hc, lane); break;
        case 12: not_pd = chol_forward<12, MP>(gc, hc, lane); break;
        case 8: not_pd = chol_forward<8, MP>(gc, hc, lane); break;
        case 11: not_pd = chol_forward<11, MP>(gc, hc, lane); break;
        case 14: not_pd = chol_forward<14, MP>(gc, hc, lane); break;
        default: not_pd = chol_forward<MP, MP>(gc, hc, lane); break;   // padded pivots are identity rows
      }
      if (not_pd && lane == 0) atomicOr(&d.status[b], 1);
#pragma unroll
      for (int i = 0; i < MP; ++i) {
        if (lane < MP) sm.G[i * LDM + lane] = (i >= lane) ? gc[i] : 0.0;
        if (lane < NX) sm.H[i * LD + lane] = hc[i]; else if (lane == NX) sm.g[i] = hc[i];
      }
    } else {
      gemm_tiles<3, 3, 6, true, false, false, false, 3, 0>(sm.At, LD, sm.SA, LD, sm.S, LD, vw, lane);
    }
    __syncthreads();
    // ---- phase 5: S -= Y^T Y ; s' = qt + At^T sb - Y^T yg ; write Y, yg, L for the policy kernel
    gemm_tiles<3, 3, 4, true, false, true, true, 4, 0>(sm.H, LD, sm.H, LD, sm.S, LD, warp, lane);
    if (warp == mw && lane < NX) {
      const int c = lane;
      double a = sm.qt[c];
      for (int r = 0; r < NX; ++r) a += sm.At[r * LD + c] * sm.sb[r];
      for (int r = 0; r < MP; ++r) a -= sm.H[r * LD + c] * sm.g[r];
      sm.snew[c] = a;
    }
    for (int r = warp; r < MP; r += 4) if (lane < NX) ric[R::K_Y + r * NX + lane] = sm.H[r * LD + lane];
    for (int r = 2 * warp + (lane >> 4); r < MP; r += 8) ric[R::K_L + r * MP + (lane & 15)] = sm.G[r * LDM + (lane & 15)];
    if (tid < MP) ric[R::K_YG + tid] = sm.g[tid];
    __syncthreads();
    // ---- phase 6: add Qt and symmetrise (each unordered pair (r, c) is owned by one thread)
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
      const int r = warp + 4 * q, c = lane;
      if (r < NX && c < NX && c >= r) {
        const double a = 0.5 * (sm.S[r * LD + c] + sm.S[c * LD + r]) + qreg[q];
        sm.S[r * LD + c] = a; sm.S[c * LD + r] = a;
      }
    }
    if (tid < NX) sm.s[tid] = sm.snew[tid];
    __syncthreads();
  }
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
  static constexpr int LDH = 44;
  alignas(16) double rec[SDims<NJ>::TMA_DOUBLES];   // AB | bt | qt | rt | meta (TMA destination)
  alignas(16) double HG[16 * LDH];                   // [H | G] fragments -> column layout for the Cholesky; afterwards [Y | L]
  double sb[24], gv[16];
  alignas(16) unsigned long long bar;
};

template <int NJ>
__global__ void __launch_bounds__(32 * RIC_WPC, RIC_BLOCKS) k_riccati_warp(Dev d) {
  using D = Dims<NJ>; using R = RDims<NJ>; using S = SDims<NJ>; using SM = RicWarpSmem<NJ>;
  constexpr int NX = D::NX, MP = S::MP, LDA = S::LDA, LDH = SM::LDH;
  constexpr unsigned TMA_BYTES = S::TMA_DOUBLES * sizeof(double);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  __syncwarp();
  if (lane == 0 && N >= 1) tma_load_1d(sm.rec, d.stage + (nb + N - 1) * S::SREC, TMA_BYTES, &sm.bar);
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
        if (lane == 0 && k >= 1) { fence_proxy_async(); tma_load_1d(sm.rec, d.stage + (nb + k - 1) * S::SREC, TMA_BYTES, &sm.bar); }
        continue;
      }
    }
    // accumulator initialisers, fragment ordered in global memory: [Pt | Rt] now, Qt below (in flight during the products)
    double HGf[2][5][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 5; ++c) { const double2 v = *reinterpret_cast<const double2*>(grec + S::S_PRF + (a * 5 + c) * 64 + 2 * lane); HGf[a][c][0] = v.x; HGf[a][c][1] = v.y; }
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
      for (int c = 0; c < 3; ++c) { const double2 v = *reinterpret_cast<const double2*>(grec + S::S_QF + (a * 3 + c) * 64 + 2 * lane); Sn[a][c][0] = v.x; Sn[a][c][1] = v.y; }
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
      if (lane < MP) sm.gv[lane] = gval;
    }
#pragma unroll
    for (int kb = 0; kb < 3; ++kb)
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        const double a0 = AB[(8 * kb + 2 * q) * LDA + 8 * mt + g], a1 = AB[(8 * kb + 2 * q + 1) * LDA + 8 * mt + g];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) { dmma884(Sn[mt][nt][0], Sn[mt][nt][1], a0, Z[nt][kb][0]); dmma884(Sn[mt][nt][0], Sn[mt][nt][1], a1, Z[nt][kb][1]); }
      }
    // ---- the staged record is free: prefetch the next stage while the Cholesky chain runs
    __syncwarp();
    if (lane == 0 && k >= 1) { fence_proxy_async(); tma_load_1d(sm.rec, d.stage + (nb + k - 1) * S::SREC, TMA_BYTES, &sm.bar); }
    // ---- step C: [H | G] fragments -> shared memory -> one column per lane ; Cholesky of G fused with the forward substitution of [H | g]
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 5; ++c) *reinterpret_cast<double2*>(&sm.HG[(8 * a + g) * LDH + 8 * c + 2 * q]) = make_double2(HGf[a][c][0], HGf[a][c][1]);
    __syncwarp();
    {
      double gc[MP], hc[MP];
#pragma unroll
      for (int i = 0; i < MP; ++i) { gc[i] = (lane < MP) ? sm.HG[i * LDH + 24 + lane] : 0.0; hc[i] = (lane < 24) ? sm.HG[i * LDH + lane] : ((lane == 24) ? sm.gv[i] : 0.0); }
      __syncwarp();
      bool not_pd;
      switch (m) {   // reduced input dimensions that occur: H1 6 / 9 / 12 (FLY / single stance / double stance), G1 8 / 11 / 14
        case 6: not_pd = chol_forward<6, MP>(gc, hc, lane); break;
        case 9: not_pd = chol_forward<9, MP>(gc, hc, lane); break;
        case 12: not_pd = chol_forward<12, MP>(gc, hc, lane); break;
        case 8: not_pd = chol_forward<8, MP>(gc, hc, lane); break;
        case 11: not_pd = chol_forward<11, MP>(gc, hc, lane); break;
        case 14: not_pd = chol_forward<14, MP>(gc, hc, lane); break;
        default: not_pd = chol_forward<MP, MP>(gc, hc, lane); break;   // padded pivots are identity rows
      }
      if (not_pd && lane == 0) atomicOr(&d.status[b], 1);
#pragma unroll
      for (int i = 0; i < MP; ++i) {
        if (lane < 24) sm.HG[i * LDH + lane] = hc[i];                       // Y
        if (lane < NX) ric[R::K_Y + i * NX + lane] = hc[i];
        if (lane < MP) ric[R::K_L + i * MP + lane] = (i >= lane) ? gc[i] : 0.0;   // L (reciprocal pivots on the diagonal)
        if (lane == 24) { sm.gv[i] = hc[i]; ric[R::K_YG + i] = hc[i]; }       // yg
      }
      __syncwarp();
      if (lane < 24) {   // s' -= Y^T yg
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < MP; ++i) a += hc[i] * sm.gv[i];
        s_l -= a;
      }
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
          for (int nt = 0; nt < 3; ++nt) dmma884(Sn[mt][nt][0], Sn[mt][nt][1], -y[mt], y[nt]);
      }
    }
    // ---- symmetrise: S = (S' + S'^T) / 2.  Element (8 nt + 2q + s, 8 mt + g) of tile (nt, mt) lives in lane (2q + s) * 4 + g / 2, slot g & 1
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
#ifdef RIC_NOSYM
          Sf[mt][nt][sl] = Sn[mt][nt][sl];
#else
          const int src = (2 * q + sl) * 4 + (g >> 1);
          const double t0 = __shfl_sync(0xffffffffu, Sn[nt][mt][0], src), t1 = __shfl_sync(0xffffffffu, Sn[nt][mt][1], src);
          Sf[mt][nt][sl] = 0.5 * (Sn[mt][nt][sl] + ((g & 1) ? t1 : t0));
#endif
        }
    __syncwarp();   // HG / gv / sb are rewritten by the next stage
  }
}

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

template <int NJ>
__global__ void __launch_bounds__(128, 4) k_policy_expand(Dev d) {
  using D = Dims<NJ>; using R = RDims<NJ>; using S = SDims<NJ>; using PS = PolSmem<NJ>;
  constexpr int NX = D::NX, NU = D::NU, NXA = D::NXA, MP = S::MP, WPB = 4, NT = PS::NT, LDK = PS::LDK, NTILES = 3 * NT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  const double imass = 1.0 / c_model.total_mass;
  const double* __restrict__ prj = d.proj + (nb + k) * D::PREC;
  const int lr = lane >> 2, lc = lane & 3;
  // ---- issue every global load up front (independent: their latency overlaps with the back substitution below)
  const bool active = lane <= NX;
  double z[MP];
#pragma unroll
  for (int i = 0; i < MP; ++i) z[i] = active ? (lane < NX ? ric[R::K_Y + i * NX + lane] : ric[R::K_YG + i]) : 0.0;
  double lreg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) lreg[i] = ric[R::K_L + lane + 32 * i];
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
  // ---- back substitution Kt = -L^-T Y (lane c < NX: column c; lane NX: kt from yg)
#pragma unroll
  for (int i = MP - 1; i >= 0; --i) {
    double a = z[i];
#pragma unroll
    for (int l = i + 1; l < MP; ++l) a -= sm.L[l][i] * z[l];
    z[i] = (i < m) ? a * sm.L[i][i] : 0.0;   // the record stores 1 / L[i][i] on the diagonal
  }
#pragma unroll
  for (int i = 0; i < MP; ++i) { z[i] = -z[i]; if (lane < LDK) sm.Kt[i * LDK + lane] = z[i]; }
  __syncwarp();
  // ---- [Phi | phi] = [At | bt] + Bt [Kt | kt] on the FP64 tensor cores: 9 tiles x 4 k-steps, results stored straight from the fragments
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
#pragma unroll
    for (int t = 0; t < NTILES; ++t) dmma884(c0[t], c1[t], af[t / NT][kk], sm.Kt[(4 * kk + lc) * LDK + 8 * (t % NT) + lr]);
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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    d.alpha[b] = 1.0; d.done[b] = 0;
  }
}

// ------------------------------------------------------------------------------------------------ K4: line-search trial evaluation, one thread per (instance, stage)
template <int NJ>
__global__ void __launch_bounds__(64, LS_BLOCKS) k_linesearch_eval(Dev d) {
  using D = Dims<NJ>;
  constexpr int NX = D::NX, NU = D::NU;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = gid / d.NS, k = gid % d.NS;
  if (b >= d.B) return;
  if (d.done[b]) return;
  const int N = d.n_nodes[b] - 1;
  if (k >= N) return;
  const size_t nb = (size_t)b * d.NS;
  const double al = d.alpha[b];
  double* out = d.perf_trial + (nb + k) * 3;
  double x[NX], xn[NX], u[NU];
  for (int i = 0; i < NX; ++i) { x[i] = d.s_x[(nb + k) * NX + i] + al * d.dx[(nb + k) * NX + i]; xn[i] = d.s_x[(nb + k + 1) * NX + i] + al * d.dx[(nb + k + 1) * NX + i]; }
  if (d.node_ev[nb + k] == 1) {
    double s = 0.0; for (int i = 0; i < NX; ++i) { const double e = x[i] - xn[i]; s += e * e; }
    out[0] = 0.0; out[1] = s; out[2] = 0.0; return;
  }
  for (int i = 0; i < NU; ++i) u[i] = d.s_u[(nb + k) * NU + i] + al * d.du[(nb + k) * NU + i];
  const double dt = d.st_dt[nb + k]; const int mode = d.st_mode[nb + k];
  ModelEval<NJ> E1;
  model_eval<NJ, 0>(x, u, E1, nullptr);
  double x2[NX], k1[NX];
  for (int i = 0; i < NX; ++i) { k1[i] = E1.f[i]; x2[i] = x[i] + dt * k1[i]; }
  v3 vc[NCON]; for (int c = 0; c < NCON; ++c) vc[c] = E1.vc[c];
  model_eval<NJ, 0>(x2, u, E1, nullptr);
  double s = 0.0;
  for (int i = 0; i < NX; ++i) { const double e = x[i] + 0.5 * dt * (k1[i] + E1.f[i]) - xn[i]; s += e * e; }
  double peq = 0.0;
  for (int leg = 0; leg < 2; ++leg) {
    const int ca = 2 * leg, cb = 2 * leg + 1;
    if (leg_in_stance(mode, leg)) peq += dot(vc[ca], vc[ca]) + dot(vc[cb], vc[cb]);
    else {
      const double zr = d.zref[(nb + k) * 2 + leg];
      for (int t = 0; t < 2; ++t) { const int c0 = t == 0 ? ca : cb; const double ev = vc[c0].z - zr; peq += ev * ev + u[3 * c0] * u[3 * c0] + u[3 * c0 + 1] * u[3 * c0 + 1] + u[3 * c0 + 2] * u[3 * c0 + 2]; }
    }
  }
  out[0] = dt * stage_cost_value<NJ>(mode, x, u, d.xref + (nb + k) * NX);
  out[1] = dt * s; out[2] = dt * peq;
}

// K4 (default): the same trial evaluation on the streaming, register-only flow map (model_values): no per-joint arrays, no local memory
template <int NJ>
__global__ void __launch_bounds__(64, LS2_BLOCKS) k_linesearch_eval2(Dev d) {
  using D = Dims<NJ>;
  constexpr int NX = D::NX, NU = D::NU;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = gid / d.NS, k = gid % d.NS;
  if (b >= d.B) return;
  if (d.done[b]) return;
  const int N = d.n_nodes[b] - 1;
  if (k >= N) return;
  const size_t nb = (size_t)b * d.NS;
  const double al = d.alpha[b];
  double* out = d.perf_trial + (nb + k) * 3;
  const double* __restrict__ gx = d.s_x + (nb + k) * NX; const double* __restrict__ gdx = d.dx + (nb + k) * NX;
  const double* __restrict__ gu = d.s_u + (nb + k) * NU; const double* __restrict__ gdu = d.du + (nb + k) * NU;
  if (d.node_ev[nb + k] == 1) {
    double s = 0.0;
    for (int i = 0; i < NX; ++i) { const double e = gx[i] + al * gdx[i] - (gx[NX + i] + al * gdx[NX + i]); s += e * e; }
    out[0] = 0.0; out[1] = s; out[2] = 0.0; return;
  }
  const double dt = d.st_dt[nb + k]; const int mode = d.st_mode[nb + k];
  double xb[12], qj[NJ], uf[12], qd[NJ];
#pragma unroll
  for (int i = 0; i < 12; ++i) { xb[i] = gx[i] + al * gdx[i]; uf[i] = gu[i] + al * gdu[i]; }
#pragma unroll
  for (int j = 0; j < NJ; ++j) { qj[j] = gx[12 + j] + al * gdx[12 + j]; qd[j] = gu[12 + j] + al * gdu[12 + j]; }
  // stage cost at (x, u): tracking + soft friction cones (cost/BipedalRobotQuadraticTrackingCost.h:57-63, common/utils.h:63-77)
  const DevModel& M = c_model;
  const bool st0 = leg_in_stance(mode, 0), st1 = leg_in_stance(mode, 1);
  double cost = 0.0;
  {
    const double* __restrict__ xr = d.xref + (nb + k) * NX;
#pragma unroll
    for (int i = 0; i < 12; ++i) { const double e = xb[i] - xr[i]; cost += 0.5 * M.Qdiag[i] * e * e; }
#pragma unroll
    for (int j = 0; j < NJ; ++j) { const double e = qj[j] - xr[12 + j]; cost += 0.5 * M.Qdiag[12 + j] * e * e; }
    const int nst = 2 * (int(st0) + int(st1));
    const double fz = nst > 0 ? M.total_mass * 9.81 / nst : 0.0;
#pragma unroll
    for (int c = 0; c < NCON; ++c) {
      const bool st = (c / 2 == 0) ? st0 : st1;
      const double ex = uf[3 * c], ey = uf[3 * c + 1], ez = uf[3 * c + 2] - (st ? fz : 0.0);
      cost += 0.5 * (M.Rforce[3 * c] * ex * ex + M.Rforce[3 * c + 1] * ey * ey + M.Rforce[3 * c + 2] * ez * ez);
      if (st) { double p, dp, ddp; barrier_penalty(friction_cone(uf[3 * c], uf[3 * c + 1], uf[3 * c + 2]), p, dp, ddp); cost += p; }
    }
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < NJ; ++j) s += M.Rjoint[i * NJ + j] * qd[j];
      cost += 0.5 * qd[i] * s;
    }
  }
  // RK2 (Heun) defect against the next node; rows 12.. of the flow map are qd, so their defect is x + dt qd - x_next
  double sdef = 0.0;
#pragma unroll
  for (int j = 0; j < NJ; ++j) { const double e = qj[j] + dt * qd[j] - (gx[NX + 12 + j] + al * gdx[NX + 12 + j]); sdef += e * e; }
  double k1[12], k2[12]; v3 vc[NCON], vc2[NCON];
  model_values<NJ>(xb, qj, uf, qd, k1, vc);
  double peq = 0.0;
#pragma unroll
  for (int leg = 0; leg < 2; ++leg) {
    const int ca = 2 * leg, cb = 2 * leg + 1;
    if (leg == 0 ? st0 : st1) peq += dot(vc[ca], vc[ca]) + dot(vc[cb], vc[cb]);
    else {
      const double zr = d.zref[(nb + k) * 2 + leg];
#pragma unroll
      for (int t = 0; t < 2; ++t) { const int c0 = t == 0 ? ca : cb; const double ev = vc[c0].z - zr; peq += ev * ev + uf[3 * c0] * uf[3 * c0] + uf[3 * c0 + 1] * uf[3 * c0 + 1] + uf[3 * c0 + 2] * uf[3 * c0 + 2]; }
    }
  }
  double xb2[12], qj2[NJ];
#pragma unroll
  for (int i = 0; i < 12; ++i) xb2[i] = xb[i] + dt * k1[i];
#pragma unroll
  for (int j = 0; j < NJ; ++j) qj2[j] = qj[j] + dt * qd[j];
  model_values<NJ>(xb2, qj2, uf, qd, k2, vc2);
#pragma unroll
  for (int i = 0; i < 12; ++i) { const double e = xb[i] + 0.5 * dt * (k1[i] + k2[i]) - (gx[NX + i] + al * gdx[NX + i]); sdef += e * e; }
  out[0] = dt * cost; out[1] = dt * sdef; out[2] = dt * peq;
}

// ------------------------------------------------------------------------------------------------ K5: filter line search acceptance, one warp per instance
// [UPSTREAM] FilterLinesearch::acceptStep (g_max, g_min: task.info:72-73; gamma_c 1e-6, armijoFactor 1e-4, alpha_decay 0.5, alpha_min 1e-4)
template <int NJ>
__global__ void __launch_bounds__(128) k_accept(Dev d) {
  constexpr int NX = Dims<NJ>::NX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= d.B) return;
  if (d.done[b]) return;
  const int N = d.n_nodes[b] - 1;
  const size_t nb = (size_t)b * d.NS;
  double pc = 0.0, pd = 0.0, pe = 0.0;
  for (int k = lane; k < N; k += 32) { const double* p = d.perf_trial + (nb + k) * 3; pc += p[0]; pd += p[1]; pe += p[2]; }
  const double al = d.alpha[b];
  if (lane < NX) { const double e = d.x0[(size_t)b * NX + lane] - (d.s_x[nb * NX + lane] + al * d.dx[nb * NX + lane]); pd += e * e; }
  for (int o = 16; o > 0; o >>= 1) { pc += __shfl_xor_sync(0xffffffffu, pc, o); pd += __shfl_xor_sync(0xffffffffu, pd, o); pe += __shfl_xor_sync(0xffffffffu, pe, o); }
  if (lane == 0) {
    double* pf = d.perf + (size_t)b * 8;
    const double th0 = sqrt(pf[1] + pf[2]), th = sqrt(pd + pe);
    const double gamma_c = 1e-6, armijoFactor = 1e-4, alpha_decay = 0.5, alpha_min = 1e-4;
    const double armijo = pf[7];
    bool acc;
    if (th > c_model.g_max) acc = th < (1.0 - gamma_c) * th0;
    else if (th < c_model.g_min && th0 < c_model.g_min && armijo < 0.0) acc = pc < pf[0] + armijoFactor * al * armijo;
    else acc = (pc < pf[0] - gamma_c * th0) || (th < (1.0 - gamma_c) * th0);
    if (!(pc == pc) || !(pd == pd) || !(pe == pe)) { acc = false; atomicOr(&d.status[b], 8); }
    if (acc) { pf[3] = pc; pf[4] = pd; pf[5] = pe; pf[6] = al; d.done[b] = 1; }
    else {
      const double an = al * alpha_decay;
      const bool small = an * d.norms[2 * b] < c_model.delta_tol && an * d.norms[2 * b + 1] < c_model.delta_tol;
      if (small || an < alpha_min) { pf[3] = pf[0]; pf[4] = pf[1]; pf[5] = pf[2]; pf[6] = 0.0; d.alpha[b] = 0.0; d.done[b] = 1; atomicOr(&d.status[b], 16); }
      else { d.alpha[b] = an; atomicAdd(&d.counters[0], 1); }
    }
  }
}

// ------------------------------------------------------------------------------------------------ K6: take the step, finish the policy
template <int NJ>
__global__ void k_update(Dev d) {   // one thread per (instance, node, component): coalesced x += alpha dx, u += alpha du, uff += alpha kappa
  using R = RDims<NJ>;
  constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t node = gid / NX; const int i = (int)(gid % NX);
  const int b = (int)(node / d.NS), k = (int)(node % d.NS);
  if (b >= d.B) return;
  const int n = d.n_nodes[b];
  if (k >= n) return;
  const size_t nb = (size_t)b * d.NS;
  const double al = d.alpha[b];
  d.s_x[(nb + k) * NX + i] += al * d.dx[(nb + k) * NX + i];
  if (i < NU && k < n - 1 && d.node_ev[nb + k] != 1) {
    d.s_u[(nb + k) * NU + i] += al * d.du[(nb + k) * NU + i];
    d.s_uff[(nb + k) * NU + i] += al * d.ric[(nb + k) * R::KREC + R::K_KAP + i];
  }
}
// event nodes and the terminal node copy input / feedforward / gain of the previous node ([UPSTREAM] toPrimalSolution)
template <int NJ>
__global__ void k_policy_fill(Dev d) {
  constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU;
  const int b = blockIdx.x;
  const int n = d.n_nodes[b];
  const size_t nb = (size_t)b * d.NS;
  for (int k = 1; k < n; ++k) {
    const bool copy = (k == n - 1) || d.node_ev[nb + k] == 1;
    if (!copy) continue;
    for (int i = threadIdx.x; i < NU; i += blockDim.x) { d.s_u[(nb + k) * NU + i] = d.s_u[(nb + k - 1) * NU + i]; d.s_uff[(nb + k) * NU + i] = d.s_uff[(nb + k - 1) * NU + i]; }
    for (int i = threadIdx.x; i < NU * NX; i += blockDim.x) d.s_K[(nb + k) * (size_t)(NU * NX) + i] = d.s_K[(nb + k - 1) * (size_t)(NU * NX) + i];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ batched policy evaluation
// [UPSTREAM] MPC_MRT_Interface::evaluatePolicy + LinearController::computeInput (linear interpolation of uff and K in time)
template <int NJ>
__global__ void k_evaluate_policy(int B, int NS, int ME, const int* n_nodes, const double* times, const double* sx, const double* suff, const double* sK,
                                  const int* n_ev, const double* ev_t, const int* ev_mode, const double* tq, const double* xq, double* xo, double* uo, int* mo) {
  constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU;
  const int b = blockIdx.x;
  if (b >= B) return;
  const size_t nb = (size_t)b * NS;
  const int n = n_nodes[b];
  int idx; double al; time_segment(times + nb, n, tq[b], idx, al);
  const int i1 = min(idx + 1, n - 1);
  for (int i = threadIdx.x; i < NX; i += blockDim.x) xo[(size_t)b * NX + i] = al * sx[(nb + idx) * NX + i] + (1.0 - al) * sx[(nb + i1) * NX + i];
  for (int r = threadIdx.x; r < NU; r += blockDim.x) {
    double a = al * suff[(nb + idx) * NU + r] + (1.0 - al) * suff[(nb + i1) * NU + r];
    const double* K0 = sK + (nb + idx) * (size_t)(NU * NX) + r * NX; const double* K1 = sK + (nb + i1) * (size_t)(NU * NX) + r * NX;
    for (int c = 0; c < NX; ++c) a += (al * K0[c] + (1.0 - al) * K1[c]) * xq[(size_t)b * NX + c];
    uo[(size_t)b * NU + r] = a;
  }
  if (threadIdx.x == 0) mo[b] = ev_mode[(size_t)b * (ME + 1) + lower_bound_d(ev_t + (size_t)b * ME, n_ev[b], tq[b])];
}

// ------------------------------------------------------------------------------------------------ observation / target helpers (device-resident drivers)
// Next observation under a perfect model: t0 += dt, x0 = optimized state trajectory interpolated at the new time
// (what MRT_ROS_Dummy_Loop's policy rollout [UPSTREAM] provides between MPC ticks, without re-integration).
template <int NJ>
__global__ void k_shift_observations(int B, int NS, double dt, const int* n_nodes, const double* times, const double* sx, double* t0, double* x0) {
  constexpr int NX = Dims<NJ>::NX;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double t = t0[b] + dt;
  interp_vec(times + (size_t)b * NS, sx + (size_t)b * NS * NX, n_nodes[b], NX, t, x0 + (size_t)b * NX);
  t0[b] = t;
}
// TargetTrajectoriesPublisher::cmdVelToTargetTrajectories (bipedal_controllers/src/TargetTrajectoriesPublisher.cpp:76-99) on device
template <int NJ>
__global__ void k_cmd_vel_targets(int B, int TP, const double* t0, const double* x0, const double* cmd, double ttt, double com_height, const double* default_joints, double* tgt_t, double* tgt_x) {
  constexpr int NX = Dims<NJ>::NX;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* x = x0 + (size_t)b * NX; const double* c = cmd + (size_t)b * 4;
  double sz, cz, sy, cy, sx, cx;
  sincos(x[9], &sz, &cz); sincos(x[10], &sy, &cy); sincos(x[11], &sx, &cx);
  const double R[9] = {cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx, sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx, -sy, cy * sx, cy * cx};
  const double vr[3] = {R[0] * c[0] + R[1] * c[1] + R[2] * c[2], R[3] * c[0] + R[4] * c[1] + R[5] * c[2], R[6] * c[0] + R[7] * c[1] + R[8] * c[2]};
  double* s0 = tgt_x + (size_t)b * TP * NX; double* s1 = s0 + NX;
  for (int i = 0; i < 2 * NX; ++i) s0[i] = 0.0;
  s0[0] = s1[0] = vr[0]; s0[1] = s1[1] = vr[1]; s0[2] = s1[2] = vr[2];
  s0[6] = x[6]; s0[7] = x[7]; s0[8] = com_height; s0[9] = x[9];
  s1[6] = x[6] + vr[0] * ttt; s1[7] = x[7] + vr[1] * ttt; s1[8] = com_height; s1[9] = x[9] + c[3] * ttt;
  for (int j = 0; j < NJ; ++j) { s0[12 + j] = default_joints[j]; s1[12 + j] = default_joints[j]; }
  tgt_t[(size_t)b * TP] = t0[b]; tgt_t[(size_t)b * TP + 1] = t0[b] + ttt;
}

}  // namespace bmpc
