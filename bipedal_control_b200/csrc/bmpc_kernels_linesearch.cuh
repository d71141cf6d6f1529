// K4-K6: line-search trial evaluation, filter acceptance, step update, policy fill
// (part of bmpc_kernels.cuh: include that header, not this file)
#pragma once

namespace bmpc {

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
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // broadcast: lets the compiler treat the warp index as warp-uniform
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

}  // namespace bmpc
