// K4-K6: line-search trial evaluation, filter acceptance, step update, policy fill
// (part of bmpc_kernels.cuh: include that header, not this file)
#pragma once

namespace bmpc {

// ------------------------------------------------------------------------------------------------ K4-K6: filter line search + step + policy completion
// ONE CTA PER INSTANCE: thread = stage.  The backtracking loop of FilterLinesearch runs inside the CTA (instances are independent, so no
// global synchronisation, no host round trip, no per-trial launches): every trial evaluates the RK2 defect, the cost and the constraint
// violation at (x + alpha dx, u + alpha du) for all stages of the instance on the streaming, register-only flow map (model_values: no per-joint
// arrays, no local memory), reduces them over the block and lets thread 0 take the accept / halve decision.  The accepted step size is applied by
// the wide kernels below (k_update: x += alpha dx, u += alpha du, uff = uff0 + alpha kappa; k_policy_fill: [UPSTREAM] toPrimalSolution).
// [UPSTREAM] FilterLinesearch::acceptStep (g_max, g_min: task.info:72-73; gamma_c 1e-6, armijoFactor 1e-4, alpha_decay 0.5, alpha_min 1e-4)
//
// Instances whose solve failed numerically (Riccati lost positive definiteness, rank anomaly of the constraint Jacobian, NaN) store nothing of
// this tick: they keep their previous policy, or none (n_nodes = 0) if there is no previous one, so the next tick starts from valid data (upstream
// throws in these cases and the controller stops, BipedalController.cpp:344-348; here bmpc_advance / bmpc_synchronize return BMPC_ERR_NUMERIC and
// the other instances of the batch are unaffected).
constexpr int FAIL_MASK = 1 | 2 | 8;
constexpr int LS_THREADS = 128;

// performance of one stage at step size al: out = {dt cost, dt |defect|^2, dt |equality constraints|^2}
template <int NJ>
__device__ __forceinline__ void stage_trial(const Dev& d, size_t nb, int k, double al, double (&out)[3]) {
  using D = Dims<NJ>;
  constexpr int NX = D::NX, NU = D::NU;
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
  double k1[12], k2[12]; v3 vc[NCON], vc2[NCON], pc[NCON];
  const bool pg = M.gain != 0.0;   // positionErrorGain: the z rows also see gain * (p_z - z_ref)
  model_values<NJ>(xb, qj, uf, qd, k1, vc, pg ? pc : nullptr);
  double peq = 0.0;
#pragma unroll
  for (int c = 0; c < NCON; ++c) {   // ZeroVelocityConstraintCppAd / NormalVelocityConstraintCppAd + ZeroForceConstraint, incl. positionErrorGain
    if ((c / 2 == 0) ? st0 : st1) { const double ez = vc[c].z + (pg ? M.gain * pc[c].z : 0.0); peq += vc[c].x * vc[c].x + vc[c].y * vc[c].y + ez * ez; }
    else {
      const double ev = vc[c].z - d.zref[(nb + k) * 4 + c / 2] + (pg ? M.gain * (pc[c].z - d.zref[(nb + k) * 4 + 2 + c / 2]) : 0.0);
      peq += ev * ev + uf[3 * c] * uf[3 * c] + uf[3 * c + 1] * uf[3 * c + 1] + uf[3 * c + 2] * uf[3 * c + 2];
    }
  }
  double xb2[12], qj2[NJ];
#pragma unroll
  for (int i = 0; i < 12; ++i) xb2[i] = xb[i] + dt * k1[i];
#pragma unroll
  for (int j = 0; j < NJ; ++j) qj2[j] = qj[j] + dt * qd[j];
  model_values<NJ>(xb2, qj2, uf, qd, k2, vc2, nullptr);
#pragma unroll
  for (int i = 0; i < 12; ++i) { const double e = xb[i] + 0.5 * dt * (k1[i] + k2[i]) - (gx[NX + i] + al * gdx[NX + i]); sdef += e * e; }
  out[0] = dt * cost; out[1] = dt * sdef; out[2] = dt * peq;
}

template <int NJ>
__global__ void __launch_bounds__(LS_THREADS, LS2_BLOCKS) k_linesearch(Dev d) {
  constexpr int NX = Dims<NJ>::NX, NWARP = LS_THREADS / 32;
  __shared__ double sred[NWARP][3];
  __shared__ double s_alpha;
  __shared__ int s_done;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = d.n_nodes[b], N = n - 1;
  const size_t nb = (size_t)b * d.NS;
  double* pf = d.perf + (size_t)b * 8;
  double al = 1.0;
  int trials = 0;
  const bool pre_fail = (d.status[b] & FAIL_MASK) != 0;
  for (;;) {
    double acc3[3] = {0.0, 0.0, 0.0};
    for (int k = tid; k < N; k += LS_THREADS) {
      double o[3]; stage_trial<NJ>(d, nb, k, al, o);
      acc3[0] += o[0]; acc3[1] += o[1]; acc3[2] += o[2];
    }
    if (tid < NX) { const double e = d.x0[(size_t)b * NX + tid] - (d.s_x[nb * NX + tid] + al * d.dx[nb * NX + tid]); acc3[1] += e * e; }   // initial-state defect
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { acc3[0] += __shfl_xor_sync(0xffffffffu, acc3[0], o); acc3[1] += __shfl_xor_sync(0xffffffffu, acc3[1], o); acc3[2] += __shfl_xor_sync(0xffffffffu, acc3[2], o); }
    if (lane == 0) { sred[wid][0] = acc3[0]; sred[wid][1] = acc3[1]; sred[wid][2] = acc3[2]; }
    __syncthreads();
    ++trials;
    if (tid == 0) {
      double pc = 0.0, pd = 0.0, pe = 0.0;
#pragma unroll
      for (int w = 0; w < NWARP; ++w) { pc += sred[w][0]; pd += sred[w][1]; pe += sred[w][2]; }
      const double th0 = sqrt(pf[1] + pf[2]), th = sqrt(pd + pe);
      const double gamma_c = 1e-6, armijoFactor = 1e-4, alpha_decay = 0.5, alpha_min = 1e-4;
      const double armijo = pf[7];
      bool acc;
      if (th > c_model.g_max) acc = th < (1.0 - gamma_c) * th0;
      else if (th < c_model.g_min && th0 < c_model.g_min && armijo < 0.0) acc = pc < pf[0] + armijoFactor * al * armijo;
      else acc = (pc < pf[0] - gamma_c * th0) || (th < (1.0 - gamma_c) * th0);
      const bool nan_ = !(pc == pc) || !(pd == pd) || !(pe == pe) || !(armijo == armijo);
      int done = 0; double anext = al;
      if (nan_ || pre_fail) { pf[3] = pf[0]; pf[4] = pf[1]; pf[5] = pf[2]; pf[6] = 0.0; anext = 0.0; done = 1; if (nan_) atomicOr(&d.status[b], 8); }
      else if (acc) { pf[3] = pc; pf[4] = pd; pf[5] = pe; pf[6] = al; done = 1; }
      else {
        const double an = al * alpha_decay;
        const bool small = an * d.norms[2 * b] < c_model.delta_tol && an * d.norms[2 * b + 1] < c_model.delta_tol;
        if (small || an < alpha_min) { pf[3] = pf[0]; pf[4] = pf[1]; pf[5] = pf[2]; pf[6] = 0.0; anext = 0.0; done = 1; atomicOr(&d.status[b], 16); }
        else anext = an;
      }
      s_alpha = anext; s_done = done;
    }
    __syncthreads();
    al = s_alpha;
    if (s_done) break;
  }
  if (tid == 0) {
    d.alpha[b] = al;
    atomicAdd(&d.counters[CNT_TRIALS], trials); atomicMax(&d.counters[CNT_MAXTRIALS], trials);
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
  const double al = d.alpha[b];
  if (al == 0.0 || (d.status[b] & FAIL_MASK) != 0) return;   // rejected step / failed instance: nothing is taken
  const size_t nb = (size_t)b * d.NS;
  d.s_x[(nb + k) * NX + i] += al * d.dx[(nb + k) * NX + i];
  if (i < NU && k < n - 1 && d.node_ev[nb + k] != 1) {
    d.s_u[(nb + k) * NU + i] += al * d.du[(nb + k) * NU + i];
    d.s_uff[(nb + k) * NU + i] += al * d.ric[(nb + k) * R::KREC + R::K_KAP + i];
  }
}
// once per tick, one CTA per instance: event nodes and the terminal node copy input / feedforward / gain of the previous node
// ([UPSTREAM] toPrimalSolution); a failed instance gets its previous policy back (see above); per-tick status summary for the host
template <int NJ>
__global__ void k_policy_fill(Dev d) {
  constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const size_t nb = (size_t)b * d.NS;
  const int st = d.status[b];
  if (tid == 0 && st) { atomicOr(&d.counters[CNT_STATUS], st); if (st & FAIL_MASK) atomicAdd(&d.counters[CNT_FAIL], 1); }
  if ((st & FAIL_MASK) != 0) {
    // the instance keeps its previous policy (time grid included), or has none (n_nodes = 0) if there is no previous one: the next tick then
    // cold-starts it from the initializer.  Nothing computed by this tick is stored.
    const int pn = d.p_n ? d.p_n[b] : 0;
    for (int i = tid; i < pn; i += nt) { d.node_t[nb + i] = d.p_t[nb + i]; d.node_ev[nb + i] = d.p_ev[nb + i]; }
    for (int i = tid; i < pn * NX; i += nt) d.s_x[nb * NX + i] = d.p_x[nb * NX + i];
    for (int i = tid; i < pn * NU; i += nt) { d.s_u[nb * NU + i] = d.p_u[nb * NU + i]; d.s_uff[nb * NU + i] = d.p_uff[nb * NU + i]; }
    for (size_t i = tid; i < (size_t)pn * NU * NX; i += nt) d.s_K[nb * (size_t)(NU * NX) + i] = d.p_K[nb * (size_t)(NU * NX) + i];
    __syncthreads();
    if (tid == 0) d.n_nodes[b] = pn;
    return;
  }
  const int n = d.n_nodes[b];
  for (int k = 1; k < n; ++k) {
    const bool copy = (k == n - 1) || d.node_ev[nb + k] == 1;
    if (!copy) continue;
    for (int i = tid; i < NU; i += nt) { d.s_u[(nb + k) * NU + i] = d.s_u[(nb + k - 1) * NU + i]; d.s_uff[(nb + k) * NU + i] = d.s_uff[(nb + k - 1) * NU + i]; }
    for (int i = tid; i < NU * NX; i += nt) d.s_K[(nb + k) * (size_t)(NU * NX) + i] = d.s_K[(nb + k - 1) * (size_t)(NU * NX) + i];
    __syncthreads();
  }
}

}  // namespace bmpc
