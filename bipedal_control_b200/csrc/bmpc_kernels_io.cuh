// batched policy evaluation and the device-resident observation / target helpers
// (part of bmpc_kernels.cuh: include that header, not this file)
#pragma once

namespace bmpc {

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
  if (n < 1) {   // no policy (the instance was reset, or its only solve so far failed): state echoed, zero input
    for (int i = threadIdx.x; i < NX; i += blockDim.x) xo[(size_t)b * NX + i] = xq[(size_t)b * NX + i];
    for (int r = threadIdx.x; r < NU; r += blockDim.x) uo[(size_t)b * NU + r] = 0.0;
    if (threadIdx.x == 0) mo[b] = ev_mode[(size_t)b * (ME + 1) + lower_bound_d(ev_t + (size_t)b * ME, n_ev[b], tq[b])];
    return;
  }
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
  if (n_nodes[b] < 1) { t0[b] = t; return; }   // no policy: the observation only advances in time
  interp_vec(times + (size_t)b * NS, sx + (size_t)b * NS * NX, n_nodes[b], NX, t, x0 + (size_t)b * NX);
  t0[b] = t;
}
// ------------------------------------------------------------------------------------------------ feedback-policy rollout between MPC ticks
// [UPSTREAM MRT_BASE::rolloutPolicy -> TimeTriggeredRollout::run] (BipedalController.cpp:322 installs the interface's rollout; settings
// task.info:159-167: ODE45, AbsTolODE 1e-5, RelTolODE 1e-3, timeStep 0.015): the closed loop xdot = f(x, uff(t) + K(t) x) is integrated from the
// observation time over `substeps` consecutive MRT periods.  One warp per instance; vectors are distributed (lane i owns component i of x and of the
// seven Dormand-Prince stages), the evaluation point and the input are exchanged through shared memory, every lane evaluates the flow map
// (model_values) redundantly.  Restatement of boost::odeint integrate_adaptive(make_controlled<runge_kutta_dopri5>(abs, rel), ...) and of
// RolloutBase::findActiveModesTimeInterval (sub-intervals split at the policy's event times, each started weakEpsilon late): see the oracle
// (oracle/oracle_solver.hpp, Solver::rolloutPolicy), which this kernel is checked against.
struct RolloutSmem { double xs[24], us[24], fs[12]; };

template <int NJ>
__device__ __forceinline__ double closed_loop_flow(int lane, RolloutSmem& sm, int n, const double* __restrict__ times, const double* __restrict__ suff, const double* __restrict__ sK,
                                                   double t, double xt_l) {
  constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU;
  __syncwarp();
  if (lane < NX) sm.xs[lane] = xt_l;
  __syncwarp();
  double xb[12], qj[NJ], uf[12], qd[NJ];
#pragma unroll
  for (int i = 0; i < 12; ++i) xb[i] = sm.xs[i];
#pragma unroll
  for (int j = 0; j < NJ; ++j) qj[j] = sm.xs[12 + j];
  // LinearController::computeInput: u = uff(t) + K(t) x, linear interpolation in time (row `lane`)
  int idx; double al; time_segment(times, n, t, idx, al);
  const int i1 = min(idx + 1, n - 1);
  double u_l = 0.0;
  if (lane < NU) {
    const double* K0 = sK + (size_t)idx * (NU * NX) + lane * NX; const double* K1 = sK + (size_t)i1 * (NU * NX) + lane * NX;
    double a0 = suff[(size_t)idx * NU + lane], a1 = suff[(size_t)i1 * NU + lane];
#pragma unroll
    for (int c = 0; c < 12; ++c) { a0 += K0[c] * xb[c]; a1 += K1[c] * xb[c]; }
#pragma unroll
    for (int c = 0; c < NJ; ++c) { a0 += K0[12 + c] * qj[c]; a1 += K1[12 + c] * qj[c]; }
    u_l = al * a0 + (1.0 - al) * a1;
    sm.us[lane] = u_l;
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 12; ++i) uf[i] = sm.us[i];
#pragma unroll
  for (int j = 0; j < NJ; ++j) qd[j] = sm.us[12 + j];
  double f[12]; v3 vc[NCON];
  model_values<NJ>(xb, qj, uf, qd, f, vc);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 12; ++i) sm.fs[i] = f[i];
  }
  __syncwarp();
  return lane < 12 ? sm.fs[lane] : u_l;   // rows 12.. of the flow map are the joint velocities = inputs 12..
}

template <int NJ>
__global__ void __launch_bounds__(128) k_rollout(int B, int NS, int ME, const int* __restrict__ n_nodes, const double* __restrict__ times_all, const double* __restrict__ suff_all,
                                                 const double* __restrict__ sK_all, const int* __restrict__ n_ev, const double* __restrict__ ev_t_all, double* t0, double* x0,
                                                 double time_step, int substeps, double abs_tol, double rel_tol, double dt_init, int* status) {
  constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU;
  __shared__ RolloutSmem smem[4];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= B) return;
  RolloutSmem& sm = smem[warp];
  const int n = n_nodes[b];
  const double tstart = t0[b];
  if (n < 1) { if (lane == 0) t0[b] = tstart + time_step; return; }   // no policy: the observation only advances in time
  const size_t nb = (size_t)b * NS;
  const double* times = times_all + nb; const double* suff = suff_all + nb * NU; const double* sK = sK_all + nb * (size_t)(NU * NX);
  const double* ev = ev_t_all + (size_t)b * ME; const int ne = n_ev[b];
  const double a21 = 1.0 / 5, a31 = 3.0 / 40, a32 = 9.0 / 40, a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9, a51 = 19372.0 / 6561, a52 = -25360.0 / 2187,
               a53 = 64448.0 / 6561, a54 = -212.0 / 729, a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176, a65 = -5103.0 / 18656,
               c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
  const double dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640, dc5 = c5 + 92097.0 / 339200, dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;
  double x = lane < NX ? x0[(size_t)b * NX + lane] : 0.0;
  const double hstep = time_step / substeps;
  bool failed = false;
  for (int ss = 0; ss < substeps && !failed; ++ss) {
    const double ta = tstart + ss * hstep, tf = ta + hstep;
    // sub-intervals: [ta, e1], [e1, e2], ..., [ek, tf] for the events ta < e < tf, each started WEAK_EPS late
    double seg_begin = ta;
    int ie = 0; while (ie < ne && !(ev[ie] > ta)) ++ie;   // upper_bound(events, ta)
    for (;;) {
      const bool has_ev = ie < ne && ev[ie] < tf;
      const double seg_end = has_ev ? ev[ie] : tf;
      const double begin = fmin(seg_begin + WEAK_EPS, seg_end);
      if (seg_end > begin) {
        double t = begin, dt = dt_init;
        double k1 = closed_loop_flow<NJ>(lane, sm, n, times, suff, sK, t, x);
        while (t < seg_end && (seg_end - t) > 1e-15 * fmax(1.0, fabs(seg_end))) {
          if (t + dt > seg_end) dt = seg_end - t;
          int trials = 0; bool ok = false;
          while (!ok && trials < 500) {
            ++trials;
            const double k2 = closed_loop_flow<NJ>(lane, sm, n, times, suff, sK, t + dt * (1.0 / 5), x + dt * a21 * k1);
            const double k3 = closed_loop_flow<NJ>(lane, sm, n, times, suff, sK, t + dt * (3.0 / 10), x + dt * (a31 * k1 + a32 * k2));
            const double k4 = closed_loop_flow<NJ>(lane, sm, n, times, suff, sK, t + dt * (4.0 / 5), x + dt * (a41 * k1 + a42 * k2 + a43 * k3));
            const double k5 = closed_loop_flow<NJ>(lane, sm, n, times, suff, sK, t + dt * (8.0 / 9), x + dt * (a51 * k1 + a52 * k2 + a53 * k3 + a54 * k4));
            const double k6 = closed_loop_flow<NJ>(lane, sm, n, times, suff, sK, t + dt, x + dt * (a61 * k1 + a62 * k2 + a63 * k3 + a64 * k4 + a65 * k5));
            const double xn = x + dt * (c1 * k1 + c3 * k3 + c4 * k4 + c5 * k5 + c6 * k6);
            const double k7 = closed_loop_flow<NJ>(lane, sm, n, times, suff, sK, t + dt, xn);
            const double xe = dt * (dc1 * k1 + dc3 * k3 + dc4 * k4 + dc5 * k5 + dc6 * k6 + dc7 * k7);
            double err = lane < NX ? fabs(xe) / (abs_tol + rel_tol * (fabs(x) + fabs(dt) * fabs(k1))) : 0.0;
            const bool bad = !(err == err);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) err = fmax(err, __shfl_xor_sync(0xffffffffu, err, o));
            if (__any_sync(0xffffffffu, bad)) { failed = true; break; }
            if (err > 1.0) { dt *= fmax(0.9 * pow(err, -1.0 / 3.0), 0.2); continue; }
            ok = true; t += dt; x = xn; k1 = k7;
            if (err < 0.5) { err = fmax(pow(5.0, -5.0), err); dt *= 0.9 * pow(err, -1.0 / 5.0); }
          }
          if (!ok) { failed = true; break; }
        }
      }
      if (failed || !has_ev) break;
      seg_begin = seg_end; ++ie;
    }
  }
  if (failed) { if (lane == 0) atomicOr(&status[b], 8); return; }   // observation left untouched
  if (lane < NX) x0[(size_t)b * NX + lane] = x;
  if (lane == 0) t0[b] = tstart + time_step;
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
