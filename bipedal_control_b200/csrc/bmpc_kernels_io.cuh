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
