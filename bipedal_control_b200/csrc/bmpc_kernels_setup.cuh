// K0: time grid with event nodes, per-node references, warm start
// (part of bmpc_kernels.cuh: include that header, not this file)
#pragma once

namespace bmpc {

// ------------------------------------------------------------------------------------------------ gait bookkeeping on the device (one thread per instance)
struct GaitDev { int B, cap; int* n; double* ev; int* modes; GaitTemplateArrays* tmpl; };
__device__ __forceinline__ GaitView gait_view(const GaitDev& g, int b) { return GaitView{g.cap, g.n + b, g.ev + (size_t)b * g.cap, g.modes + (size_t)b * (g.cap + 1), g.tmpl + b}; }
// reference.info initialModeSchedule / defaultModeSequenceTemplate for instances [first, first + count) (BipedalRobotInterface.cpp:209-234)
__global__ void k_gait_init(GaitDev g, int first, int count, int n0, const double* ev0, const int* modes0, GaitTemplateArrays t0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const GaitView v = gait_view(g, first + i);
  *v.n = n0;
  for (int k = 0; k < n0; ++k) v.ev[k] = ev0[k];
  for (int k = 0; k <= n0; ++k) v.modes[k] = modes0[k];
  *v.tmpl = t0;
}
// GaitSchedule::insertModeSequenceTemplate for instances [first, first + count): status[] gets GaitStatus codes (0 = ok)
__global__ void k_gait_insert(GaitDev g, int first, int count, GaitTemplateArrays t, double start, double fin, double transition, int* result) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int rc = gait_insert(gait_view(g, first + i), t, start, fin, transition);
  if (rc != GAIT_OK) atomicMax(result, rc);
}
// SwitchedModelReferenceManager::modifyReferences (SwitchedModelReferenceManager.cpp:62-69): modeSchedule = gaitSchedule.getModeSchedule(t0 - T, tf + T),
// written straight into the solver's mode-schedule inputs
__global__ void k_gait_schedule(GaitDev g, Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  const GaitView v = gait_view(g, b);
  const double t0 = d.t0[b], tf = t0 + d.horizon;
  const int rc = gait_get_mode_schedule(v, t0 - d.horizon, tf + d.horizon);
  int n = *v.n;
  if (rc != GAIT_OK || n > d.ME) { atomicOr(&d.status[b], 32); n = n < d.ME ? n : d.ME; }
  int* nev = const_cast<int*>(d.n_ev); double* evt = const_cast<double*>(d.ev_t) + (size_t)b * d.ME; int* evm = const_cast<int*>(d.ev_mode) + (size_t)b * (d.ME + 1);
  nev[b] = n;
  for (int k = 0; k < n; ++k) evt[k] = v.ev[k];
  for (int k = 0; k <= n; ++k) evm[k] = v.modes[k];
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
// swing height reference of leg `leg` at time t: velocity (returned) and position (*zpos)
// (foot_planner/SwingTrajectoryPlanner.cpp:50-118, SplineCpg.cpp:38-60, CubicSpline.cpp:38-75; the position enters only with positionErrorGain != 0)
__device__ inline double swing_zref(const double* ev, const int* modes, int ne, int leg, double t, int* status, double* zpos) {
  const int np = ne + 1;
  const int p = lower_bound_d(ev, ne, t);
  int start = -1;
  for (int ip = p - 1; ip >= 0; --ip) if (leg_in_stance(modes[ip], leg)) { start = ip; break; }
  int fin = np - 1;
  for (int ip = p + 1; ip < np; ++ip) if (leg_in_stance(modes[ip], leg)) { fin = ip - 1; break; }
  *zpos = 0.0;
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
  *zpos = ((c3 * tn + c2) * tn + c1) * tn + p_a;
  return (3.0 * c3 * tn * tn + 2.0 * c2 * tn + c1) / dts;
}

// One thread per node for the scalar lookup chains (a warp-per-node variant was 2.7x slower: every lane of a warp would repeat them), which only
// DESCRIBE the three vectors a node gets (x_ref, warm-start state, warm-start input) as "alpha * a[] + (1 - alpha) * b[]" with two source rows;
// the CTA then writes the vectors of its 128 consecutive nodes together, element by element, so the 22-wide rows are read and written coalesced
// (the per-thread row loops of the first version ran at a quarter of the HBM bandwidth).
struct VecDesc { const double* a; const double* b; double al; double fz; int kind; };   // kind 0: none, 1: blend, 2: zero, 3: initializer input (fz, stance bits in al)
__device__ __forceinline__ VecDesc blend_desc(const double* ta, const double* data, int n, int dim, double t) {
  VecDesc v; v.fz = 0.0; v.kind = 1;
  if (n <= 1) { v.a = data; v.b = data; v.al = 1.0; return v; }
  int idx; double al; time_segment(ta, n, t, idx, al);
  v.a = data + (size_t)idx * dim; v.b = v.a + dim; v.al = al;
  return v;
}
template <int NJ>
__global__ void __launch_bounds__(128) k_node_setup(Dev d) {
  constexpr int NX = Dims<NJ>::NX, NU = Dims<NJ>::NU;
  __shared__ VecDesc sd[3][128];
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = gid / d.NS, k = gid % d.NS;
  VecDesc dr, dx, du; dr.kind = 0; dx.kind = 0; du.kind = 0;
  dr.a = dr.b = dx.a = dx.b = du.a = du.b = nullptr; dr.al = dx.al = du.al = 0.0; dr.fz = dx.fz = du.fz = 0.0;
  const int n = b < d.B ? d.n_nodes[b] : 0;
  if (k < n) {
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
        dr = blend_desc(d.tgt_t + (size_t)b * d.TP, d.tgt_x + (size_t)b * d.TP * NX, d.npts, NX, t);
        for (int leg = 0; leg < 2; ++leg) {   // per leg: reference height velocity, and position (used only with positionErrorGain != 0); terrain height 0 in stance
          double zp = 0.0;
          const double zv = leg_in_stance(mode, leg) ? 0.0 : swing_zref(ev, modes, ne, leg, t, &d.status[b], &zp);
          d.zref[(nb + k) * 4 + leg] = zv; d.zref[(nb + k) * 4 + 2 + leg] = zp;
        }
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
    if (j == 0) {
      const double tinit = nev[0] == 2 ? nt[0] + WEAK_EPS : nt[0];
      if (tinit < stateTill) dx = blend_desc(pt, px, pn, NX, tinit);
      else { dx.kind = 1; dx.a = dx.b = d.x0 + (size_t)b * NX; dx.al = 1.0; }
    } else dx = blend_desc(pt, px, pn, NX, stt[j - 1] + std_[j - 1]);
    // input of stage k
    if (k < N) {
      if (nev[k] == 1) du.kind = 2;
      else if (interval_uses_initializer(k)) {   // initialization/BipedalRobotInitializer.cpp:56-63 + common/utils.h:63-77
        const int mode = modes[lower_bound_d(ev, ne, stt[k])];
        const bool s0 = leg_in_stance(mode, 0), s1 = leg_in_stance(mode, 1);
        const int ns = 2 * (int(s0) + int(s1));
        du.kind = 3; du.fz = ns > 0 ? c_model.total_mass * 9.81 / ns : 0.0; du.al = (s0 ? 1.0 : 0.0) + (s1 ? 2.0 : 0.0);
      } else du = blend_desc(pt, pu, pn, NU, stt[k]);
    }
  }
  sd[0][threadIdx.x] = dr; sd[1][threadIdx.x] = dx; sd[2][threadIdx.x] = du;
  __syncthreads();
  // ---- the CTA's 128 nodes are consecutive rows of xref / s_x / s_u: element e of the block = (node e / dim, component e % dim)
  const size_t row0 = (size_t)blockIdx.x * blockDim.x;
  static_assert(NX == NU, "one element loop serves the three vectors");
  for (int e = threadIdx.x; e < 128 * NX; e += 128) {
    const int nd = e / NX, c = e - nd * NX;
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const VecDesc q = sd[v][nd];
      if (q.kind == 0) continue;
      double* out = (v == 0 ? d.xref : (v == 1 ? d.s_x : d.s_u)) + (row0 + nd) * NX + c;
      double val;
      if (q.kind == 1) val = q.al * q.a[c] + (1.0 - q.al) * q.b[c];
      else if (q.kind == 2) val = 0.0;
      else { const int st = (int)q.al; val = ((c == 2 || c == 5) && (st & 1)) || ((c == 8 || c == 11) && (st & 2)) ? q.fz : 0.0; }
      *out = val;
    }
  }
}

}  // namespace bmpc
