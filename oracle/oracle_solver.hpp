// TEST INFRASTRUCTURE - CPU oracle (see oracle_model.hpp header).  PARITY UNPINNED.
//
// oracle_solver.hpp : gait / swing / reference logic of ocs2_bipedal_robot and one multiple-shooting SQP
// iteration ([UPSTREAM] ocs2_sqp::SqpSolver as configured by task.info:66-83), dense FP64.
#pragma once
#include <algorithm>
#include <cassert>
#include <limits>
#include <memory>
#include "oracle_model.hpp"

namespace orc {

constexpr double WEAK_EPS = 1e-6;                                     // [UPSTREAM] numeric_traits::weakEpsilon
constexpr double LIMIT_EPS = std::numeric_limits<double>::epsilon();  // [UPSTREAM] numeric_traits::limitEpsilon
enum Mode { FLY = 0, LF = 1, RF = 2, STANCE = 3 };                    // gait/MotionPhaseDefinition.h:47-52

// gait/MotionPhaseDefinition.h:57-76
inline void mode_to_contact_flags(int mode, bool* f) {
  f[0] = f[1] = (mode == LF || mode == STANCE);
  f[2] = f[3] = (mode == RF || mode == STANCE);
}

// [UPSTREAM] lookup::findIndexInTimeArray: index of first element >= t
inline int find_index(const std::vector<double>& ta, double t) { return int(std::lower_bound(ta.begin(), ta.end(), t) - ta.begin()); }
// [UPSTREAM] lookup::findIntervalInTimeArray: i with ta[i] < t <= ta[i+1]; first interval closed on the left
inline int find_interval(const std::vector<double>& ta, double t) {
  if (ta.empty()) return 0;
  const int idx = find_index(ta, t);
  if (idx == 0 && t == ta.front()) return 0;
  return idx - 1;
}

struct ModeSchedule {
  std::vector<double> eventTimes;
  std::vector<int> modeSequence;
  int modeAtTime(double t) const { return modeSequence[find_index(eventTimes, t)]; }  // [UPSTREAM] ModeSchedule::modeAtTime
};

// ---------------------------------------------------------------- gait/GaitSchedule.cpp:38-137
struct GaitSchedule {
  ModeSchedule ms;
  GaitTemplate tmpl;
  double phaseTransitionStanceTime = 0.4;

  void tile(double startTime, double finalTime) {  // GaitSchedule.cpp:107-137
    auto& et = ms.eventTimes; auto& seq = ms.modeSequence;
    const size_t n = tmpl.modes.size();
    if (n == 0) return;
    if (!et.empty() && startTime <= et.back()) throw std::runtime_error("The initial time for template-tiling is not greater than the last event time.");
    et.push_back(startTime);
    while (et.back() < finalTime) {
      for (size_t i = 0; i < n; ++i) { seq.push_back(tmpl.modes[i]); et.push_back(et.back() + (tmpl.times[i + 1] - tmpl.times[i])); }
    }
    seq.push_back(STANCE);
  }
  void insertModeSequenceTemplate(const GaitTemplate& t, double startTime, double finalTime) {  // GaitSchedule.cpp:46-73
    tmpl = t;
    auto& et = ms.eventTimes; auto& seq = ms.modeSequence;
    const size_t index = std::lower_bound(et.begin(), et.end(), startTime) - et.begin();
    if (index < et.size()) { et.erase(et.begin() + index, et.end()); seq.erase(seq.begin() + index + 1, seq.end()); }
    double pts = phaseTransitionStanceTime;
    if (!seq.empty() && seq.back() == STANCE) pts = 0.0;
    if (pts > 0.0) { et.push_back(startTime); seq.push_back(STANCE); }
    tile(startTime + pts, finalTime);
  }
  ModeSchedule getModeSchedule(double lower, double upper) {  // GaitSchedule.cpp:78-102
    auto& et = ms.eventTimes; auto& seq = ms.modeSequence;
    const size_t index = std::lower_bound(et.begin(), et.end(), lower) - et.begin();
    if (index > 0) {
      et.erase(et.begin(), et.begin() + index - 1);
      seq.erase(seq.begin(), seq.begin() + index - 1);
      seq.front() = STANCE;
    }
    const double tilingStart = et.empty() ? upper : et.back();
    et.erase(et.end() - 1, et.end());
    seq.erase(seq.end() - 1, seq.end());
    tile(tilingStart, upper);
    return ms;
  }
};

// ---------------------------------------------------------------- foot_planner/CubicSpline.cpp:38-107, SplineCpg.cpp:38-83
struct CubicSpline {
  double t0, t1, dt, c0, c1, c2, c3;
  CubicSpline() : t0(0), t1(1), dt(1), c0(0), c1(0), c2(0), c3(0) {}
  CubicSpline(double ts, double ps, double vs, double te, double pe, double ve) {
    t0 = ts; t1 = te; dt = te - ts;
    const double dp = pe - ps, dv = ve - vs;
    const double dc0 = 0.0, dc1 = vs, dc2 = -(3.0 * vs + dv), dc3 = (2.0 * vs + dv);
    c0 = dc0 * dt + ps; c1 = dc1 * dt; c2 = dc2 * dt + 3.0 * dp; c3 = dc3 * dt - 2.0 * dp;
  }
  double position(double t) const { const double tn = (t - t0) / dt; return c3 * tn * tn * tn + c2 * tn * tn + c1 * tn + c0; }
  double velocity(double t) const { const double tn = (t - t0) / dt; return (3.0 * c3 * tn * tn + 2.0 * c2 * tn + c1) / dt; }
};
struct SplineCpg {
  double midTime; CubicSpline left, right;
  SplineCpg() : midTime(0.5) {}
  SplineCpg(double tl, double pl, double vl, double midHeight, double tt, double pt, double vt)
      : midTime((tl + tt) / 2), left(tl, pl, vl, midTime, midHeight, 0.0), right(midTime, midHeight, 0.0, tt, pt, vt) {}
  double position(double t) const { return t < midTime ? left.position(t) : right.position(t); }
  double velocity(double t) const { return t < midTime ? left.velocity(t) : right.velocity(t); }
};

// ---------------------------------------------------------------- foot_planner/SwingTrajectoryPlanner.cpp:50-219
struct SwingPlanner {
  double liftOffVelocity, touchDownVelocity, swingHeight, swingTimeScale;
  std::vector<SplineCpg> traj[NC];
  std::vector<double> events;
  void update(const ModeSchedule& ms, double terrainHeight) {
    const auto& seq = ms.modeSequence; const auto& et = ms.eventTimes;
    const int np = (int)seq.size();
    for (int leg = 0; leg < NC; ++leg) {
      std::vector<bool> flags(np);
      for (int p = 0; p < np; ++p) { bool f[NC]; mode_to_contact_flags(seq[p], f); flags[p] = f[leg]; }
      traj[leg].clear();
      for (int p = 0; p < np; ++p) {
        if (!flags[p]) {
          int start = -1; for (int ip = p - 1; ip >= 0; --ip) if (flags[ip]) { start = ip; break; }
          int fin = np - 1; for (int ip = p + 1; ip < np; ++ip) if (flags[ip]) { fin = ip - 1; break; }
          if (start < 0) throw std::runtime_error("The time of take-off for the first swing of the EE with ID " + std::to_string(leg) + " is not defined.");
          if (fin >= np - 1) throw std::runtime_error("The time of touch-down for the last swing of the EE with ID " + std::to_string(leg) + " is not defined.");
          const double ts = et[start], tf = et[fin];
          const double scaling = std::min(1.0, (tf - ts) / swingTimeScale);
          const double mid = std::min(terrainHeight, terrainHeight) + scaling * swingHeight;
          traj[leg].emplace_back(ts, terrainHeight, scaling * liftOffVelocity, mid, tf, terrainHeight, scaling * touchDownVelocity);
        } else {
          traj[leg].emplace_back(0.0, terrainHeight, 0.0, terrainHeight, 1.0, terrainHeight, 0.0);
        }
      }
    }
    events = et;
  }
  double zvel(int leg, double t) const { return traj[leg][find_index(events, t)].velocity(t); }
  double zpos(int leg, double t) const { return traj[leg][find_index(events, t)].position(t); }
};

// ---------------------------------------------------------------- [UPSTREAM] LinearInterpolation::timeSegment / interpolate
inline std::pair<int, double> time_segment(double t, const std::vector<double>& ta) {
  const int index = find_interval(ta, t);
  const int last = (int)ta.size() - 1;
  if (index >= 0) {
    if (index < last) {
      const double len = ta[index + 1] - ta[index];
      const double till = ta[index + 1] - t;
      if (len > 2.0 * LIMIT_EPS) return {index, till / len};
      return (till < 0.5 * len) ? std::make_pair(index, 0.0) : std::make_pair(index, 1.0);
    }
    return {std::max(last - 1, 0), 0.0};
  }
  return {0, 1.0};
}
inline void interpolate(double t, const std::vector<double>& ta, const std::vector<std::vector<double>>& data, std::vector<double>& out) {
  if (data.size() == 1 || ta.size() <= 1) { out = data.front(); return; }
  auto seg = time_segment(t, ta);
  const auto& a = data[seg.first]; const auto& b = data[seg.first + 1];
  out.resize(a.size());
  for (size_t i = 0; i < a.size(); ++i) out[i] = seg.second * a[i] + (1.0 - seg.second) * b[i];
}

struct TargetTrajectories {
  std::vector<double> times; std::vector<std::vector<double>> states;
  void desiredState(double t, std::vector<double>& out) const { interpolate(t, times, states, out); }
};

// bipedal_controllers/src/TargetTrajectoriesPublisher.cpp:41-99 (cmdVelToTargetTrajectories)
inline TargetTrajectories cmd_vel_to_target(const Model& M, double t_obs, const double* x_obs, const double cmd[4], double timeToTarget) {
  const double* pose = x_obs + 6;
  M3<double> R = euler_zyx<double>(pose[3], pose[4], pose[5]);
  const V3<double> vr = R * V3<double>(cmd[0], cmd[1], cmd[2]);
  std::vector<double> s0(M.nx, 0.0), s1(M.nx, 0.0);
  double cur[6] = {pose[0], pose[1], M.com_height, pose[3], 0.0, 0.0};
  double tgt[6] = {pose[0] + vr.x * timeToTarget, pose[1] + vr.y * timeToTarget, M.com_height, pose[3] + cmd[3] * timeToTarget, 0.0, 0.0};
  for (int i = 0; i < 6; ++i) { s0[6 + i] = cur[i]; s1[6 + i] = tgt[i]; }
  for (int j = 0; j < M.nj; ++j) { s0[12 + j] = M.default_joint_state[j]; s1[12 + j] = M.default_joint_state[j]; }
  s0[0] = s1[0] = vr.x; s0[1] = s1[1] = vr.y; s0[2] = s1[2] = vr.z;
  TargetTrajectories tt; tt.times = {t_obs, t_obs + timeToTarget}; tt.states = {s0, s1};
  return tt;
}

// ---------------------------------------------------------------- [UPSTREAM] timeDiscretizationWithEvents
struct AnnotatedTime { double time; int event; };  // event: 0 None, 1 PreEvent, 2 PostEvent
inline std::vector<AnnotatedTime> time_discretization_with_events(double t0, double tf, double dt, const std::vector<double>& eventTimes) {
  const double dt_min = 10.0 * WEAK_EPS;
  std::vector<AnnotatedTime> td;
  td.push_back({t0, 0});
  size_t nextEvent = (size_t)find_index(eventTimes, t0);
  AnnotatedTime next = td.back();
  while (td.back().time < tf) {
    next.time = next.time + dt; next.event = 0;
    if (nextEvent < eventTimes.size() && next.time >= eventTimes[nextEvent]) { next.time = eventTimes[nextEvent]; next.event = 1; ++nextEvent; }
    if (next.time >= tf) { next.time = tf; next.event = 0; }
    if (next.time > td.back().time + dt_min) td.push_back(next); else td.back() = next;
    if (next.event == 1) { next.event = 2; td.push_back(next); }
  }
  return td;
}
inline double interval_start(const AnnotatedTime& a) { return a.event == 2 ? a.time + WEAK_EPS : a.time; }
inline double interval_end(const AnnotatedTime& a) { return a.event == 1 ? a.time - WEAK_EPS : a.time; }

// ---------------------------------------------------------------- dense helpers
struct Mat {
  int r = 0, c = 0; std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};
inline Mat matmul(const Mat& A, const Mat& B) { Mat C(A.r, B.c); for (int i = 0; i < A.r; ++i) for (int k = 0; k < A.c; ++k) { const double a = A(i, k); if (a == 0.0) continue; for (int j = 0; j < B.c; ++j) C(i, j) += a * B(k, j); } return C; }
inline Mat matmulT(const Mat& A, const Mat& B) { /* A^T B */ Mat C(A.c, B.c); for (int k = 0; k < A.r; ++k) for (int i = 0; i < A.c; ++i) { const double a = A(k, i); if (a == 0.0) continue; for (int j = 0; j < B.c; ++j) C(i, j) += a * B(k, j); } return C; }
inline std::vector<double> matvec(const Mat& A, const std::vector<double>& x) { std::vector<double> y(A.r, 0.0); for (int i = 0; i < A.r; ++i) { double s = 0; for (int j = 0; j < A.c; ++j) s += A(i, j) * x[j]; y[i] = s; } return y; }
inline std::vector<double> matTvec(const Mat& A, const std::vector<double>& x) { std::vector<double> y(A.c, 0.0); for (int i = 0; i < A.r; ++i) for (int j = 0; j < A.c; ++j) y[j] += A(i, j) * x[i]; return y; }
inline double dotv(const std::vector<double>& a, const std::vector<double>& b) { double s = 0; for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i]; return s; }

// In-place Cholesky (lower) of SPD n x n; returns false if not PD.
inline bool cholesky(Mat& G) {
  const int n = G.r;
  for (int j = 0; j < n; ++j) {
    double d = G(j, j); for (int k = 0; k < j; ++k) d -= G(j, k) * G(j, k);
    if (!(d > 0.0)) return false;
    d = std::sqrt(d); G(j, j) = d;
    for (int i = j + 1; i < n; ++i) { double s = G(i, j); for (int k = 0; k < j; ++k) s -= G(i, k) * G(j, k); G(i, j) = s / d; }
  }
  return true;
}
inline void chol_solve(const Mat& L, double* b /* n */) {
  const int n = L.r;
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L(i, k) * b[k]; b[i] = s / L(i, i); }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= L(k, i) * b[k]; b[i] = s / L(i, i); }
}

// ---------------------------------------------------------------- per-node LQ data (dense, un-projected and projected)
struct NodeLQ {
  int type = 0;        // 0 intermediate, 1 event (PreEvent node: nu = 0), 2 terminal
  int mode = STANCE; double t = 0, dt = 0;
  int nc_rows = 0, m = 0, rank = 0;   // equality rows, reduced input dimension, rank(D)
  Mat A, B; std::vector<double> b;            // dx+ = A dx + B du + b   (un-projected)
  Mat Q, R, P; std::vector<double> q, r; double c = 0;   // stage cost (already * dt)
  Mat C, D; std::vector<double> e;            // C dx + D du + e = 0
  Mat Px, Pu; std::vector<double> Pe;         // du = Pe + Px dx + Pu dut
  Mat At, Bt; std::vector<double> bt;         // projected dynamics
  Mat Qt, Rt, Pt; std::vector<double> qt, rt; double ct = 0;
  // performance contributions of this node at the linearisation point
  double perf_cost = 0, perf_dyn = 0, perf_eq = 0;
  // Riccati outputs
  Mat Kt; std::vector<double> kt; Mat K;      // reduced gain/feedforward, full gain K = Px + Pu Kt
};

struct Performance { double cost = 0, dynSSE = 0, eqSSE = 0; double merit() const { return cost; } double theta() const { return std::sqrt(dynSSE + eqSSE); } };

// ---------------------------------------------------------------- problem evaluation at one (t, x, u)
struct Problem {
  const Model* M = nullptr;
  ModeSchedule modeSchedule;
  SwingPlanner swing;
  TargetTrajectories target;

  // common/utils.h:63-77
  void weightCompensatingInput(int mode, std::vector<double>& u) const {
    bool f[NC]; mode_to_contact_flags(mode, f);
    int ns = 0; for (int i = 0; i < NC; ++i) ns += f[i];
    u.assign(M->nu, 0.0);
    if (ns > 0) { const double fz = M->total_mass * 9.81 / ns; for (int i = 0; i < NC; ++i) if (f[i]) u[3 * i + 2] = fz; }
  }

  // relaxed barrier [UPSTREAM RelaxedBarrierPenalty]; config BipedalRobotInterface.cpp:296-316
  void barrier(double h, double& p, double& dp, double& ddp) const {
    const double mu = M->barrier_mu, de = M->barrier_delta;
    if (h > de) { p = -mu * std::log(h); dp = -mu / h; ddp = mu / (h * h); }
    else { const double dh = (h - 2.0 * de) / de; p = mu * (-std::log(de) + 0.5 * dh * dh - 0.5); dp = mu * (h - 2.0 * de) / (de * de); ddp = mu / (de * de); }
  }

  // constraint/FrictionConeConstraint.cpp:129-166 (terrain rotation = identity, FrictionConeConstraint.h:143)
  void frictionCone(const double* F, double& h, double g[3], double H[3][3]) const {
    const double reg = M->fr_reg, mu = M->mu_f;
    const double fx2 = F[0] * F[0], fy2 = F[1] * F[1];
    const double ts = fx2 + fy2 + reg, tn = std::sqrt(ts), t32 = tn * ts;
    h = mu * (F[2] + M->fr_grip) - tn;
    g[0] = -F[0] / tn; g[1] = -F[1] / tn; g[2] = mu;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) H[i][j] = 0.0;
    H[0][0] = -(fy2 + reg) / t32; H[0][1] = H[1][0] = F[0] * F[1] / t32; H[1][1] = -(fx2 + reg) / t32;
  }

  // Stage cost value: tracking cost (cost/BipedalRobotQuadraticTrackingCost.h:57-63) + soft friction cones.
  double costValue(double t, int mode, const double* x, const double* u) const {
    const int nx = M->nx, nu = M->nu, nj = M->nj;
    std::vector<double> xr, un; target.desiredState(t, xr); weightCompensatingInput(mode, un);
    double c = 0;
    for (int i = 0; i < nx; ++i) { const double d = x[i] - xr[i]; c += 0.5 * M->Q_diag[i] * d * d; }
    std::vector<double> du(nu); for (int i = 0; i < nu; ++i) du[i] = u[i] - un[i];
    for (int i = 0; i < 12; ++i) c += 0.5 * M->R_force_diag[i] * du[i] * du[i];
    for (int i = 0; i < nj; ++i) for (int j = 0; j < nj; ++j) c += 0.5 * du[12 + i] * M->R_joint[i * nj + j] * du[12 + j];
    bool f[NC]; mode_to_contact_flags(mode, f);
    for (int k = 0; k < NC; ++k) if (f[k]) { double h, g[3], H[3][3], p, dp, ddp; frictionCone(u + 3 * k, h, g, H); barrier(h, p, dp, ddp); c += p; }
    return c;
  }

  // number of equality rows for a mode, in the stacking order of BipedalRobotInterface.cpp:187-191
  static int numEqRows(int mode) { bool f[NC]; mode_to_contact_flags(mode, f); int n = 0; for (int i = 0; i < NC; ++i) n += f[i] ? 3 : 4; return n; }

  // equality constraint values given contact positions / velocities (EndEffectorLinearConstraint.cpp:74-87,
  // ZeroForceConstraint.cpp:57-59, BipedalRobotPreComputation.cpp:71-80, BipedalRobotInterface.cpp:350-359)
  void eqValues(double t, int mode, const double* u, const V3<double>* pos, const V3<double>* vel, std::vector<double>& e) const {
    bool f[NC]; mode_to_contact_flags(mode, f);
    e.clear();
    const double gain = M->pos_err_gain;
    for (int i = 0; i < NC; ++i) {
      if (!f[i]) { e.push_back(u[3 * i]); e.push_back(u[3 * i + 1]); e.push_back(u[3 * i + 2]); }       // zeroForce
      if (f[i]) { e.push_back(vel[i].x); e.push_back(vel[i].y); e.push_back(vel[i].z + (gain != 0.0 ? gain * pos[i].z : 0.0)); }  // zeroVelocity
      if (!f[i]) {                                                                                      // normalVelocity
        double v = vel[i].z - swing.zvel(i, t);
        if (gain != 0.0) v += gain * (pos[i].z - swing.zpos(i, t));
        e.push_back(v);
      }
    }
  }
};

// Heun / RK2 value discretisation [UPSTREAM ocs2 integrator RK2]
inline void rk2_value(const Model& M, const double* x, const double* u, double dt, double* xn, V3<double>* pos, V3<double>* vel) {
  const int nx = M.nx;
  double k1[MAXX], k2[MAXX], xt[MAXX];
  flow_map<double>(M, x, u, k1, pos, vel);
  for (int i = 0; i < nx; ++i) xt[i] = x[i] + dt * k1[i];
  flow_map<double>(M, xt, u, k2);
  for (int i = 0; i < nx; ++i) xn[i] = x[i] + 0.5 * dt * (k1[i] + k2[i]);
}

constexpr int ND = MAXX + MAXU;
using DualN = Dual<ND>;

// Flow map + Jacobians + contact kinematics + Jacobians at (x,u) with forward-mode duals.
struct Lin { std::vector<double> f; Mat A, B; V3<double> pos[NC], vel[NC]; Mat dpdx[NC], dvdx[NC], dvdu[NC]; };
inline void linearize(const Model& M, const double* x, const double* u, Lin& L, bool with_contacts) {
  const int nx = M.nx, nu = M.nu;
  DualN xd[MAXX], ud[MAXU], fd[MAXX];
  for (int i = 0; i < nx; ++i) { xd[i] = DualN(x[i]); xd[i].d[i] = 1.0; }
  for (int i = 0; i < nu; ++i) { ud[i] = DualN(u[i]); ud[i].d[nx + i] = 1.0; }
  V3<DualN> pd[NC], vd[NC];
  flow_map<DualN>(M, xd, ud, fd, with_contacts ? pd : nullptr, with_contacts ? vd : nullptr);
  L.f.resize(nx); L.A = Mat(nx, nx); L.B = Mat(nx, nu);
  for (int i = 0; i < nx; ++i) { L.f[i] = fd[i].v; for (int j = 0; j < nx; ++j) L.A(i, j) = fd[i].d[j]; for (int j = 0; j < nu; ++j) L.B(i, j) = fd[i].d[nx + j]; }
  if (with_contacts) for (int c = 0; c < NC; ++c) {
    L.dpdx[c] = Mat(3, nx); L.dvdx[c] = Mat(3, nx); L.dvdu[c] = Mat(3, nu);
    for (int r = 0; r < 3; ++r) {
      L.pos[c][r] = pd[c][r].v; L.vel[c][r] = vd[c][r].v;
      for (int j = 0; j < nx; ++j) { L.dpdx[c](r, j) = pd[c][r].d[j]; L.dvdx[c](r, j) = vd[c][r].d[j]; }
      for (int j = 0; j < nu; ++j) L.dvdu[c](r, j) = vd[c][r].d[nx + j];
    }
  }
}

// ---------------------------------------------------------------- constraint projection
// du = Pe + Px dx + Pu dut with D Pu = 0, Px = -D^+ C, Pe = -D^+ e (D^+ = Moore-Penrose pseudo-inverse).
// Moore-Penrose alternative to upstream's projection (selectable: projection_mode() = 0, product option "projection_mode" 0).
// [UPSTREAM] luConstraintProjection uses Eigen::FullPivLU (project_constraints_fullpivlu below, the default); its particular solution
// differs from the min-norm one only inside null(D) (which the QP re-optimises) except for the rank-deficient stance-foot rows
// (SURVEY.md Appendix B.6).  Algorithm: row-space orthonormalisation by
// modified Gram-Schmidt in natural row order with a relative rank tolerance, D = T W, D^+ = W^T (T^T T)^-1 T^T;
// null-space basis by pivoted Gram-Schmidt of the unit vectors against W.
constexpr double RANK_TOL = 1e-9;
inline void project_constraints(const Mat& C, const Mat& D, const std::vector<double>& e, Mat& Px, Mat& Pu, std::vector<double>& Pe, int& rank) {
  const int nr = D.r, nu = D.c, nx = C.c;
  Mat W(nr, nu), T(nr, nr);
  std::vector<int> act(nr, 0);
  for (int k = 0; k < nr; ++k) {
    std::vector<double> row(nu); double n0 = 0;
    for (int j = 0; j < nu; ++j) { row[j] = D(k, j); n0 += row[j] * row[j]; }
    for (int pass = 0; pass < 2; ++pass)
      for (int l = 0; l < k; ++l) if (act[l]) {
        double cdot = 0; for (int j = 0; j < nu; ++j) cdot += row[j] * W(l, j);
        for (int j = 0; j < nu; ++j) row[j] -= cdot * W(l, j);
        T(k, l) += cdot;
      }
    double res = 0; for (int j = 0; j < nu; ++j) res += row[j] * row[j];
    if (n0 > 0.0 && res > RANK_TOL * RANK_TOL * n0) { act[k] = 1; const double nrm = std::sqrt(res); T(k, k) = nrm; for (int j = 0; j < nu; ++j) W(k, j) = row[j] / nrm; }
  }
  std::vector<int> idx; for (int k = 0; k < nr; ++k) if (act[k]) idx.push_back(k);
  rank = (int)idx.size();
  // Mm = T^T T on the active columns
  Mat Mm(rank, rank);
  for (int a = 0; a < rank; ++a) for (int b2 = 0; b2 < rank; ++b2) { double s = 0; for (int k = 0; k < nr; ++k) s += T(k, idx[a]) * T(k, idx[b2]); Mm(a, b2) = s; }
  if (rank > 0 && !cholesky(Mm)) throw std::runtime_error("[oracle] projection: T^T T not PD");
  auto pinv_apply = [&](const std::vector<double>& g, std::vector<double>& y) {  // y = D^+ g
    std::vector<double> z(rank, 0.0);
    for (int a = 0; a < rank; ++a) { double s = 0; for (int k = 0; k < nr; ++k) s += T(k, idx[a]) * g[k]; z[a] = s; }
    if (rank > 0) chol_solve(Mm, z.data());
    y.assign(nu, 0.0);
    for (int a = 0; a < rank; ++a) for (int j = 0; j < nu; ++j) y[j] += W(idx[a], j) * z[a];
  };
  Px = Mat(nu, nx); Pe.assign(nu, 0.0);
  std::vector<double> g(nr), y;
  for (int c = 0; c < nx; ++c) { for (int k = 0; k < nr; ++k) g[k] = C(k, c); pinv_apply(g, y); for (int j = 0; j < nu; ++j) Px(j, c) = -y[j]; }
  pinv_apply(e, y); for (int j = 0; j < nu; ++j) Pe[j] = -y[j];
  // null-space basis: pivoted Gram-Schmidt of e_0..e_{nu-1} against W
  const int m = nu - rank;
  Pu = Mat(nu, m);
  Mat cand(nu, nu);
  for (int i = 0; i < nu; ++i) {
    cand(i, i) = 1.0;
    for (int pass = 0; pass < 2; ++pass)
      for (int a = 0; a < rank; ++a) { double cdot = 0; for (int j = 0; j < nu; ++j) cdot += cand(i, j) * W(idx[a], j); for (int j = 0; j < nu; ++j) cand(i, j) -= cdot * W(idx[a], j); }
  }
  std::vector<int> used(nu, 0);
  for (int s = 0; s < m; ++s) {
    int best = -1; double bn = -1;
    for (int i = 0; i < nu; ++i) if (!used[i]) { double n2 = 0; for (int j = 0; j < nu; ++j) n2 += cand(i, j) * cand(i, j); if (n2 > bn) { bn = n2; best = i; } }
    used[best] = 1;
    const double nrm = std::sqrt(bn);
    for (int j = 0; j < nu; ++j) Pu(j, s) = cand(best, j) / nrm;
    for (int i = 0; i < nu; ++i) if (!used[i])
      for (int pass = 0; pass < 2; ++pass) { double cdot = 0; for (int j = 0; j < nu; ++j) cdot += cand(i, j) * Pu(j, s); for (int j = 0; j < nu; ++j) cand(i, j) -= cdot * Pu(j, s); }
  }
}

// ---------------------------------------------------------------- [UPSTREAM] luConstraintProjection, emulated
// What upstream does (ocs2_core LinearAlgebra::luConstraintProjection; the reference selects it with projectStateInputEqualityConstraints,
// task.info:76):  lu = Eigen::FullPivLU(D);  Pu = lu.kernel();  Px = -lu.solve(C);  Pe = -lu.solve(e).
// Restated from Eigen's documented algorithm: Gaussian elimination with complete pivoting; rank = number of pivots with
// |pivot| > eps * min(rows, cols) * |largest pivot|; solve() forward-substitutes with the unit-lower factor, back-substitutes the leading
// rank x rank block of U and sets the free unknowns to zero, i.e. it satisfies the first `rank` pivot rows and silently ignores the rest.
// This is the oracle's default since round 2 (the CUDA product implements the same elimination in-warp, k_project<NJ, true>).
constexpr double PIVOT_TIE = 1e-10;
inline void project_constraints_fullpivlu(const Mat& C, const Mat& D, const std::vector<double>& e, Mat& Px, Mat& Pu, std::vector<double>& Pe, int& rank) {
  const int nr = D.r, nu = D.c, nx = C.c, sd = std::min(nr, nu);
  Mat lu = D;
  std::vector<int> rowp(nr), colp(nu);
  for (int i = 0; i < nr; ++i) rowp[i] = i;
  for (int j = 0; j < nu; ++j) colp[j] = j;
  double maxpivot = 0.0; int nonzero = sd;
  for (int k = 0; k < sd; ++k) {
    int br = k, bc = k; double big = -1.0;
    for (int j = k; j < nu; ++j) for (int i = k; i < nr; ++i) big = std::max(big, std::fabs(lu(i, j)));
    if (big == 0.0) { nonzero = k; break; }
    // Pivot = the largest remaining coefficient (Eigen: maxCoeff visitor).  The two sole points of a stance foot make the LAST pivot of that
    // foot's block a mathematically exact tie between two rows (their residuals along the dependent direction have equal magnitude), which
    // floating-point noise would decide -- in Eigen as well.  To make the choice reproducible across implementations, coefficients within
    // PIVOT_TIE of the maximum count as tied and the first one in (original column, original row) order wins (the CUDA kernel applies the same rule).
    {
      const double tie = big * (1.0 - PIVOT_TIE);
      long bestkey = -1;
      for (int j = k; j < nu; ++j) for (int i = k; i < nr; ++i) if (std::fabs(lu(i, j)) >= tie) {
        const long key = (long)colp[j] * 4096 + rowp[i];
        if (bestkey < 0 || key < bestkey) { bestkey = key; br = i; bc = j; }
      }
    }
    maxpivot = std::max(maxpivot, big);
    if (br != k) { for (int j = 0; j < nu; ++j) std::swap(lu(k, j), lu(br, j)); std::swap(rowp[k], rowp[br]); }
    if (bc != k) { for (int i = 0; i < nr; ++i) std::swap(lu(i, k), lu(i, bc)); std::swap(colp[k], colp[bc]); }
    for (int i = k + 1; i < nr; ++i) {
      const double f = lu(i, k) / lu(k, k);
      lu(i, k) = f;
      for (int j = k + 1; j < nu; ++j) lu(i, j) -= f * lu(k, j);
    }
  }
  const double thr = std::numeric_limits<double>::epsilon() * sd * maxpivot;
  rank = 0;
  for (int k = 0; k < nonzero; ++k) if (std::fabs(lu(k, k)) > thr) ++rank;
  // Eigen counts the pivots above the threshold; with complete pivoting they are the leading ones
  auto solve = [&](const std::vector<double>& rhs, std::vector<double>& x) {
    std::vector<double> c(nr);
    for (int i = 0; i < nr; ++i) c[i] = rhs[rowp[i]];
    for (int i = 0; i < sd; ++i) for (int l = 0; l < i; ++l) c[i] -= lu(i, l) * c[l];       // unit lower
    for (int i = sd; i < nr; ++i) for (int l = 0; l < sd; ++l) c[i] -= lu(i, l) * c[l];
    for (int i = rank - 1; i >= 0; --i) { for (int j = i + 1; j < rank; ++j) c[i] -= lu(i, j) * c[j]; c[i] /= lu(i, i); }
    x.assign(nu, 0.0);
    for (int i = 0; i < rank; ++i) x[colp[i]] = c[i];
  };
  Px = Mat(nu, nx); Pe.assign(nu, 0.0);
  std::vector<double> g(nr), y;
  for (int c = 0; c < nx; ++c) { for (int k = 0; k < nr; ++k) g[k] = C(k, c); solve(g, y); for (int j = 0; j < nu; ++j) Px(j, c) = -y[j]; }
  solve(e, y); for (int j = 0; j < nu; ++j) Pe[j] = -y[j];
  // kernel: one basis vector per free (non-pivot) column f: x_f = 1, U11 x_p = -U12 e_f
  const int m = nu - rank;
  Pu = Mat(nu, m);
  for (int s_ = 0; s_ < m; ++s_) {
    const int f = rank + s_;
    std::vector<double> xp(rank);
    for (int i = rank - 1; i >= 0; --i) { double v = -lu(i, f); for (int j = i + 1; j < rank; ++j) v -= lu(i, j) * xp[j]; xp[i] = v / lu(i, i); }
    for (int i = 0; i < rank; ++i) Pu(colp[i], s_) = xp[i];
    Pu(colp[f], s_) = 1.0;
  }
}
// 1 (default): FullPivLU emulation = upstream's luConstraintProjection (what the CUDA product implements by default), 0: Moore-Penrose (the product's "projection_mode" 0)
inline int& projection_mode() { static int mode = 1; return mode; }
inline void project_constraints_dispatch(const Mat& C, const Mat& D, const std::vector<double>& e, Mat& Px, Mat& Pu, std::vector<double>& Pe, int& rank) {
  if (projection_mode() == 1) project_constraints_fullpivlu(C, D, e, Px, Pu, Pe, rank);
  else project_constraints(C, D, e, Px, Pu, Pe, rank);
}

// ---------------------------------------------------------------- node transcription [UPSTREAM multiple_shooting::setupIntermediateNode + projectTranscription]
inline void setup_intermediate_node(const Problem& P, double t, double dt, int mode, const double* x, const double* u, const double* xnext, NodeLQ& n) {
  const Model& M = *P.M; const int nx = M.nx, nu = M.nu, nj = M.nj;
  n.type = 0; n.mode = mode; n.t = t; n.dt = dt;
  // --- dynamics: RK2 with sensitivities [UPSTREAM SensitivityIntegrator RK2]
  Lin k1, k2;
  linearize(M, x, u, k1, true);
  std::vector<double> xt(nx); for (int i = 0; i < nx; ++i) xt[i] = x[i] + dt * k1.f[i];
  linearize(M, xt.data(), u, k2, false);
  Mat k2B = k2.B, k2A = k2.A;
  { Mat t1 = matmul(k2.A, k1.B); for (size_t i = 0; i < k2B.a.size(); ++i) k2B.a[i] += dt * t1.a[i]; }
  { Mat t2 = matmul(k2.A, k1.A); for (size_t i = 0; i < k2A.a.size(); ++i) k2A.a[i] += dt * t2.a[i]; }
  n.A = Mat(nx, nx); n.B = Mat(nx, nu); n.b.assign(nx, 0.0);
  for (int i = 0; i < nx; ++i) {
    for (int j = 0; j < nx; ++j) n.A(i, j) = 0.5 * dt * (k1.A(i, j) + k2A(i, j)) + (i == j ? 1.0 : 0.0);
    for (int j = 0; j < nu; ++j) n.B(i, j) = 0.5 * dt * (k1.B(i, j) + k2B(i, j));
    n.b[i] = x[i] + 0.5 * dt * (k1.f[i] + k2.f[i]) - xnext[i];
  }
  // --- cost (x dt)
  std::vector<double> xr, un; P.target.desiredState(t, xr); P.weightCompensatingInput(mode, un);
  n.Q = Mat(nx, nx); n.R = Mat(nu, nu); n.P = Mat(nu, nx); n.q.assign(nx, 0.0); n.r.assign(nu, 0.0);
  for (int i = 0; i < nx; ++i) { n.Q(i, i) = M.Q_diag[i]; n.q[i] = M.Q_diag[i] * (x[i] - xr[i]); }
  for (int i = 0; i < 12; ++i) { n.R(i, i) = M.R_force_diag[i]; n.r[i] = M.R_force_diag[i] * (u[i] - un[i]); }
  for (int i = 0; i < nj; ++i) for (int j = 0; j < nj; ++j) { n.R(12 + i, 12 + j) = M.R_joint[i * nj + j]; n.r[12 + i] += M.R_joint[i * nj + j] * (u[12 + j] - un[12 + j]); }
  double cval = P.costValue(t, mode, x, u);
  bool fl[NC]; mode_to_contact_flags(mode, fl);
  for (int k = 0; k < NC; ++k) if (fl[k]) {  // soft friction cone: [UPSTREAM StateInputSoftConstraint / MultidimensionalPenalty]
    double h, g[3], H[3][3], p, dp, ddp; P.frictionCone(u + 3 * k, h, g, H); P.barrier(h, p, dp, ddp);
    for (int a = 0; a < 3; ++a) { n.r[3 * k + a] += dp * g[a]; for (int b2 = 0; b2 < 3; ++b2) n.R(3 * k + a, 3 * k + b2) += ddp * g[a] * g[b2] + dp * H[a][b2]; }
    // FrictionConeConstraint.cpp:192-206: the Hessian shift is applied to the whole uu and xx diagonals
    for (int i = 0; i < nu; ++i) n.R(i, i) += dp * (-M.fr_shift);
    for (int i = 0; i < nx; ++i) n.Q(i, i) += dp * (-M.fr_shift);
  }
  for (auto& v : n.Q.a) v *= dt; for (auto& v : n.R.a) v *= dt; for (auto& v : n.q) v *= dt; for (auto& v : n.r) v *= dt;
  n.c = cval * dt;
  // --- equality constraints, stacking order BipedalRobotInterface.cpp:187-191
  const int nr = Problem::numEqRows(mode);
  n.nc_rows = nr; n.C = Mat(nr, nx); n.D = Mat(nr, nu);
  P.eqValues(t, mode, u, k1.pos, k1.vel, n.e);
  const double gain = M.pos_err_gain;
  int row = 0;
  for (int i = 0; i < NC; ++i) {
    if (!fl[i]) { for (int a = 0; a < 3; ++a) n.D(row + a, 3 * i + a) = 1.0; row += 3; }
    if (fl[i]) {
      for (int a = 0; a < 3; ++a) { for (int j = 0; j < nx; ++j) n.C(row + a, j) = k1.dvdx[i](a, j); for (int j = 0; j < nu; ++j) n.D(row + a, j) = k1.dvdu[i](a, j); }
      if (gain != 0.0) for (int j = 0; j < nx; ++j) n.C(row + 2, j) += gain * k1.dpdx[i](2, j);
      row += 3;
    }
    if (!fl[i]) {
      for (int j = 0; j < nx; ++j) n.C(row, j) = k1.dvdx[i](2, j) + (gain != 0.0 ? gain * k1.dpdx[i](2, j) : 0.0);
      for (int j = 0; j < nu; ++j) n.D(row, j) = k1.dvdu[i](2, j);
      row += 1;
    }
  }
  // --- performance at the linearisation point [UPSTREAM computeMetrics / toPerformanceIndex]
  n.perf_cost = n.c;
  n.perf_dyn = 0; for (int i = 0; i < nx; ++i) n.perf_dyn += n.b[i] * n.b[i]; n.perf_dyn *= dt;
  n.perf_eq = 0; for (double v : n.e) n.perf_eq += v * v; n.perf_eq *= dt;
  // --- projection + change of input variables
  project_constraints_dispatch(n.C, n.D, n.e, n.Px, n.Pu, n.Pe, n.rank);
  n.m = nu - n.rank;
  std::vector<double> RPe = matvec(n.R, n.Pe);
  std::vector<double> rr(nu); for (int i = 0; i < nu; ++i) rr[i] = n.r[i] + RPe[i];
  n.ct = n.c + dotv(n.r, n.Pe) + 0.5 * dotv(n.Pe, RPe);
  n.qt = n.q; { auto t1 = matTvec(n.Px, rr); auto t2 = matTvec(n.P, n.Pe); for (int i = 0; i < nx; ++i) n.qt[i] += t1[i] + t2[i]; }
  n.rt = matTvec(n.Pu, rr);
  Mat RPx = matmul(n.R, n.Px);
  n.Qt = n.Q; { Mat t1 = matmulT(n.Px, RPx), t2 = matmulT(n.Px, n.P), t3 = matmulT(n.P, n.Px); for (size_t i = 0; i < n.Qt.a.size(); ++i) n.Qt.a[i] += t1.a[i] + t2.a[i] + t3.a[i]; }
  { Mat PRPx = n.P; for (size_t i = 0; i < PRPx.a.size(); ++i) PRPx.a[i] += RPx.a[i]; n.Pt = matmulT(n.Pu, PRPx); }
  n.Rt = matmulT(n.Pu, matmul(n.R, n.Pu));
  n.At = n.A; { Mat t1 = matmul(n.B, n.Px); for (size_t i = 0; i < n.At.a.size(); ++i) n.At.a[i] += t1.a[i]; }
  n.bt = n.b; { auto t1 = matvec(n.B, n.Pe); for (int i = 0; i < nx; ++i) n.bt[i] += t1[i]; }
  n.Bt = matmul(n.B, n.Pu);
}

// [UPSTREAM multiple_shooting::setupEventNode] identity jump map, no input, no cost
inline void setup_event_node(const Problem& P, double t, const double* x, const double* xnext, NodeLQ& n) {
  const int nx = P.M->nx;
  n.type = 1; n.t = t; n.dt = 0; n.m = 0; n.rank = 0; n.nc_rows = 0;
  n.At = Mat(nx, nx); for (int i = 0; i < nx; ++i) n.At(i, i) = 1.0;
  n.A = n.At; n.Bt = Mat(nx, 0); n.B = Mat(nx, 0);
  n.bt.assign(nx, 0.0); for (int i = 0; i < nx; ++i) n.bt[i] = x[i] - xnext[i];
  n.b = n.bt;
  n.Qt = Mat(nx, nx); n.qt.assign(nx, 0.0); n.Rt = Mat(0, 0); n.Pt = Mat(0, nx); n.rt.clear(); n.ct = 0;
  n.perf_cost = 0; n.perf_eq = 0; n.perf_dyn = 0; for (int i = 0; i < nx; ++i) n.perf_dyn += n.bt[i] * n.bt[i];
}

// ---------------------------------------------------------------- Riccati [UPSTREAM HPIPM, equality-only OCP-QP; SURVEY.md Appendix B.7]
// nodes[0..N-1] stages, terminal value function zero (no terminal cost is installed: SURVEY.md a7).
inline bool riccati_solve(std::vector<NodeLQ>& nodes, int nx, const std::vector<double>& dx0, std::vector<std::vector<double>>& dx, std::vector<std::vector<double>>& dut) {
  const int N = (int)nodes.size();
  Mat S(nx, nx); std::vector<double> s(nx, 0.0);
  for (int k = N - 1; k >= 0; --k) {
    NodeLQ& n = nodes[k]; const int m = n.m;
    Mat SA = matmul(S, n.At);
    std::vector<double> Sb = matvec(S, n.bt); for (int i = 0; i < nx; ++i) Sb[i] += s[i];   // s + S b
    Mat Snew = n.Qt; { Mat t = matmulT(n.At, SA); for (size_t i = 0; i < Snew.a.size(); ++i) Snew.a[i] += t.a[i]; }
    std::vector<double> snew = n.qt; { auto t = matTvec(n.At, Sb); for (int i = 0; i < nx; ++i) snew[i] += t[i]; }
    n.Kt = Mat(m, nx); n.kt.assign(m, 0.0);
    if (m > 0) {
      Mat SB = matmul(S, n.Bt);
      Mat G = n.Rt; { Mat t = matmulT(n.Bt, SB); for (size_t i = 0; i < G.a.size(); ++i) G.a[i] += t.a[i]; }
      Mat H = n.Pt; { Mat t = matmulT(n.Bt, SA); for (size_t i = 0; i < H.a.size(); ++i) H.a[i] += t.a[i]; }
      std::vector<double> g = n.rt; { auto t = matTvec(n.Bt, Sb); for (int i = 0; i < m; ++i) g[i] += t[i]; }
      if (!cholesky(G)) return false;
      std::vector<double> col(m);
      for (int j = 0; j < nx; ++j) { for (int i = 0; i < m; ++i) col[i] = H(i, j); chol_solve(G, col.data()); for (int i = 0; i < m; ++i) n.Kt(i, j) = -col[i]; }
      for (int i = 0; i < m; ++i) col[i] = g[i]; chol_solve(G, col.data()); for (int i = 0; i < m; ++i) n.kt[i] = -col[i];
      { Mat t = matmulT(H, n.Kt); for (size_t i = 0; i < Snew.a.size(); ++i) Snew.a[i] += t.a[i]; }
      { auto t = matTvec(H, n.kt); for (int i = 0; i < nx; ++i) snew[i] += t[i]; }
    }
    for (int i = 0; i < nx; ++i) for (int j = i + 1; j < nx; ++j) { const double v = 0.5 * (Snew(i, j) + Snew(j, i)); Snew(i, j) = v; Snew(j, i) = v; }
    S = Snew; s = snew;
  }
  dx.assign(N + 1, std::vector<double>(nx, 0.0)); dut.assign(N, std::vector<double>());
  dx[0] = dx0;
  for (int k = 0; k < N; ++k) {
    const NodeLQ& n = nodes[k];
    dut[k] = matvec(n.Kt, dx[k]); for (int i = 0; i < n.m; ++i) dut[k][i] += n.kt[i];
    dx[k + 1] = matvec(n.At, dx[k]);
    if (n.m > 0) { auto t = matvec(n.Bt, dut[k]); for (int i = 0; i < nx; ++i) dx[k + 1][i] += t[i]; }
    for (int i = 0; i < nx; ++i) dx[k + 1][i] += n.bt[i];
  }
  return true;
}

// ---------------------------------------------------------------- primal solution + solver
struct PrimalSolution {
  std::vector<double> times; std::vector<int> events;   // node times (interpolation times) and event annotation
  std::vector<std::vector<double>> x, u;                // N+1 each (u filled at event/terminal nodes by copying)
  std::vector<std::vector<double>> uff; std::vector<Mat> K;   // N+1 each
  ModeSchedule modeSchedule;
  bool empty() const { return times.empty(); }
};

struct SolveInfo {
  Performance before, after; double step = 0; int trials = 0; int n_nodes = 0; int status = 0;
  double dx_norm = 0, du_norm = 0, armijo = 0;
};

struct Solver {
  Model M;
  Problem P;
  GaitSchedule gait;
  bool explicitSchedule = false;     // true: modeSchedule set by the caller each tick, GaitSchedule bypassed
  double dt, horizon;
  PrimalSolution sol;
  SolveInfo info;
  std::vector<NodeLQ> nodes;         // kept for inspection by tests
  std::vector<AnnotatedTime> td;
  std::vector<std::vector<double>> x_lin, u_lin, dx, du;   // linearisation trajectories and QP step (un-projected du)

  explicit Solver(const Model& m) : M(m) {
    P.M = &M;
    dt = M.sqp_dt; horizon = M.time_horizon;
    P.swing.liftOffVelocity = M.liftoff_vel; P.swing.touchDownVelocity = M.touchdown_vel;
    P.swing.swingHeight = M.swing_height; P.swing.swingTimeScale = M.swing_time_scale;
    gait.ms.eventTimes = M.init_events; gait.ms.modeSequence = M.init_modes;
    gait.tmpl = M.default_template; gait.phaseTransitionStanceTime = M.phase_transition_stance_time;
  }
  void reset() { sol = PrimalSolution(); }

  // [UPSTREAM SqpSolver::computePerformance]
  Performance computePerformance(const std::vector<double>& x0, const std::vector<std::vector<double>>& x, const std::vector<std::vector<double>>& u) const {
    const int N = (int)td.size() - 1; const int nx = M.nx;
    Performance pf;
    for (int i = 0; i < N; ++i) {
      if (td[i].event == 1) {
        double s = 0; for (int k = 0; k < nx; ++k) { const double d = x[i][k] - x[i + 1][k]; s += d * d; }
        pf.dynSSE += s;
      } else {
        const double ti = interval_start(td[i]); const double dti = interval_end(td[i + 1]) - ti;
        const int mode = P.modeSchedule.modeAtTime(ti);
        double xn[MAXX]; V3<double> pos[NC], vel[NC];
        rk2_value(M, x[i].data(), u[i].data(), dti, xn, pos, vel);
        double s = 0; for (int k = 0; k < nx; ++k) { const double d = xn[k] - x[i + 1][k]; s += d * d; }
        pf.dynSSE += dti * s;
        pf.cost += dti * P.costValue(ti, mode, x[i].data(), u[i].data());
        std::vector<double> e; P.eqValues(ti, mode, u[i].data(), pos, vel, e);
        double se = 0; for (double v : e) se += v * v;
        pf.eqSSE += dti * se;
      }
    }
    double s0 = 0; for (int k = 0; k < nx; ++k) { const double d = x0[k] - x[0][k]; s0 += d * d; }
    pf.dynSSE += s0;
    return pf;
  }

  // [UPSTREAM MPC_BASE::run -> SolverBase::run -> SqpSolver::runImpl], one tick
  void run(double t0, const std::vector<double>& x0) {
    const int nx = M.nx, nu = M.nu;
    const double tf = t0 + horizon;
    info = SolveInfo();
    // preSolverRun: SwitchedModelReferenceManager::modifyReferences (SwitchedModelReferenceManager.cpp:62-69)
    if (!explicitSchedule) P.modeSchedule = gait.getModeSchedule(t0 - horizon, tf + horizon);
    P.swing.update(P.modeSchedule, 0.0);
    td = time_discretization_with_events(t0, tf, dt, P.modeSchedule.eventTimes);
    const int N = (int)td.size() - 1;
    info.n_nodes = N + 1;
    // [UPSTREAM multiple_shooting::initializeStateInputTrajectories]
    std::vector<std::vector<double>> x(N + 1), u(N);
    double stateTill = td.front().time, inputTill = td.front().time;
    if (sol.times.size() >= 2) { stateTill = sol.times.back(); inputTill = sol.times[sol.times.size() - 2]; }
    const double tinit = interval_start(td[0]);
    if (tinit < stateTill) interpolate(tinit, sol.times, sol.x, x[0]); else x[0] = x0;
    for (int i = 0; i < N; ++i) {
      if (td[i].event == 1) { u[i].assign(nu, 0.0); x[i + 1] = x[i]; continue; }
      const double ti = interval_start(td[i]), tn = interval_end(td[i + 1]);
      if (ti > inputTill || tn > stateTill) {  // initializer/BipedalRobotInitializer.cpp:56-63 (extendNormalizedMomentum = true)
        P.weightCompensatingInput(P.modeSchedule.modeAtTime(ti), u[i]); x[i + 1] = x[i];
      } else { interpolate(ti, sol.times, sol.u, u[i]); interpolate(tn, sol.times, sol.x, x[i + 1]); }
    }
    for (int iter = 0; iter < M.sqp_iterations; ++iter) {
      // setupQuadraticSubproblem
      nodes.assign(N, NodeLQ());
      Performance base;
      for (int i = 0; i < N; ++i) {
        if (td[i].event == 1) setup_event_node(P, td[i].time, x[i].data(), x[i + 1].data(), nodes[i]);
        else {
          const double ti = interval_start(td[i]); const double dti = interval_end(td[i + 1]) - ti;
          setup_intermediate_node(P, ti, dti, P.modeSchedule.modeAtTime(ti), x[i].data(), u[i].data(), x[i + 1].data(), nodes[i]);
        }
        base.cost += nodes[i].perf_cost; base.dynSSE += nodes[i].perf_dyn; base.eqSSE += nodes[i].perf_eq;
      }
      std::vector<double> dx0(nx); double s0 = 0;
      for (int k = 0; k < nx; ++k) { dx0[k] = x0[k] - x[0][k]; s0 += dx0[k] * dx0[k]; }
      base.dynSSE += s0;
      if (iter == 0) info.before = base;
      // QP
      std::vector<std::vector<double>> dut;
      if (!riccati_solve(nodes, nx, dx0, dx, dut)) { info.status = 1; return; }
      // armijoDescentMetric on the projected cost, then remap the input [UPSTREAM SqpSolver::getOCPSolution]
      double armijo = 0;
      du.assign(N, std::vector<double>(nu, 0.0));
      for (int i = 0; i < N; ++i) {
        NodeLQ& n = nodes[i];
        armijo += dotv(n.qt, dx[i]);
        if (n.type == 0) {
          armijo += dotv(n.rt, dut[i]);
          auto a = matvec(n.Px, dx[i]); auto b2 = matvec(n.Pu, dut[i]);
          for (int k = 0; k < nu; ++k) du[i][k] = n.Pe[k] + a[k] + b2[k];
          n.K = matmul(n.Pu, n.Kt); for (size_t k = 0; k < n.K.a.size(); ++k) n.K.a[k] += n.Px.a[k];
        } else n.K = Mat(nu, nx);
      }
      info.armijo = armijo;
      x_lin = x; u_lin = u;
      // takeStep: filter line search [UPSTREAM SqpSolver::takeStep, FilterLinesearch::acceptStep]
      double dxn = 0, dun = 0;
      for (int i = 0; i <= N; ++i) for (double v : dx[i]) dxn += v * v;
      for (int i = 0; i < N; ++i) if (nodes[i].type == 0) for (double v : du[i]) dun += v * v;
      dxn = std::sqrt(dxn); dun = std::sqrt(dun);
      info.dx_norm = dxn; info.du_norm = dun;
      const double alpha_decay = 0.5, alpha_min = 1e-4, gamma_c = 1e-6, armijoFactor = 1e-4;
      double alpha = 1.0; bool accepted = false; Performance pnew;
      std::vector<std::vector<double>> xn(N + 1), unew(N);
      do {
        for (int i = 0; i <= N; ++i) { xn[i] = x[i]; for (int k = 0; k < nx; ++k) xn[i][k] += alpha * dx[i][k]; }
        for (int i = 0; i < N; ++i) { unew[i] = u[i]; if (nodes[i].type == 0) for (int k = 0; k < nu; ++k) unew[i][k] += alpha * du[i][k]; }
        pnew = computePerformance(x0, xn, unew);
        ++info.trials;
        const double th0 = base.theta(), th = pnew.theta();
        if (th > M.g_max) accepted = th < (1.0 - gamma_c) * th0;
        else if (th < M.g_min && th0 < M.g_min && armijo < 0.0) accepted = pnew.merit() < base.merit() + armijoFactor * alpha * armijo;
        else accepted = pnew.merit() < base.merit() - gamma_c * th0 || th < (1.0 - gamma_c) * th0;
        if (accepted) break;
        alpha *= alpha_decay;
        if (alpha * dxn < M.delta_tol && alpha * dun < M.delta_tol) break;
      } while (alpha >= alpha_min);
      if (accepted) { x = xn; u = unew; info.step = alpha; info.after = pnew; }
      else { info.step = 0.0; info.after = base; }
    }
    // toPrimalSolution with feedback [UPSTREAM multiple_shooting::toPrimalSolution]
    PrimalSolution ps;
    ps.modeSchedule = P.modeSchedule;
    ps.times.resize(N + 1); ps.events.resize(N + 1);
    for (int i = 0; i <= N; ++i) { ps.times[i] = td[i].time; ps.events[i] = td[i].event; }
    ps.x = x; ps.u.resize(N + 1); ps.uff.resize(N + 1); ps.K.resize(N + 1);
    for (int i = 0; i < N; ++i) {
      if (td[i].event == 1 && i > 0) { ps.u[i] = ps.u[i - 1]; ps.uff[i] = ps.uff[i - 1]; ps.K[i] = ps.K[i - 1]; }
      else {
        ps.u[i] = u[i]; ps.K[i] = nodes[i].K;
        ps.uff[i] = u[i]; auto kx = matvec(ps.K[i], x[i]); for (int k = 0; k < nu; ++k) ps.uff[i][k] -= kx[k];
      }
    }
    ps.u[N] = ps.u[N - 1]; ps.uff[N] = ps.uff[N - 1]; ps.K[N] = ps.K[N - 1];
    sol = ps;
  }

  // [UPSTREAM MPC_MRT_Interface::evaluatePolicy / LinearController::computeInput]
  void evaluatePolicy(double t, const double* xm, double* xOpt, double* uOpt, int* mode) const {
    const int nx = M.nx, nu = M.nu;
    auto seg = time_segment(t, sol.times);
    const int i = seg.first; const double a = seg.second;
    const int i1 = std::min(i + 1, (int)sol.times.size() - 1);
    for (int k = 0; k < nx; ++k) xOpt[k] = a * sol.x[i][k] + (1 - a) * sol.x[i1][k];
    for (int k = 0; k < nu; ++k) {
      double v = a * sol.uff[i][k] + (1 - a) * sol.uff[i1][k];
      for (int j = 0; j < nx; ++j) v += (a * sol.K[i](k, j) + (1 - a) * sol.K[i1](k, j)) * xm[j];
      uOpt[k] = v;
    }
    *mode = sol.modeSchedule.modeAtTime(t);
  }

  // ---- feedback-policy rollout between MPC ticks
  // [UPSTREAM MRT_BASE::rolloutPolicy -> TimeTriggeredRollout::run] (installed by mpcMrtInterface_->initRollout(&interface.getRollout()),
  // bipedal_controllers/src/BipedalController.cpp:322; settings task.info:159-167: ODE45, AbsTolODE 1e-5, RelTolODE 1e-3, timeStep 0.015).
  // The closed-loop system xdot = f(x, uff(t) + K(t) x) is integrated from t to t + timeStep.  Restated from OCS2 / boost::odeint:
  //  * RolloutBase::findActiveModesTimeInterval splits [t, t + timeStep] at the event times of the policy's mode schedule and starts every
  //    sub-interval weakEpsilon late (the state is carried over unchanged; the jump map is the identity);
  //  * each sub-interval runs odeint::integrate_adaptive(make_controlled<runge_kutta_dopri5>(AbsTol, RelTol), sys, x, begin, end, dtInitial = timeStep):
  //    Dormand-Prince 5(4) with the default error checker  err = max_i |xerr_i| / (abs + rel (|x_i| + dt |dxdt_i|))  and the default step
  //    adjuster  (reject: dt *= max(0.9 err^(-1/3), 0.2); accept with err < 0.5: dt *= 0.9 max(err, 5^-5)^(-1/5)), a fresh stepper per interval.
  struct RolloutSettings { double absTol = 1e-5, relTol = 1e-3, timeStep = 0.015; int maxTrials = 500; };
  RolloutSettings rollout;
  void closedLoopFlow(double t, const double* x, double* f) const {
    double xo[MAXX], u[MAXU]; int mode;
    evaluatePolicy(t, x, xo, u, &mode);
    flow_map<double>(M, x, u, f);
  }
  int integrateAdaptiveDopri5(double* x, double t0, double t1, double dtInit) const {
    static const double a21 = 1.0 / 5, a31 = 3.0 / 40, a32 = 9.0 / 40, a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9, a51 = 19372.0 / 6561, a52 = -25360.0 / 2187,
                        a53 = 64448.0 / 6561, a54 = -212.0 / 729, a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176, a65 = -5103.0 / 18656,
                        c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
    static const double dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640, dc5 = c5 + 92097.0 / 339200, dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;
    const int nx = M.nx;
    double k1[MAXX], k2[MAXX], k3[MAXX], k4[MAXX], k5[MAXX], k6[MAXX], k7[MAXX], xt[MAXX], xn[MAXX];
    double t = t0, dt = dtInit; int steps = 0;
    closedLoopFlow(t, x, k1);   // first call of the FSAL stepper
    while (t < t1 && (t1 - t) > 1e-15 * std::max(1.0, std::fabs(t1))) {
      if (t + dt > t1) dt = t1 - t;
      int trials = 0; bool ok = false;
      while (!ok && trials < rollout.maxTrials) {
        ++trials;
        for (int i = 0; i < nx; ++i) xt[i] = x[i] + dt * a21 * k1[i];
        closedLoopFlow(t + dt * (1.0 / 5), xt, k2);
        for (int i = 0; i < nx; ++i) xt[i] = x[i] + dt * (a31 * k1[i] + a32 * k2[i]);
        closedLoopFlow(t + dt * (3.0 / 10), xt, k3);
        for (int i = 0; i < nx; ++i) xt[i] = x[i] + dt * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
        closedLoopFlow(t + dt * (4.0 / 5), xt, k4);
        for (int i = 0; i < nx; ++i) xt[i] = x[i] + dt * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
        closedLoopFlow(t + dt * (8.0 / 9), xt, k5);
        for (int i = 0; i < nx; ++i) xt[i] = x[i] + dt * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
        closedLoopFlow(t + dt, xt, k6);
        for (int i = 0; i < nx; ++i) xn[i] = x[i] + dt * (c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c5 * k5[i] + c6 * k6[i]);
        closedLoopFlow(t + dt, xn, k7);
        double err = 0.0;
        for (int i = 0; i < nx; ++i) {
          const double xe = dt * (dc1 * k1[i] + dc3 * k3[i] + dc4 * k4[i] + dc5 * k5[i] + dc6 * k6[i] + dc7 * k7[i]);
          err = std::max(err, std::fabs(xe) / (rollout.absTol + rollout.relTol * (std::fabs(x[i]) + std::fabs(dt) * std::fabs(k1[i]))));
        }
        if (err > 1.0) { dt *= std::max(0.9 * std::pow(err, -1.0 / 3.0), 0.2); continue; }
        ok = true; t += dt; ++steps;
        for (int i = 0; i < nx; ++i) { x[i] = xn[i]; k1[i] = k7[i]; }
        if (err < 0.5) { err = std::max(std::pow(5.0, -5.0), err); dt *= 0.9 * std::pow(err, -1.0 / 5.0); }
      }
      if (!ok) throw std::runtime_error("[oracle] rollout: step size underflow");
    }
    return steps;
  }
  // x <- state after the closed-loop rollout from t to t + timeStep; returns the number of accepted integration steps
  int rolloutPolicy(double t, double* x, double timeStep) const {
    const double tf = t + timeStep;
    const auto& ev = sol.modeSchedule.eventTimes;
    std::vector<double> sw; sw.push_back(t);
    for (auto it = std::upper_bound(ev.begin(), ev.end(), t); it != ev.end() && *it < tf; ++it) sw.push_back(*it);   // events inside (t, tf)
    sw.push_back(tf);
    int steps = 0;
    for (size_t i = 0; i + 1 < sw.size(); ++i) {
      const double begin = std::min(sw[i] + WEAK_EPS, sw[i + 1]), end = sw[i + 1];
      if (end > begin) steps += integrateAdaptiveDopri5(x, begin, end, rollout.timeStep);
    }
    return steps;
  }
};

}  // namespace orc
