// TEST INFRASTRUCTURE - C ABI of the CPU oracle for ctypes (tests/, smoke, bench cpu_baseline only).
// PARITY UNPINNED (see oracle_model.hpp).  Build: make -C oracle
#include <atomic>
#include <chrono>
#include <thread>
#include "oracle_solver.hpp"

using namespace orc;

namespace {
thread_local std::string g_err;
struct Batch {
  std::vector<std::unique_ptr<Solver>> s;
  std::vector<double> t0; std::vector<std::vector<double>> x0;
  std::vector<std::array<double, 4>> cmd; std::vector<char> has_cmd; double ttt = 1.0;
};
}  // namespace

#define ORC_TRY try {
#define ORC_CATCH(ret) } catch (const std::exception& e) { g_err = e.what(); return ret; }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_create(const char* model_path) { ORC_TRY return new Solver(load_model(model_path)); ORC_CATCH(nullptr) }
void orc_destroy(void* h) { delete static_cast<Solver*>(h); }
void orc_dims(void* h, int* nx, int* nu, int* nj) { auto* s = static_cast<Solver*>(h); *nx = s->M.nx; *nu = s->M.nu; *nj = s->M.nj; }
double orc_total_mass(void* h) { return static_cast<Solver*>(h)->M.total_mass; }
void orc_get_initial_state(void* h, double* x) { auto* s = static_cast<Solver*>(h); for (int i = 0; i < s->M.nx; ++i) x[i] = s->M.initial_state[i]; }
void orc_set_dt_horizon(void* h, double dt, double horizon) { auto* s = static_cast<Solver*>(h); s->dt = dt; s->horizon = horizon; }
void orc_set_sqp_iterations(void* h, int it) { static_cast<Solver*>(h)->M.sqp_iterations = it; }
void orc_reset(void* h) { static_cast<Solver*>(h)->reset(); }

void orc_set_mode_schedule(void* h, int n_events, const double* et, const int* modes) {
  auto* s = static_cast<Solver*>(h);
  s->explicitSchedule = true;
  s->P.modeSchedule.eventTimes.assign(et, et + n_events);
  s->P.modeSchedule.modeSequence.assign(modes, modes + n_events + 1);
}
void orc_use_gait_schedule(void* h) { static_cast<Solver*>(h)->explicitSchedule = false; }
void orc_set_target(void* h, int npts, const double* times, const double* states) {
  auto* s = static_cast<Solver*>(h);
  s->P.target.times.assign(times, times + npts);
  s->P.target.states.resize(npts);
  for (int i = 0; i < npts; ++i) s->P.target.states[i].assign(states + (size_t)i * s->M.nx, states + (size_t)(i + 1) * s->M.nx);
}
void orc_set_target_cmd_vel(void* h, double t_obs, const double* x_obs, const double* cmd, double time_to_target) {
  auto* s = static_cast<Solver*>(h);
  s->P.target = cmd_vel_to_target(s->M, t_obs, x_obs, cmd, time_to_target);
}
int orc_get_target(void* h, double* times, double* states) {
  auto* s = static_cast<Solver*>(h);
  const int n = (int)s->P.target.times.size();
  for (int i = 0; i < n; ++i) { times[i] = s->P.target.times[i]; for (int k = 0; k < s->M.nx; ++k) states[i * s->M.nx + k] = s->P.target.states[i][k]; }
  return n;
}

// GaitSchedule (gait/GaitSchedule.cpp)
int orc_gait_insert(void* h, int n_modes, const int* modes, const double* times, double start, double fin) {
  ORC_TRY auto* s = static_cast<Solver*>(h);
  GaitTemplate t; t.modes.assign(modes, modes + n_modes); t.times.assign(times, times + n_modes + 1);
  s->gait.insertModeSequenceTemplate(t, start, fin); return 0; ORC_CATCH(-1)
}
int orc_gait_get(void* h, double lower, double upper, int cap, double* et, int* modes) {
  ORC_TRY auto* s = static_cast<Solver*>(h);
  ModeSchedule ms = s->gait.getModeSchedule(lower, upper);
  const int n = (int)ms.eventTimes.size();
  if (n > cap) { g_err = "capacity"; return -1; }
  for (int i = 0; i < n; ++i) et[i] = ms.eventTimes[i];
  for (int i = 0; i <= n; ++i) modes[i] = ms.modeSequence[i];
  return n; ORC_CATCH(-1)
}
int orc_gait_peek(void* h, int cap, double* et, int* modes) {
  auto* s = static_cast<Solver*>(h);
  const int n = (int)s->gait.ms.eventTimes.size();
  if (n > cap) return -1;
  for (int i = 0; i < n; ++i) et[i] = s->gait.ms.eventTimes[i];
  for (int i = 0; i <= n; ++i) modes[i] = s->gait.ms.modeSequence[i];
  return n;
}

int orc_run(void* h, double t0, const double* x0) {
  ORC_TRY auto* s = static_cast<Solver*>(h);
  s->run(t0, std::vector<double>(x0, x0 + s->M.nx)); return s->info.status; ORC_CATCH(-1)
}
int orc_num_nodes(void* h) { return (int)static_cast<Solver*>(h)->sol.times.size(); }
void orc_get_times(void* h, double* t, int* ev) { auto* s = static_cast<Solver*>(h); for (size_t i = 0; i < s->sol.times.size(); ++i) { t[i] = s->sol.times[i]; ev[i] = s->sol.events[i]; } }
void orc_get_solution(void* h, double* x, double* u, double* uff, double* K) {
  auto* s = static_cast<Solver*>(h); const int nx = s->M.nx, nu = s->M.nu; const size_t n = s->sol.times.size();
  for (size_t i = 0; i < n; ++i) {
    if (x) for (int k = 0; k < nx; ++k) x[i * nx + k] = s->sol.x[i][k];
    if (u) for (int k = 0; k < nu; ++k) u[i * nu + k] = s->sol.u[i][k];
    if (uff) for (int k = 0; k < nu; ++k) uff[i * nu + k] = s->sol.uff[i][k];
    if (K) for (int k = 0; k < nu * nx; ++k) K[i * nu * nx + k] = s->sol.K[i].a[k];
  }
}
// info: [0..2] before {cost, dynSSE, eqSSE}, [3..5] after, [6] step, [7] trials, [8] armijo, [9] dx_norm, [10] du_norm, [11] n_nodes
void orc_get_info(void* h, double* o) {
  auto* s = static_cast<Solver*>(h); const auto& I = s->info;
  o[0] = I.before.cost; o[1] = I.before.dynSSE; o[2] = I.before.eqSSE; o[3] = I.after.cost; o[4] = I.after.dynSSE; o[5] = I.after.eqSSE;
  o[6] = I.step; o[7] = I.trials; o[8] = I.armijo; o[9] = I.dx_norm; o[10] = I.du_norm; o[11] = I.n_nodes;
}
void orc_get_step(void* h, double* dx, double* du, double* xl, double* ul) {
  auto* s = static_cast<Solver*>(h); const int nx = s->M.nx, nu = s->M.nu; const size_t N = s->du.size();
  for (size_t i = 0; i <= N; ++i) for (int k = 0; k < nx; ++k) { if (dx) dx[i * nx + k] = s->dx[i][k]; if (xl) xl[i * nx + k] = s->x_lin[i][k]; }
  for (size_t i = 0; i < N; ++i) for (int k = 0; k < nu; ++k) { if (du) du[i * nu + k] = s->du[i][k]; if (ul) ul[i * nu + k] = s->u_lin[i][k]; }
}
// dense LQ data of node k: sizes nx*nx, nx*nu, nx, nx*nx, nu*nu, nx, nu, 16*nx, 16*nu, 16 ; meta = {type, mode, nc_rows, m, rank}; tdt = {t, dt}
int orc_get_node_lq(void* h, int k, double* A, double* B, double* b, double* Q, double* R, double* q, double* r, double* C, double* D, double* e, int* meta, double* tdt) {
  auto* s = static_cast<Solver*>(h); const int nx = s->M.nx, nu = s->M.nu;
  if (k < 0 || k >= (int)s->nodes.size()) return -1;
  const NodeLQ& n = s->nodes[k];
  meta[0] = n.type; meta[1] = n.mode; meta[2] = n.nc_rows; meta[3] = n.m; meta[4] = n.rank; tdt[0] = n.t; tdt[1] = n.dt;
  if (n.type != 0) { for (int i = 0; i < nx; ++i) b[i] = n.bt[i]; return 0; }
  for (int i = 0; i < nx * nx; ++i) { A[i] = n.A.a[i]; Q[i] = n.Q.a[i]; }
  for (int i = 0; i < nx * nu; ++i) B[i] = n.B.a[i];
  for (int i = 0; i < nu * nu; ++i) R[i] = n.R.a[i];
  for (int i = 0; i < nx; ++i) { b[i] = n.b[i]; q[i] = n.q[i]; }
  for (int i = 0; i < nu; ++i) r[i] = n.r[i];
  for (int i = 0; i < n.nc_rows * nx; ++i) C[i] = n.C.a[i];
  for (int i = 0; i < n.nc_rows * nu; ++i) D[i] = n.D.a[i];
  for (int i = 0; i < n.nc_rows; ++i) e[i] = n.e[i];
  return 0;
}
// projection of node k: Px (nu*nx), Pu (nu*m), Pe (nu), full gain K (nu*nx)
int orc_get_node_projection(void* h, int k, double* Px, double* Pu, double* Pe, double* K) {
  auto* s = static_cast<Solver*>(h);
  if (k < 0 || k >= (int)s->nodes.size()) return -1;
  const NodeLQ& n = s->nodes[k];
  if (n.type != 0) return 1;
  for (size_t i = 0; i < n.Px.a.size(); ++i) Px[i] = n.Px.a[i];
  for (size_t i = 0; i < n.Pu.a.size(); ++i) Pu[i] = n.Pu.a[i];
  for (size_t i = 0; i < n.Pe.size(); ++i) Pe[i] = n.Pe[i];
  if (K) for (size_t i = 0; i < n.K.a.size(); ++i) K[i] = n.K.a[i];
  return 0;
}
void orc_evaluate_policy(void* h, double t, const double* xm, double* xOpt, double* uOpt, int* mode) { static_cast<Solver*>(h)->evaluatePolicy(t, xm, xOpt, uOpt, mode); }
// MRT_BASE::rolloutPolicy: x <- closed-loop state at t + time_step under the current policy; `substeps` consecutive calls of time_step / substeps
int orc_rollout_policy(void* h, double t, double* x, double time_step, int substeps) {
  ORC_TRY auto* s = static_cast<Solver*>(h); int steps = 0;
  const double hstep = time_step / substeps;
  for (int i = 0; i < substeps; ++i) steps += s->rolloutPolicy(t + i * hstep, x, hstep);
  return steps; ORC_CATCH(-1)
}

// ---- model maths
void orc_flow_map(void* h, const double* x, const double* u, double* f, double* pos, double* vel) {
  auto* s = static_cast<Solver*>(h); V3<double> p[NC], v[NC];
  flow_map<double>(s->M, x, u, f, p, v);
  for (int c = 0; c < NC; ++c) for (int r = 0; r < 3; ++r) { if (pos) pos[3 * c + r] = p[c][r]; if (vel) vel[3 * c + r] = v[c][r]; }
}
void orc_linearize(void* h, const double* x, const double* u, double* f, double* A, double* B, double* dpdx, double* dvdx, double* dvdu) {
  auto* s = static_cast<Solver*>(h); const int nx = s->M.nx, nu = s->M.nu;
  Lin L; linearize(s->M, x, u, L, true);
  for (int i = 0; i < nx; ++i) f[i] = L.f[i];
  for (int i = 0; i < nx * nx; ++i) A[i] = L.A.a[i];
  for (int i = 0; i < nx * nu; ++i) B[i] = L.B.a[i];
  for (int c = 0; c < NC; ++c) {
    for (int i = 0; i < 3 * nx; ++i) { if (dpdx) dpdx[c * 3 * nx + i] = L.dpdx[c].a[i]; if (dvdx) dvdx[c * 3 * nx + i] = L.dvdx[c].a[i]; }
    for (int i = 0; i < 3 * nu; ++i) if (dvdu) dvdu[c * 3 * nu + i] = L.dvdu[c].a[i];
  }
}
void orc_cmm(void* h, const double* q, double* A, double* com) {
  auto* s = static_cast<Solver*>(h); Kin<double> K; forward_kinematics<double>(s->M, q, K);
  centroidal_momentum_matrix<double>(s->M, K, A);
  for (int r = 0; r < 3; ++r) com[r] = K.com[r];
}
// body world COMs (nj+1)*3, masses (nj+1), world inertias (nj+1)*9 for brute-force checks
void orc_bodies(void* h, const double* q, double* c, double* m, double* I) {
  auto* s = static_cast<Solver*>(h); Kin<double> K; forward_kinematics<double>(s->M, q, K);
  for (int b = 0; b <= s->M.nj; ++b) { m[b] = K.mbody[b]; for (int r = 0; r < 3; ++r) { c[3 * b + r] = K.cbody[b][r]; for (int cc = 0; cc < 3; ++cc) I[9 * b + 3 * r + cc] = K.Ibody[b].m[r][cc]; } }
}
void orc_friction(void* h, const double* F, double* hval, double* g, double* H, double* pen) {
  auto* s = static_cast<Solver*>(h); double HH[3][3];
  s->P.frictionCone(F, *hval, g, HH); for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) H[3 * i + j] = HH[i][j];
  s->P.barrier(*hval, pen[0], pen[1], pen[2]);
}
void orc_barrier(void* h, double hv, double* pen) { static_cast<Solver*>(h)->P.barrier(hv, pen[0], pen[1], pen[2]); }
int orc_time_discretization(double t0, double tf, double dt, int ne, const double* ev, int cap, double* t, int* e) {
  auto td = time_discretization_with_events(t0, tf, dt, std::vector<double>(ev, ev + ne));
  if ((int)td.size() > cap) return -1;
  for (size_t i = 0; i < td.size(); ++i) { t[i] = td[i].time; e[i] = td[i].event; }
  return (int)td.size();
}
// swing reference for an explicit mode schedule: returns zvel/zpos of the 4 contacts at the n query times
int orc_swing(void* h, int ne, const double* et, const int* modes, int n, const double* tq, double* zvel, double* zpos) {
  ORC_TRY auto* s = static_cast<Solver*>(h);
  ModeSchedule ms; ms.eventTimes.assign(et, et + ne); ms.modeSequence.assign(modes, modes + ne + 1);
  SwingPlanner sp = s->P.swing; sp.update(ms, 0.0);
  for (int i = 0; i < n; ++i) for (int c = 0; c < NC; ++c) { zvel[i * NC + c] = sp.zvel(c, tq[i]); zpos[i * NC + c] = sp.zpos(c, tq[i]); }
  return 0; ORC_CATCH(-1)
}
void orc_spline(double ts, double ps, double vs, double mid, double tf, double pf, double vf, int n, const double* tq, double* pos, double* vel) {
  SplineCpg sp(ts, ps, vs, mid, tf, pf, vf);
  for (int i = 0; i < n; ++i) { pos[i] = sp.position(tq[i]); vel[i] = sp.velocity(tq[i]); }
}
// 0: Moore-Penrose projection (default = the CUDA product), 1: emulation of upstream's Eigen::FullPivLU projection (process-wide switch, tests only)
void orc_set_projection_mode(int mode) { projection_mode() = mode; }
// exact operation counts of ONE evaluation of the flow map + contact kinematics at (x, u): out = {add, mul, div, trig} (counting scalar)
void orc_count_flow_map(void* h, const double* x, const double* u, unsigned long long* out) {
  auto* s = static_cast<Solver*>(h); const Model& M = s->M;
  Counted xc[MAXX], uc[MAXU], fc[MAXX]; V3<Counted> pc[NC], vc[NC];
  for (int i = 0; i < M.nx; ++i) xc[i] = Counted(x[i]);
  for (int i = 0; i < M.nu; ++i) uc[i] = Counted(u[i]);
  flop_counter() = FlopCount();
  flow_map<Counted>(M, xc, uc, fc, pc, vc);
  const FlopCount c = flop_counter();
  out[0] = c.add; out[1] = c.mul; out[2] = c.div; out[3] = c.trig;
}
int orc_project(int nr, int nu, int nx, const double* C, const double* D, const double* e, double* Px, double* Pu, double* Pe) {
  ORC_TRY Mat Cm(nr, nx), Dm(nr, nu); Cm.a.assign(C, C + nr * nx); Dm.a.assign(D, D + nr * nu);
  std::vector<double> ev(e, e + nr), pe; Mat px, pu; int rank = 0;
  project_constraints_dispatch(Cm, Dm, ev, px, pu, pe, rank);
  for (size_t i = 0; i < px.a.size(); ++i) Px[i] = px.a[i];
  for (size_t i = 0; i < pu.a.size(); ++i) Pu[i] = pu.a[i];
  for (size_t i = 0; i < pe.size(); ++i) Pe[i] = pe[i];
  return rank; ORC_CATCH(-1)
}
// generic Riccati KAT: N stages, each with m[k] inputs; arrays are concatenated stage by stage
int orc_riccati(int N, int nx, const int* m, const double* A, const double* B, const double* b, const double* Q, const double* R, const double* P,
                const double* q, const double* r, const double* dx0, double* dx, double* du, double* Kt) {
  ORC_TRY std::vector<NodeLQ> nodes(N);
  size_t oB = 0, oR = 0, oP = 0, or_ = 0;
  for (int k = 0; k < N; ++k) {
    NodeLQ& n = nodes[k]; n.m = m[k]; n.type = m[k] > 0 ? 0 : 1;
    n.At = Mat(nx, nx); n.At.a.assign(A + (size_t)k * nx * nx, A + (size_t)(k + 1) * nx * nx);
    n.Qt = Mat(nx, nx); n.Qt.a.assign(Q + (size_t)k * nx * nx, Q + (size_t)(k + 1) * nx * nx);
    n.bt.assign(b + (size_t)k * nx, b + (size_t)(k + 1) * nx); n.qt.assign(q + (size_t)k * nx, q + (size_t)(k + 1) * nx);
    n.Bt = Mat(nx, m[k]); n.Bt.a.assign(B + oB, B + oB + (size_t)nx * m[k]); oB += (size_t)nx * m[k];
    n.Rt = Mat(m[k], m[k]); n.Rt.a.assign(R + oR, R + oR + (size_t)m[k] * m[k]); oR += (size_t)m[k] * m[k];
    n.Pt = Mat(m[k], nx); n.Pt.a.assign(P + oP, P + oP + (size_t)m[k] * nx); oP += (size_t)m[k] * nx;
    n.rt.assign(r + or_, r + or_ + m[k]); or_ += m[k];
  }
  std::vector<std::vector<double>> dxv, duv;
  if (!riccati_solve(nodes, nx, std::vector<double>(dx0, dx0 + nx), dxv, duv)) return 1;
  size_t ou = 0, oK = 0;
  for (int k = 0; k <= N; ++k) for (int i = 0; i < nx; ++i) dx[(size_t)k * nx + i] = dxv[k][i];
  for (int k = 0; k < N; ++k) { for (int i = 0; i < m[k]; ++i) du[ou + i] = duv[k][i]; ou += m[k]; for (size_t i = 0; i < nodes[k].Kt.a.size(); ++i) Kt[oK + i] = nodes[k].Kt.a[i]; oK += nodes[k].Kt.a.size(); }
  return 0; ORC_CATCH(-1)
}

// ---- batch of independent instances (CPU baseline: one std::thread per host core over instances)
void* orc_batch_create(const char* model_path, int B) {
  ORC_TRY Model m = load_model(model_path); auto* b = new Batch();
  for (int i = 0; i < B; ++i) b->s.emplace_back(new Solver(m));
  b->t0.assign(B, 0.0); b->x0.assign(B, m.initial_state); b->cmd.assign(B, std::array<double, 4>{0, 0, 0, 0}); b->has_cmd.assign(B, 0); return b; ORC_CATCH(nullptr)
}
void orc_batch_destroy(void* h) { delete static_cast<Batch*>(h); }
void* orc_batch_instance(void* h, int i) { return static_cast<Batch*>(h)->s[i].get(); }
void orc_batch_set_observation(void* h, int i, double t0, const double* x0) { auto* b = static_cast<Batch*>(h); b->t0[i] = t0; b->x0[i].assign(x0, x0 + b->s[i]->M.nx); }
void orc_batch_set_cmd_vel(void* h, int i, const double* cmd, double ttt) { auto* b = static_cast<Batch*>(h); for (int k = 0; k < 4; ++k) b->cmd[i][k] = cmd[k]; b->has_cmd[i] = 1; b->ttt = ttt; }
void orc_batch_get_observation(void* h, int i, double* t0, double* x0) { auto* b = static_cast<Batch*>(h); *t0 = b->t0[i]; for (size_t k = 0; k < b->x0[i].size(); ++k) x0[k] = b->x0[i][k]; }
// runs one tick for instances [first, first+count) on `threads` threads; returns wall seconds, -1 on failure.
// shift_dt > 0: closed loop under a perfect model: t0 += shift_dt, x0 = previous optimized state at the new time; a cmd_vel target
// (if set) is rebuilt from the new observation (TargetTrajectoriesPublisher.cpp:76-99) before the solve.
double orc_batch_run(void* h, int first, int count, int threads, double shift_dt) {
  auto* b = static_cast<Batch*>(h);
  std::atomic<int> next(first); std::atomic<int> fail(0);
  const auto t_begin = std::chrono::steady_clock::now();
  auto work = [&]() {
    for (;;) { const int i = next.fetch_add(1); if (i >= first + count) break;
      try {
        Solver& S = *b->s[i];
        if (shift_dt > 0.0 && !S.sol.empty()) {
          std::vector<double> xo(S.M.nx), uo(S.M.nu); int mode;
          b->t0[i] += shift_dt;
          S.evaluatePolicy(b->t0[i], b->x0[i].data(), xo.data(), uo.data(), &mode);
          b->x0[i] = xo;
        }
        if (b->has_cmd[i]) S.P.target = cmd_vel_to_target(S.M, b->t0[i], b->x0[i].data(), b->cmd[i].data(), b->ttt);
        S.run(b->t0[i], b->x0[i]); if (S.info.status != 0) fail++;
      } catch (...) { fail++; } }
  };
  std::vector<std::thread> th; for (int t = 1; t < threads; ++t) th.emplace_back(work);
  work(); for (auto& t : th) t.join();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
  return fail.load() ? -1.0 : sec;
}

}  // extern "C"
