// TEST INFRASTRUCTURE - CPU oracle for the batched bipedal MPC hot path.
//
// This directory restates, on the CPU and in plain FP64 C++, the algorithm of the reference's
// hot path (zitongbai/bipedal_control: ocs2_bipedal_robot + the un-vendored OCS2 / Pinocchio /
// HPIPM code it calls).  It is the *checker* for the CUDA product: only tests/, the smoke test and
// bench.py's cpu_baseline / --impl reference legs may load it.  The product never calls it.
//
// PARITY UNPINNED: the reference ships no tests, no golden vectors, and its solver stack (OCS2,
// Pinocchio, CppADCodeGen, HPIPM) is neither vendored nor buildable offline.
// Everything marked [UPSTREAM] is a restatement of leggedrobotics/ocs2 from its published
// algorithm; known-answer tests in tests/ pin this oracle against independent numpy computations
// (finite differences, brute-force momentum sums, dense KKT solves).
//
// oracle_model.hpp : compact model file reader + scalar-templated rigid-body kinematics and
// centroidal quantities (restates [UPSTREAM] ocs2_centroidal_model / Pinocchio as used from
// ocs2_bipedal_robot/src/BipedalRobotInterface.cpp:117-123,169-178).
#pragma once
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

constexpr int MAXJ = 12;      // leg joints (H1: 10, G1: 12)
constexpr int NC = 4;         // 3-DoF contact points (two per foot), Types.h:39-41
constexpr int MAXX = 12 + MAXJ;
constexpr int MAXU = 12 + MAXJ;

// ---------------------------------------------------------------- forward-mode dual numbers
// The reference differentiates with CppAD (exact first-order AD, BipedalRobotDynamicsAD.cpp:53-56).
// The oracle mirrors that with forward-mode duals over all nx+nu directions.
template <int N>
struct Dual {
  double v;
  double d[N];
  Dual() : v(0.0) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
  Dual(double a) : v(a) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
};
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; const double ib = 1.0 / b.v; r.v = a.v * ib; for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
template <int N> inline Dual<N> operator*(double a, const Dual<N>& b) { Dual<N> r; r.v = a * b.v; for (int i = 0; i < N; ++i) r.d[i] = a * b.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& b, double a) { return a * b; }
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double b, const Dual<N>& a) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { return a * (1.0 / b); }
template <int N> inline Dual<N>& operator+=(Dual<N>& a, const Dual<N>& b) { a.v += b.v; for (int i = 0; i < N; ++i) a.d[i] += b.d[i]; return a; }
template <int N> inline Dual<N>& operator-=(Dual<N>& a, const Dual<N>& b) { a.v -= b.v; for (int i = 0; i < N; ++i) a.d[i] -= b.d[i]; return a; }
template <int N> inline Dual<N> sin(const Dual<N>& a) { Dual<N> r; r.v = std::sin(a.v); const double c = std::cos(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
template <int N> inline Dual<N> cos(const Dual<N>& a) { Dual<N> r; r.v = std::cos(a.v); const double s = -std::sin(a.v); for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
// ---------------------------------------------------------------- counting scalar (exact flop counts instead of estimates)
// A double that counts the arithmetic it takes part in.  flow_map<Counted> gives the exact operation count of one evaluation of the model;
// the count of the forward-mode linearisation follows from it (every + / - costs 1 + N, every * costs 1 + 3N flops with N tangent directions).
struct FlopCount { unsigned long long add = 0, mul = 0, div = 0, trig = 0; };
inline FlopCount& flop_counter() { static thread_local FlopCount c; return c; }
struct Counted {
  double v;
  Counted() : v(0.0) {}
  Counted(double a) : v(a) {}
};
inline Counted operator+(const Counted& a, const Counted& b) { ++flop_counter().add; return Counted(a.v + b.v); }
inline Counted operator-(const Counted& a, const Counted& b) { ++flop_counter().add; return Counted(a.v - b.v); }
inline Counted operator-(const Counted& a) { return Counted(-a.v); }
inline Counted operator*(const Counted& a, const Counted& b) { ++flop_counter().mul; return Counted(a.v * b.v); }
inline Counted operator/(const Counted& a, const Counted& b) { ++flop_counter().div; return Counted(a.v / b.v); }
inline Counted operator*(double a, const Counted& b) { ++flop_counter().mul; return Counted(a * b.v); }
inline Counted operator*(const Counted& b, double a) { ++flop_counter().mul; return Counted(a * b.v); }
inline Counted operator+(const Counted& a, double b) { ++flop_counter().add; return Counted(a.v + b); }
inline Counted operator+(double b, const Counted& a) { ++flop_counter().add; return Counted(a.v + b); }
inline Counted operator-(const Counted& a, double b) { ++flop_counter().add; return Counted(a.v - b); }
inline Counted operator/(const Counted& a, double b) { ++flop_counter().mul; return Counted(a.v * (1.0 / b)); }
inline Counted& operator+=(Counted& a, const Counted& b) { ++flop_counter().add; a.v += b.v; return a; }
inline Counted& operator-=(Counted& a, const Counted& b) { ++flop_counter().add; a.v -= b.v; return a; }
inline Counted sin(const Counted& a) { ++flop_counter().trig; return Counted(std::sin(a.v)); }
inline Counted cos(const Counted& a) { ++flop_counter().trig; return Counted(std::cos(a.v)); }
inline double value_of(const Counted& a) { return a.v; }
inline double sin(double a) { return std::sin(a); }
inline double cos(double a) { return std::cos(a); }
inline double value_of(double a) { return a; }
template <int N> inline double value_of(const Dual<N>& a) { return a.v; }

// ---------------------------------------------------------------- tiny 3-vector / 3x3 templates
template <class S> struct V3 {
  S x, y, z;
  V3() : x(0.0), y(0.0), z(0.0) {}
  V3(S a, S b, S c) : x(a), y(b), z(c) {}
  S& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  const S& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class S> inline V3<S> operator+(const V3<S>& a, const V3<S>& b) { return V3<S>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class S> inline V3<S> operator-(const V3<S>& a, const V3<S>& b) { return V3<S>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class S> inline V3<S> operator*(const S& s, const V3<S>& a) { return V3<S>(s * a.x, s * a.y, s * a.z); }
template <class S> inline V3<S> cross(const V3<S>& a, const V3<S>& b) { return V3<S>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
template <class S> inline S dot(const V3<S>& a, const V3<S>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class S> struct M3 {
  S m[3][3];
  M3() { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = S(0.0); }
  static M3 identity() { M3 r; for (int i = 0; i < 3; ++i) r.m[i][i] = S(1.0); return r; }
};
template <class S> inline M3<S> operator*(const M3<S>& a, const M3<S>& b) { M3<S> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { S s(0.0); for (int k = 0; k < 3; ++k) s = s + a.m[i][k] * b.m[k][j]; r.m[i][j] = s; } return r; }
template <class S> inline V3<S> operator*(const M3<S>& a, const V3<S>& v) { return V3<S>(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z, a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z); }
template <class S> inline M3<S> transpose(const M3<S>& a) { M3<S> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i]; return r; }
template <class S, class T> inline M3<S> castM(const M3<T>& a) { M3<S> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = S(a.m[i][j]); return r; }
template <class S, class T> inline V3<S> castV(const V3<T>& a) { return V3<S>(S(a.x), S(a.y), S(a.z)); }

// Rodrigues rotation about a constant unit axis.
template <class S> inline M3<S> rot_axis(const V3<double>& a, const S& ang) {
  const S s = sin(ang), c = cos(ang);
  const S t = S(1.0) - c;
  M3<S> r;
  r.m[0][0] = c + t * (a.x * a.x); r.m[0][1] = t * (a.x * a.y) - s * a.z; r.m[0][2] = t * (a.x * a.z) + s * a.y;
  r.m[1][0] = t * (a.x * a.y) + s * a.z; r.m[1][1] = c + t * (a.y * a.y); r.m[1][2] = t * (a.y * a.z) - s * a.x;
  r.m[2][0] = t * (a.x * a.z) - s * a.y; r.m[2][1] = t * (a.y * a.z) + s * a.x; r.m[2][2] = c + t * (a.z * a.z);
  return r;
}

// ---------------------------------------------------------------- model
struct GaitTemplate { std::string name; std::vector<int> modes; std::vector<double> times; };

struct Model {
  std::string name;
  int nj = 0, nc = NC, nx = 0, nu = 0, nq = 0;
  double base_mass = 0; V3<double> base_com; M3<double> base_I;
  struct Joint { int parent; M3<double> R; V3<double> p; V3<double> axis; double mass; V3<double> com; M3<double> I; double lo, hi; };
  Joint joints[MAXJ];
  struct Contact { int parent; V3<double> off; };
  Contact contacts[NC];
  double total_mass = 0;
  std::vector<double> initial_state, Q_diag, R_force_diag, R_joint, default_joint_state;
  double com_height = 0, target_disp_vel = 0, target_rot_vel = 0;
  double mu_f = 0.5, barrier_mu = 0.1, barrier_delta = 5.0, fr_reg = 25.0, fr_grip = 0.0, fr_shift = 1e-6;
  double pos_err_gain = 0.0, phase_transition_stance_time = 0.4;
  double liftoff_vel = 0.05, touchdown_vel = 0.0, swing_height = 0.05, swing_time_scale = 0.15;
  double sqp_dt = 0.015, delta_tol = 1e-4, g_max = 1e-2, g_min = 1e-6, time_horizon = 1.0;
  int sqp_iterations = 1;
  std::vector<int> init_modes; std::vector<double> init_events;
  GaitTemplate default_template;
  std::vector<GaitTemplate> gaits;
};

struct ModelFile {
  std::map<std::string, std::vector<double>> d;
  std::map<std::string, std::vector<long>> i;
  std::map<std::string, std::string> s;
  explicit ModelFile(const std::string& path) {
    std::ifstream fh(path);
    if (!fh) throw std::invalid_argument("[oracle] model file not found: " + path);
    std::string line;
    while (std::getline(fh, line)) {
      if (line.empty() || line[0] == '#') continue;
      std::istringstream is(line);
      std::string k, t; size_t n;
      is >> k >> t >> n;
      if (t == "s") { std::string v; is >> v; s[k] = v; }
      else if (t == "i") { std::vector<long> v(n); for (auto& e : v) is >> e; i[k] = v; }
      else { std::vector<double> v(n); for (auto& e : v) { std::string tok; is >> tok; e = std::strtod(tok.c_str(), nullptr); } d[k] = v; }
    }
  }
  const std::vector<double>& D(const std::string& k) const { auto it = d.find(k); if (it == d.end()) throw std::runtime_error("[oracle] missing key " + k); return it->second; }
  const std::vector<long>& I(const std::string& k) const { auto it = i.find(k); if (it == i.end()) throw std::runtime_error("[oracle] missing key " + k); return it->second; }
  double D1(const std::string& k) const { return D(k)[0]; }
  long I1(const std::string& k) const { return I(k)[0]; }
};

inline Model load_model(const std::string& path) {
  ModelFile f(path);
  Model m;
  m.name = f.s.count("name") ? f.s["name"] : "robot";
  m.nj = (int)f.I1("nj");
  m.nc = (int)f.I1("nc");
  if (m.nc != NC || m.nj > MAXJ) throw std::runtime_error("[oracle] unsupported model dimensions");
  m.nx = 12 + m.nj; m.nu = 3 * m.nc + m.nj; m.nq = 6 + m.nj;
  auto v3 = [](const std::vector<double>& v) { return V3<double>(v[0], v[1], v[2]); };
  auto m3 = [](const std::vector<double>& v) { M3<double> r; for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) r.m[a][b] = v[3 * a + b]; return r; };
  m.base_mass = f.D1("base_mass"); m.base_com = v3(f.D("base_com")); m.base_I = m3(f.D("base_inertia"));
  for (int j = 0; j < m.nj; ++j) {
    const std::string p = "joint" + std::to_string(j) + "_";
    auto& J = m.joints[j];
    J.parent = (int)f.I1(p + "parent"); J.R = m3(f.D(p + "R")); J.p = v3(f.D(p + "p")); J.axis = v3(f.D(p + "axis"));
    J.mass = f.D1(p + "mass"); J.com = v3(f.D(p + "com")); J.I = m3(f.D(p + "inertia"));
    J.lo = f.D(p + "limits")[0]; J.hi = f.D(p + "limits")[1];
  }
  for (int c = 0; c < m.nc; ++c) {
    const std::string p = "contact" + std::to_string(c) + "_";
    m.contacts[c].parent = (int)f.I1(p + "parent"); m.contacts[c].off = v3(f.D(p + "offset"));
  }
  m.total_mass = f.D1("total_mass");
  m.initial_state = f.D("initial_state"); m.Q_diag = f.D("Q_diag"); m.R_force_diag = f.D("R_force_diag"); m.R_joint = f.D("R_joint");
  m.default_joint_state = f.D("default_joint_state");
  m.com_height = f.D1("com_height"); m.target_disp_vel = f.D1("target_displacement_velocity"); m.target_rot_vel = f.D1("target_rotation_velocity");
  m.mu_f = f.D1("friction_coefficient"); m.barrier_mu = f.D1("barrier_mu"); m.barrier_delta = f.D1("barrier_delta");
  m.fr_reg = f.D1("friction_regularization"); m.fr_grip = f.D1("friction_gripper_force"); m.fr_shift = f.D1("friction_hessian_shift");
  m.pos_err_gain = f.D1("position_error_gain"); m.phase_transition_stance_time = f.D1("phase_transition_stance_time");
  m.liftoff_vel = f.D1("swing_liftoff_velocity"); m.touchdown_vel = f.D1("swing_touchdown_velocity");
  m.swing_height = f.D1("swing_height"); m.swing_time_scale = f.D1("swing_time_scale");
  m.sqp_dt = f.D1("sqp_dt"); m.sqp_iterations = (int)f.I1("sqp_iterations"); m.delta_tol = f.D1("sqp_delta_tol");
  m.g_max = f.D1("sqp_g_max"); m.g_min = f.D1("sqp_g_min"); m.time_horizon = f.D1("mpc_time_horizon");
  for (long v : f.I("initial_mode_sequence")) m.init_modes.push_back((int)v);
  m.init_events = f.D("initial_event_times");
  for (long v : f.I("default_template_modes")) m.default_template.modes.push_back((int)v);
  m.default_template.times = f.D("default_template_times");
  m.default_template.name = "default";
  const int ng = (int)f.I1("n_gaits");
  for (int g = 0; g < ng; ++g) {
    GaitTemplate t; const std::string p = "gait" + std::to_string(g) + "_";
    t.name = f.s[p + "name"]; for (long v : f.I(p + "modes")) t.modes.push_back((int)v); t.times = f.D(p + "times");
    m.gaits.push_back(t);
  }
  return m;
}

// ---------------------------------------------------------------- kinematics
// q = [base xyz, ZYX Euler (yaw,pitch,roll), joints]; generalized velocity v = dq/dt (Euler RATES for the
// base orientation, as the Pinocchio composite root joint Translation+SphericalZYX has) [UPSTREAM,
// SURVEY.md Appendix B.3].  Everything is expressed in the world frame.
template <class S>
struct Kin {
  int nj = 0, nq = 0;
  M3<S> Rb; V3<S> pb;
  M3<S> Rw[MAXJ]; V3<S> ow[MAXJ]; V3<S> aw[MAXJ];   // joint frame rotation / origin / axis in world
  V3<S> cbody[MAXJ + 1]; M3<S> Ibody[MAXJ + 1]; double mbody[MAXJ + 1];  // body index 0 = base, j+1 = joint j body
  V3<S> pc[NC];                                      // contact points
  V3<S> com;
  // the 6 base DoF are written as 1-DoF joints: 3 prismatic (world axes) then revolute z, y', x''
  V3<S> base_axis[6];
};

template <class S>
inline M3<S> euler_zyx(const S& z, const S& y, const S& x) {
  const S cz = cos(z), sz = sin(z), cy = cos(y), sy = sin(y), cx = cos(x), sx = sin(x);
  M3<S> R;
  R.m[0][0] = cz * cy; R.m[0][1] = cz * sy * sx - sz * cx; R.m[0][2] = cz * sy * cx + sz * sx;
  R.m[1][0] = sz * cy; R.m[1][1] = sz * sy * sx + cz * cx; R.m[1][2] = sz * sy * cx - cz * sx;
  R.m[2][0] = -sy;     R.m[2][1] = cy * sx;                R.m[2][2] = cy * cx;
  return R;
}

template <class S>
inline void forward_kinematics(const Model& M, const S* q, Kin<S>& K) {
  K.nj = M.nj; K.nq = M.nq;
  K.pb = V3<S>(q[0], q[1], q[2]);
  K.Rb = euler_zyx<S>(q[3], q[4], q[5]);
  // Euler-rate axes in world: e_z, Rz e_y, Rz Ry e_x
  const S cz = cos(q[3]), sz = sin(q[3]), cy = cos(q[4]), sy = sin(q[4]);
  K.base_axis[0] = V3<S>(S(1.0), S(0.0), S(0.0));
  K.base_axis[1] = V3<S>(S(0.0), S(1.0), S(0.0));
  K.base_axis[2] = V3<S>(S(0.0), S(0.0), S(1.0));
  K.base_axis[3] = V3<S>(S(0.0), S(0.0), S(1.0));
  K.base_axis[4] = V3<S>(-sz, cz, S(0.0));
  K.base_axis[5] = V3<S>(cz * cy, sz * cy, -sy);
  K.mbody[0] = M.base_mass;
  K.cbody[0] = K.Rb * castV<S>(M.base_com) + K.pb;
  K.Ibody[0] = K.Rb * castM<S>(M.base_I) * transpose(K.Rb);
  for (int j = 0; j < M.nj; ++j) {
    const auto& J = M.joints[j];
    const M3<S>& Rp = J.parent < 0 ? K.Rb : K.Rw[J.parent];
    const V3<S>& pp = J.parent < 0 ? K.pb : K.ow[J.parent];
    K.ow[j] = Rp * castV<S>(J.p) + pp;
    const M3<S> Rfix = Rp * castM<S>(J.R);
    K.aw[j] = Rfix * castV<S>(J.axis);
    K.Rw[j] = Rfix * rot_axis<S>(J.axis, q[6 + j]);
    K.mbody[j + 1] = J.mass;
    K.cbody[j + 1] = K.Rw[j] * castV<S>(J.com) + K.ow[j];
    K.Ibody[j + 1] = K.Rw[j] * castM<S>(J.I) * transpose(K.Rw[j]);
  }
  for (int c = 0; c < M.nc; ++c) {
    const int par = M.contacts[c].parent;
    const M3<S>& Rp = par < 0 ? K.Rb : K.Rw[par];
    const V3<S>& pp = par < 0 ? K.pb : K.ow[par];
    K.pc[c] = Rp * castV<S>(M.contacts[c].off) + pp;
  }
  V3<S> mc;
  for (int b = 0; b <= M.nj; ++b) mc = mc + S(K.mbody[b]) * K.cbody[b];
  K.com = S(1.0 / M.total_mass) * mc;
}

// is leg joint k an ancestor-or-self of leg joint j ?
inline bool is_ancestor(const Model& M, int k, int j) { while (j >= 0) { if (j == k) return true; j = M.joints[j].parent; } return false; }

// Centroidal momentum matrix A(q) (6 x nq): rows 0..2 linear momentum, rows 3..5 angular momentum about the
// COM, world axes; column k = momentum produced by unit generalized velocity k  [UPSTREAM pinocchio::computeCentroidalMap].
// Brute-force (definition) form: sum over the bodies moved by joint k.
template <class S>
inline void centroidal_momentum_matrix(const Model& M, const Kin<S>& K, S* A /* 6 x nq row-major */) {
  const int nq = M.nq;
  for (int i = 0; i < 6 * nq; ++i) A[i] = S(0.0);
  auto add_body = [&](int col, int b, const V3<S>& w, const V3<S>& cdot) {
    const S mb(K.mbody[b]);
    const V3<S> lin = mb * cdot;
    const V3<S> ang = K.Ibody[b] * w + cross(K.cbody[b] - K.com, lin);
    for (int r = 0; r < 3; ++r) { A[r * nq + col] += lin[r]; A[(3 + r) * nq + col] += ang[r]; }
  };
  for (int k = 0; k < 3; ++k)  // base translation moves every body
    for (int b = 0; b <= M.nj; ++b) add_body(k, b, V3<S>(), K.base_axis[k]);
  for (int k = 3; k < 6; ++k)  // Euler rates rotate every body about the base origin
    for (int b = 0; b <= M.nj; ++b) add_body(k, b, K.base_axis[k], cross(K.base_axis[k], K.cbody[b] - K.pb));
  for (int k = 0; k < M.nj; ++k)
    for (int j = 0; j < M.nj; ++j)
      if (is_ancestor(M, k, j)) add_body(6 + k, j + 1, K.aw[k], cross(K.aw[k], K.cbody[j + 1] - K.ow[k]));
}

// 3x3 inverse (adjugate), scalar-templated.
template <class S>
inline M3<S> inverse3(const M3<S>& a) {
  M3<S> r;
  const S c00 = a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1];
  const S c01 = a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2];
  const S c02 = a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0];
  const S det = a.m[0][0] * c00 + a.m[0][1] * c01 + a.m[0][2] * c02;
  const S id = S(1.0) / det;
  r.m[0][0] = c00 * id; r.m[0][1] = (a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2]) * id; r.m[0][2] = (a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1]) * id;
  r.m[1][0] = c01 * id; r.m[1][1] = (a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0]) * id; r.m[1][2] = (a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2]) * id;
  r.m[2][0] = c02 * id; r.m[2][1] = (a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1]) * id; r.m[2][2] = (a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0]) * id;
  return r;
}

// Generalized velocity v = [A_b^-1 (m h - A_j qd_j) ; qd_j]  [UPSTREAM CentroidalModelPinocchioMapping::
// getPinocchioJointVelocity with computeFloatingBaseCentroidalMomentumMatrixInverse: A_b = [m I, A12; 0, A22]].
template <class S>
inline void generalized_velocity(const Model& M, const S* A, const S* x, const S* u, S* v /* nq */) {
  const int nq = M.nq, nj = M.nj;
  S mom[6];
  for (int r = 0; r < 6; ++r) {
    S s = S(M.total_mass) * x[r];
    for (int j = 0; j < nj; ++j) s -= A[r * nq + 6 + j] * u[3 * M.nc + j];
    mom[r] = s;
  }
  M3<S> A22, A12;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { A22.m[r][c] = A[(3 + r) * nq + 3 + c]; A12.m[r][c] = A[r * nq + 3 + c]; }
  const M3<S> A22i = inverse3(A22);
  const V3<S> wrate = A22i * V3<S>(mom[3], mom[4], mom[5]);   // Euler rates
  const V3<S> t = A12 * wrate;
  const S im(1.0 / value_of(A[0]));  // A(0,0) = mass (constant)
  v[0] = im * (mom[0] - t.x); v[1] = im * (mom[1] - t.y); v[2] = im * (mom[2] - t.z);
  v[3] = wrate.x; v[4] = wrate.y; v[5] = wrate.z;
  for (int j = 0; j < nj; ++j) v[6 + j] = u[3 * M.nc + j];
}

// Linear velocity of contact point c (LOCAL_WORLD_ALIGNED) for generalized velocity v  [UPSTREAM
// PinocchioEndEffectorKinematicsCppAd::getVelocityCppAd].
template <class S>
inline V3<S> contact_velocity(const Model& M, const Kin<S>& K, int c, const S* v) {
  V3<S> vel(v[0], v[1], v[2]);
  for (int k = 3; k < 6; ++k) vel = vel + v[k] * cross(K.base_axis[k], K.pc[c] - K.pb);
  int j = M.contacts[c].parent;
  while (j >= 0) { vel = vel + v[6 + j] * cross(K.aw[j], K.pc[c] - K.ow[j]); j = M.joints[j].parent; }
  return vel;
}

// Flow map of the full centroidal model  [UPSTREAM PinocchioCentroidalDynamicsAD::getValueCppAd; call site
// ocs2_bipedal_robot/src/dynamics/BipedalRobotDynamicsAD.cpp:46-48].  Optionally also returns contact positions and velocities.
template <class S>
inline void flow_map(const Model& M, const S* x, const S* u, S* f, V3<S>* pos = nullptr, V3<S>* vel = nullptr) {
  Kin<S> K;
  forward_kinematics<S>(M, x + 6, K);
  S A[6 * (6 + MAXJ)];
  centroidal_momentum_matrix<S>(M, K, A);
  const double m = M.total_mass;
  V3<S> lin(S(0.0), S(0.0), S(-9.81 * m)), ang;
  for (int c = 0; c < M.nc; ++c) {
    const V3<S> F(u[3 * c], u[3 * c + 1], u[3 * c + 2]);
    lin = lin + F;
    ang = ang + cross(K.pc[c] - K.com, F);
  }
  for (int r = 0; r < 3; ++r) { f[r] = lin[r] / m; f[3 + r] = ang[r] / m; }
  S v[6 + MAXJ];
  generalized_velocity<S>(M, A, x, u, v);
  for (int i = 0; i < M.nq; ++i) f[6 + i] = v[i];
  if (pos) for (int c = 0; c < M.nc; ++c) pos[c] = K.pc[c];
  if (vel) for (int c = 0; c < M.nc; ++c) vel[c] = contact_velocity<S>(M, K, c, v);
}

}  // namespace orc
