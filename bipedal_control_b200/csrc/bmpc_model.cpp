// Model loading for libbmpc: compact derived model files and the reference's own task.info / reference.info /
// gait.info / URDF (host side of BipedalRobotInterface, ocs2_bipedal_robot/src/BipedalRobotInterface.cpp:67-204).
#include "bmpc_model.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace bmpc {

namespace {

struct KV {
  std::map<std::string, std::vector<double>> d;
  std::map<std::string, std::vector<long>> i;
  std::map<std::string, std::string> s;
  const std::vector<double>& D(const std::string& k) const { auto it = d.find(k); if (it == d.end()) throw std::runtime_error("[bmpc] model file: missing key '" + k + "'"); return it->second; }
  const std::vector<long>& I(const std::string& k) const { auto it = i.find(k); if (it == i.end()) throw std::runtime_error("[bmpc] model file: missing key '" + k + "'"); return it->second; }
  std::string S(const std::string& k) const { auto it = s.find(k); return it == s.end() ? std::string() : it->second; }
};

KV read_kv(const std::string& path) {
  std::ifstream fh(path);
  if (!fh) throw std::invalid_argument("[bmpc] model file not found: " + path);
  KV kv; std::string line;
  while (std::getline(fh, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream is(line);
    std::string k, t; size_t n = 0;
    if (!(is >> k >> t >> n)) continue;
    if (t == "s") { std::string v; is >> v; kv.s[k] = v; }
    else if (t == "i") { std::vector<long> v(n); for (auto& e : v) is >> e; kv.i[k] = v; }
    else { std::vector<double> v(n); for (auto& e : v) { std::string tok; is >> tok; e = std::strtod(tok.c_str(), nullptr); } kv.d[k] = v; }
  }
  return kv;
}

void copy_n(const std::vector<double>& v, double* dst, size_t n) { if (v.size() < n) throw std::runtime_error("[bmpc] model file: array too short"); for (size_t i = 0; i < n; ++i) dst[i] = v[i]; }

}  // namespace

// Fills the device mirror from the host description and validates the structural assumptions of the kernels.
void finalize_model(HostModel& m) {
  const int nj = m.nj;
  if (nj != 10 && nj != 12) throw std::invalid_argument("[bmpc] unsupported number of leg joints (10 or 12 expected)");
  if ((int)m.contact_parent.size() != NCON) throw std::invalid_argument("[bmpc] exactly four 3-DoF contact points are supported");
  const int nl = nj / 2;
  for (int j = 0; j < nj; ++j) {
    const int expect = (j % nl == 0) ? -1 : j - 1;
    if (m.joint_parent[j] != expect) throw std::invalid_argument("[bmpc] legs must be two serial chains of nj/2 joints attached to the base");
  }
  for (int c = 0; c < NCON; ++c)
    if (m.contact_parent[c] != (c / 2) * nl + nl - 1) throw std::invalid_argument("[bmpc] contact points must be attached to the last link of each leg (two per foot)");
  m.nx = 12 + nj; m.nu = 12 + nj;
  if ((int)m.initial_state.size() != m.nx) throw std::invalid_argument("[bmpc] initial_state must have 12 + nj entries");
  if ((int)m.default_joint_state.size() != nj) throw std::invalid_argument("[bmpc] default_joint_state must have nj entries");
  m.dev.nj = nj; m.dev.nl = nl;
}

HostModel load_compact_model(const std::string& path) {
  KV f = read_kv(path);
  HostModel m;
  m.name = f.S("name");
  m.nj = (int)f.I("nj")[0];
  const int nc = (int)f.I("nc")[0];
  if (nc != NCON || m.nj > MAXJ) throw std::invalid_argument("[bmpc] unsupported model dimensions");
  DevModel& d = m.dev;
  std::memset(&d, 0, sizeof(d));
  d.base_mass = f.D("base_mass")[0]; copy_n(f.D("base_com"), d.base_com, 3); copy_n(f.D("base_inertia"), d.base_inertia, 9);
  for (int j = 0; j < m.nj; ++j) {
    const std::string p = "joint" + std::to_string(j) + "_";
    m.joint_names.push_back(f.S(p + "name"));
    m.joint_parent.push_back((int)f.I(p + "parent")[0]);
    copy_n(f.D(p + "R"), d.Rj[j], 9); copy_n(f.D(p + "p"), d.pj[j], 3); copy_n(f.D(p + "axis"), d.axis[j], 3);
    d.mass[j] = f.D(p + "mass")[0]; copy_n(f.D(p + "com"), d.com[j], 3); copy_n(f.D(p + "inertia"), d.inertia[j], 9);
    m.joint_lo.push_back(f.D(p + "limits")[0]); m.joint_hi.push_back(f.D(p + "limits")[1]);
  }
  for (int c = 0; c < NCON; ++c) {
    const std::string p = "contact" + std::to_string(c) + "_";
    m.contact_names.push_back(f.S(p + "name"));
    m.contact_parent.push_back((int)f.I(p + "parent")[0]);
    copy_n(f.D(p + "offset"), d.coff[c], 3);
  }
  d.total_mass = f.D("total_mass")[0];
  const int nx = 12 + m.nj;
  m.initial_state = f.D("initial_state"); m.default_joint_state = f.D("default_joint_state");
  copy_n(f.D("Q_diag"), d.Qdiag, nx); copy_n(f.D("R_force_diag"), d.Rforce, 12); copy_n(f.D("R_joint"), d.Rjoint, (size_t)m.nj * m.nj);
  m.R_taskspace_diag = f.D("R_taskspace_diag");
  m.com_height = f.D("com_height")[0]; m.target_disp_vel = f.D("target_displacement_velocity")[0]; m.target_rot_vel = f.D("target_rotation_velocity")[0];
  d.mu_f = f.D("friction_coefficient")[0]; d.bar_mu = f.D("barrier_mu")[0]; d.bar_delta = f.D("barrier_delta")[0];
  d.fr_reg = f.D("friction_regularization")[0]; d.fr_grip = f.D("friction_gripper_force")[0]; d.fr_shift = f.D("friction_hessian_shift")[0];
  d.gain = f.D("position_error_gain")[0]; m.phase_transition_stance_time = f.D("phase_transition_stance_time")[0];
  d.liftoff_vel = f.D("swing_liftoff_velocity")[0]; d.touchdown_vel = f.D("swing_touchdown_velocity")[0];
  d.swing_height = f.D("swing_height")[0]; d.swing_time_scale = f.D("swing_time_scale")[0];
  m.sqp_dt = f.D("sqp_dt")[0]; m.sqp_iterations = (int)f.I("sqp_iterations")[0]; d.delta_tol = f.D("sqp_delta_tol")[0];
  d.g_max = f.D("sqp_g_max")[0]; d.g_min = f.D("sqp_g_min")[0]; m.time_horizon = f.D("mpc_time_horizon")[0];
  m.mpc_frequency = f.D("mpc_desired_frequency")[0];
  for (long v : f.I("initial_mode_sequence")) m.init_modes.push_back((int)v);
  m.init_events = f.D("initial_event_times");
  for (long v : f.I("default_template_modes")) m.default_template.modes.push_back((int)v);
  m.default_template.times = f.D("default_template_times"); m.default_template.name = "default";
  const int ng = (int)f.I("n_gaits")[0];
  for (int g = 0; g < ng; ++g) {
    GaitTemplate t; const std::string p = "gait" + std::to_string(g) + "_";
    t.name = f.S(p + "name"); for (long v : f.I(p + "modes")) t.modes.push_back((int)v); t.times = f.D(p + "times");
    m.gaits.push_back(t);
  }
  finalize_model(m);
  return m;
}

void save_compact_model(const HostModel& m, const std::string& path) {
  FILE* fh = std::fopen(path.c_str(), "w");
  if (!fh) throw std::invalid_argument("[bmpc] cannot write " + path);
  auto wd = [&](const std::string& k, const double* v, size_t n) { std::fprintf(fh, "%s d %zu", k.c_str(), n); for (size_t i = 0; i < n; ++i) std::fprintf(fh, " %.17g", v[i]); std::fprintf(fh, "\n"); };
  auto wd1 = [&](const std::string& k, double v) { wd(k, &v, 1); };
  auto wi = [&](const std::string& k, const std::vector<int>& v) { std::fprintf(fh, "%s i %zu", k.c_str(), v.size()); for (int e : v) std::fprintf(fh, " %d", e); std::fprintf(fh, "\n"); };
  auto ws = [&](const std::string& k, const std::string& v) { std::fprintf(fh, "%s s 1 %s\n", k.c_str(), v.c_str()); };
  const DevModel& d = m.dev;
  std::fprintf(fh, "# bmpc compact model file v1 (derived numbers; written by bmpc_export_model)\n");
  ws("name", m.name); wi("nj", {m.nj}); wi("nc", {NCON});
  wd1("base_mass", d.base_mass); wd("base_com", d.base_com, 3); wd("base_inertia", d.base_inertia, 9);
  for (int j = 0; j < m.nj; ++j) {
    const std::string p = "joint" + std::to_string(j) + "_";
    ws(p + "name", m.joint_names[j]); wi(p + "parent", {m.joint_parent[j]});
    wd(p + "R", d.Rj[j], 9); wd(p + "p", d.pj[j], 3); wd(p + "axis", d.axis[j], 3); wd1(p + "mass", d.mass[j]); wd(p + "com", d.com[j], 3); wd(p + "inertia", d.inertia[j], 9);
    const double lim[2] = {m.joint_lo[j], m.joint_hi[j]}; wd(p + "limits", lim, 2);
  }
  for (int c = 0; c < NCON; ++c) {
    const std::string p = "contact" + std::to_string(c) + "_";
    ws(p + "name", m.contact_names[c]); wi(p + "parent", {m.contact_parent[c]}); wd(p + "offset", d.coff[c], 3);
  }
  wd1("total_mass", d.total_mass);
  wd("initial_state", m.initial_state.data(), m.initial_state.size());
  wd("Q_diag", d.Qdiag, m.nx); wd("R_taskspace_diag", m.R_taskspace_diag.data(), m.R_taskspace_diag.size());
  wd("R_force_diag", d.Rforce, 12); wd("R_joint", d.Rjoint, (size_t)m.nj * m.nj);
  wd("default_joint_state", m.default_joint_state.data(), m.default_joint_state.size());
  wd1("com_height", m.com_height); wd1("target_displacement_velocity", m.target_disp_vel); wd1("target_rotation_velocity", m.target_rot_vel);
  wd1("friction_coefficient", d.mu_f); wd1("barrier_mu", d.bar_mu); wd1("barrier_delta", d.bar_delta);
  wd1("friction_regularization", d.fr_reg); wd1("friction_gripper_force", d.fr_grip); wd1("friction_hessian_shift", d.fr_shift);
  wd1("position_error_gain", d.gain); wd1("phase_transition_stance_time", m.phase_transition_stance_time);
  wd1("swing_liftoff_velocity", d.liftoff_vel); wd1("swing_touchdown_velocity", d.touchdown_vel); wd1("swing_height", d.swing_height); wd1("swing_time_scale", d.swing_time_scale);
  wd1("sqp_dt", m.sqp_dt); wi("sqp_iterations", {m.sqp_iterations}); wd1("sqp_delta_tol", d.delta_tol); wd1("sqp_g_max", d.g_max); wd1("sqp_g_min", d.g_min);
  wd1("mpc_time_horizon", m.time_horizon); wd1("mpc_desired_frequency", m.mpc_frequency); wi("centroidal_model_type", {0});
  wi("initial_mode_sequence", m.init_modes); wd("initial_event_times", m.init_events.data(), m.init_events.size());
  wi("default_template_modes", m.default_template.modes); wd("default_template_times", m.default_template.times.data(), m.default_template.times.size());
  wi("n_gaits", {(int)m.gaits.size()});
  for (size_t g = 0; g < m.gaits.size(); ++g) {
    const std::string p = "gait" + std::to_string(g) + "_";
    ws(p + "name", m.gaits[g].name); wi(p + "modes", m.gaits[g].modes); wd(p + "times", m.gaits[g].times.data(), m.gaits[g].times.size());
  }
  std::fclose(fh);
}

}  // namespace bmpc
