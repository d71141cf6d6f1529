// Host-side robot/problem description and its device mirror.
//
// Replaces what BipedalRobotInterface assembles on the host (ocs2_bipedal_robot/src/BipedalRobotInterface.cpp:96-204):
// reduced Pinocchio model (createPinocchioInterface, :117), CentroidalModelInfo (:120-123), Q/R (:239-291),
// friction / swing / sqp / mpc settings (:96-101, :296-316) and the gait templates of gait.info.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace bmpc {

constexpr int MAXJ = 12;   // leg joints: H1 10, G1 12
constexpr int NCON = 4;    // 3-DoF contacts (two per foot), common/Types.h:39-41

struct GaitTemplate { std::string name; std::vector<int> modes; std::vector<double> times; };

// Plain-old-data model copied to __constant__ memory.
struct DevModel {
  int nj, nl;                       // leg joints, joints per leg (two serial chains)
  double Rj[MAXJ][9], pj[MAXJ][3], axis[MAXJ][3], mass[MAXJ], com[MAXJ][3], inertia[MAXJ][9];
  double base_mass, base_com[3], base_inertia[9];
  double coff[NCON][3];             // contact offsets in the last joint frame of leg (c / 2)
  double total_mass;
  double Qdiag[12 + MAXJ], Rforce[12], Rjoint[MAXJ * MAXJ];
  double mu_f, fr_reg, fr_grip, fr_shift, bar_mu, bar_delta, gain;
  double liftoff_vel, touchdown_vel, swing_height, swing_time_scale;
  double g_max, g_min, delta_tol;
};

struct HostModel {
  std::string name;
  DevModel dev;
  int nj = 0, nx = 0, nu = 0;
  std::vector<double> initial_state, default_joint_state;
  std::vector<double> joint_lo, joint_hi;
  std::vector<std::string> joint_names, contact_names;
  std::vector<int> joint_parent, contact_parent;
  double com_height = 0, target_disp_vel = 0, target_rot_vel = 0;
  double phase_transition_stance_time = 0.4;
  double sqp_dt = 0.015, time_horizon = 1.0, mpc_frequency = 50;
  int sqp_iterations = 1;
  std::vector<double> R_taskspace_diag;
  std::vector<int> init_modes; std::vector<double> init_events;
  GaitTemplate default_template;
  std::vector<GaitTemplate> gaits;
};

// compact model file (tools/ingest.py, bmpc_export_model); throws std::invalid_argument / std::runtime_error
HostModel load_compact_model(const std::string& path);
void save_compact_model(const HostModel& m, const std::string& path);
// the reference's own files: task.info, reference.info, gait.info (may be empty), URDF
HostModel load_reference_files(const std::string& task, const std::string& reference, const std::string& gait, const std::string& urdf);

}  // namespace bmpc
