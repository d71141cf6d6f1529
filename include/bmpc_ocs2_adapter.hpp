// bmpc_ocs2_adapter.hpp - header-only C++ adapter that puts libbmpc.so behind the reference's own solver boundary.
//
// The reference holds its solver as `std::shared_ptr<ocs2::MPC_BASE> mpc_` (bipedal_controllers/include/bipedal_controllers/BipedalController.h:92),
// builds it in the virtual `setupMpc()` (BipedalController.h:67, src/BipedalController.cpp:291-308:
//     mpc_ = std::make_shared<SqpMpc>(mpcSettings, sqpSettings, ocp, initializer);
//     mpc_->getSolverPtr()->setReferenceManager(rosReferenceManagerPtr);
//     mpc_->getSolverPtr()->addSynchronizedModule(gaitReceiverPtr); )
// and drives it through MPC_MRT_Interface (BipedalController.cpp:321-351).  This header provides
//     bmpc::BmpcSolver : ocs2::SolverBase      (replaces ocs2::SqpSolver  [UPSTREAM])
//     bmpc::BmpcMpc    : ocs2::MPC_BASE        (replaces ocs2::SqpMpc     [UPSTREAM])
// so that the one-line change
//     mpc_ = std::make_shared<bmpc::BmpcMpc>(bipedalInterface_->mpcSettings(), bipedalInterface_->getOptimalControlProblem(), taskFile, urdfFile, referenceFile, gaitFile);
// swaps the CPU SQP for the B200 library while MPC_MRT_Interface, the RosReferenceManager, the GaitReceiver, bipedal_wbc and the
// visualizers keep consuming ModeSchedule / TargetTrajectories / PrimalSolution / LinearController unchanged.
//
// What crosses the boundary per MPC tick (SolverBase::run [UPSTREAM]: preRun -> runImpl -> postRun):
//   in : initTime, initState; ReferenceManager::getModeSchedule() (already updated by SwitchedModelReferenceManager::modifyReferences in preRun,
//        ocs2_bipedal_robot/src/reference_manager/SwitchedModelReferenceManager.cpp:62-69) and getTargetTrajectories()
//   out: PrimalSolution{timeTrajectory_, postEventIndices_, stateTrajectory_, inputTrajectory_, modeSchedule_, controllerPtr_ = LinearController(t, uff, K)},
//        PerformanceIndex{cost, dynamicsViolationSSE, equalityConstraintsSSE}
// `replicas` > 1 solves the same problem for B copies of the observation (optionally perturbed by the caller through replicaStateOffsets());
// replica 0 is the one handed back to OCS2.
//
// Only the OCS2 API named above is used (scalar_t / vector_t / matrix_t element access, size(), resize()); tests/stubs/ holds minimal
// stand-ins of those headers so that this file is syntax- and run-checked in CI without an OCS2 installation (tests/test_adapter.py).
#pragma once

#if !defined(BMPC_OCS2_ADAPTER_FORCE) && defined(__has_include)
#if !__has_include(<ocs2_mpc/MPC_BASE.h>)
#define BMPC_OCS2_ADAPTER_DISABLED 1
#endif
#endif

#ifndef BMPC_OCS2_ADAPTER_DISABLED

#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <ocs2_core/Types.h>
#include <ocs2_core/control/LinearController.h>
#include <ocs2_core/reference/ModeSchedule.h>
#include <ocs2_core/reference/TargetTrajectories.h>
#include <ocs2_mpc/MPC_BASE.h>
#include <ocs2_oc/oc_data/PerformanceIndex.h>
#include <ocs2_oc/oc_data/PrimalSolution.h>
#include <ocs2_oc/oc_problem/OptimalControlProblem.h>
#include <ocs2_oc/oc_solver/SolverBase.h>

#include "bmpc.h"

namespace bmpc {

class BmpcSolver final : public ocs2::SolverBase {
 public:
  struct Files { std::string task, urdf, reference, gait, model; };   // either `model` (compact file) or task + urdf + reference (+ gait)

  // ocp: kept only to answer getOptimalControlProblem() (observers / visualizers ask for it); the library builds its own problem from the files.
  BmpcSolver(const ocs2::OptimalControlProblem& ocp, const Files& files, int replicas = 1, int device = 0) : ocp_(ocp) {
    bmpc_config cfg{};
    cfg.model_file = files.model.empty() ? nullptr : files.model.c_str();
    cfg.task_file = files.task.empty() ? nullptr : files.task.c_str();
    cfg.urdf_file = files.urdf.empty() ? nullptr : files.urdf.c_str();
    cfg.reference_file = files.reference.empty() ? nullptr : files.reference.c_str();
    cfg.gait_file = files.gait.empty() ? nullptr : files.gait.c_str();
    cfg.batch = replicas; cfg.device = device;
    if (bmpc_create(&cfg, &h_) != BMPC_OK) throw std::invalid_argument(std::string("[BmpcSolver] ") + bmpc_last_error(nullptr));   // BipedalRobotInterface.cpp:71-90 throws the same type
    bmpc_get_dims(h_, &nx_, &nu_, &batch_, &maxNodes_);
    offsets_.assign(static_cast<size_t>(batch_) * nx_, 0.0);
  }
  ~BmpcSolver() override { if (h_) bmpc_destroy(h_); }
  BmpcSolver(const BmpcSolver&) = delete;
  BmpcSolver& operator=(const BmpcSolver&) = delete;

  // ---- SolverBase [UPSTREAM]
  void reset() override {
    check(bmpc_reset(h_, -1));
    primalSolution_ = ocs2::PrimalSolution();
    performanceLog_.clear();
    numIterations_ = 0;
  }
  ocs2::scalar_t getFinalTime() const override { return primalSolution_.timeTrajectory_.empty() ? 0.0 : primalSolution_.timeTrajectory_.back(); }
  void getPrimalSolution(ocs2::scalar_t /*finalTime*/, ocs2::PrimalSolution* primalSolutionPtr) const override { *primalSolutionPtr = primalSolution_; }
  size_t getNumIterations() const override { return numIterations_; }
  const ocs2::OptimalControlProblem& getOptimalControlProblem() const override { return ocp_; }
  const ocs2::PerformanceIndex& getPerformanceIndeces() const override { return performanceLog_.back(); }
  const std::vector<ocs2::PerformanceIndex>& getIterationsLog() const override {
    if (performanceLog_.empty()) throw std::runtime_error("[BmpcSolver]: No performance log yet, no problem solved yet?");
    return performanceLog_;
  }
  // as ocs2::SqpSolver: not available
  ocs2::ScalarFunctionQuadraticApproximation getValueFunction(ocs2::scalar_t, const ocs2::vector_t&) const override { throw std::runtime_error("[BmpcSolver] getValueFunction() not available."); }
  ocs2::ScalarFunctionQuadraticApproximation getHamiltonian(ocs2::scalar_t, const ocs2::vector_t&, const ocs2::vector_t&) override { throw std::runtime_error("[BmpcSolver] getHamiltonian() not available."); }
  ocs2::vector_t getStateInputEqualityConstraintLagrangian(ocs2::scalar_t, const ocs2::vector_t&) const override { throw std::runtime_error("[BmpcSolver] getStateInputEqualityConstraintLagrangian() not available."); }

  // ---- batched extras
  bmpc_handle* handle() const { return h_; }
  int replicas() const { return batch_; }
  // per-replica offsets added to the observation (replica 0 should stay zero): robustness sweeps around the measured state
  std::vector<double>& replicaStateOffsets() { return offsets_; }

 private:
  void check(int rc) const { if (rc < 0) throw std::runtime_error(std::string("[BmpcSolver] ") + bmpc_last_error(h_)); }

  void runImpl(ocs2::scalar_t initTime, const ocs2::vector_t& initState, ocs2::scalar_t finalTime) override {
    if (static_cast<int>(initState.size()) != nx_) throw std::runtime_error("[BmpcSolver] state dimension mismatch");
    (void)finalTime;   // the horizon is mpc.timeHorizon of the task file, as in MPC_BASE::run [UPSTREAM]
    // observation (MPC_MRT_Interface::setCurrentObservation -> MPC_BASE::run(t, x))
    std::vector<double> t(batch_, initTime), x(static_cast<size_t>(batch_) * nx_);
    for (int b = 0; b < batch_; ++b)
      for (int i = 0; i < nx_; ++i) x[static_cast<size_t>(b) * nx_ + i] = initState(i) + offsets_[static_cast<size_t>(b) * nx_ + i];
    check(bmpc_set_observations(h_, t.data(), x.data()));
    // references: the reference manager was updated by SolverBase::preRun (gait from GaitReceiver, targets from RosReferenceManager)
    const ocs2::ModeSchedule& ms = getReferenceManager().getModeSchedule();
    const ocs2::TargetTrajectories& tt = getReferenceManager().getTargetTrajectories();
    pushModeSchedule(ms);
    pushTargets(tt, initTime, initState);
    check(bmpc_advance(h_));          // throws on a numerical failure: caught by the MPC thread (BipedalController.cpp:344-348)
    pullPrimalSolution(ms);
    ++numIterations_;
  }
  void runImpl(ocs2::scalar_t initTime, const ocs2::vector_t& initState, ocs2::scalar_t finalTime, const ocs2::ControllerBase* externalControllerPtr) override {
    if (externalControllerPtr != nullptr) throw std::runtime_error("[BmpcSolver::run] This solver does not support external controller!");
    runImpl(initTime, initState, finalTime);
  }

  void pushModeSchedule(const ocs2::ModeSchedule& ms) {
    const int ne = static_cast<int>(ms.eventTimes.size());
    std::vector<int> n(batch_, ne), modes(static_cast<size_t>(batch_) * (ne + 1));
    std::vector<double> ev(static_cast<size_t>(batch_) * std::max(ne, 1));
    for (int b = 0; b < batch_; ++b) {
      for (int i = 0; i < ne; ++i) ev[static_cast<size_t>(b) * ne + i] = ms.eventTimes[i];
      for (int i = 0; i <= ne; ++i) modes[static_cast<size_t>(b) * (ne + 1) + i] = static_cast<int>(ms.modeSequence[i]);
    }
    check(bmpc_set_mode_schedules(h_, ne, n.data(), ev.data(), modes.data()));
  }
  void pushTargets(const ocs2::TargetTrajectories& tt, ocs2::scalar_t initTime, const ocs2::vector_t& initState) {
    int npts = static_cast<int>(tt.timeTrajectory.size());
    std::vector<double> times, states;
    if (npts == 0) {   // no target yet: hold the current state (what BipedalController::starting sends, BipedalController.cpp:145)
      npts = 1; times.assign(batch_, initTime); states.resize(static_cast<size_t>(batch_) * nx_);
      for (int b = 0; b < batch_; ++b) for (int i = 0; i < nx_; ++i) states[static_cast<size_t>(b) * nx_ + i] = initState(i);
    } else {
      times.resize(static_cast<size_t>(batch_) * npts); states.resize(static_cast<size_t>(batch_) * npts * nx_);
      for (int b = 0; b < batch_; ++b)
        for (int k = 0; k < npts; ++k) {
          times[static_cast<size_t>(b) * npts + k] = tt.timeTrajectory[k];
          for (int i = 0; i < nx_; ++i) states[(static_cast<size_t>(b) * npts + k) * nx_ + i] = tt.stateTrajectory[k](i);
        }
    }
    check(bmpc_set_target_trajectories(h_, npts, times.data(), states.data()));
  }
  void pullPrimalSolution(const ocs2::ModeSchedule& ms) {
    int n = 0;
    std::vector<double> tm(maxNodes_), xs(static_cast<size_t>(maxNodes_) * nx_), us(static_cast<size_t>(maxNodes_) * nu_), uff(static_cast<size_t>(maxNodes_) * nu_),
        K(static_cast<size_t>(maxNodes_) * nu_ * nx_), perf(static_cast<size_t>(batch_) * 8);
    std::vector<int> ev(maxNodes_);
    check(bmpc_get_policy(h_, 0, 1, &n, tm.data(), ev.data(), xs.data(), us.data(), uff.data(), K.data()));
    check(bmpc_get_performance(h_, perf.data()));
    ocs2::PrimalSolution sol;
    sol.modeSchedule_ = ms;
    sol.timeTrajectory_.assign(tm.begin(), tm.begin() + n);
    sol.stateTrajectory_.resize(n); sol.inputTrajectory_.resize(n);
    ocs2::vector_array_t bias(n); ocs2::matrix_array_t gain(n);
    for (int k = 0; k < n; ++k) {
      if (ev[k] == 2) sol.postEventIndices_.push_back(static_cast<size_t>(k));   // [UPSTREAM] toPrimalSolution / getPostEventIndices
      sol.stateTrajectory_[k].resize(nx_); sol.inputTrajectory_[k].resize(nu_); bias[k].resize(nu_); gain[k].resize(nu_, nx_);
      for (int i = 0; i < nx_; ++i) sol.stateTrajectory_[k](i) = xs[static_cast<size_t>(k) * nx_ + i];
      for (int i = 0; i < nu_; ++i) {
        sol.inputTrajectory_[k](i) = us[static_cast<size_t>(k) * nu_ + i];
        bias[k](i) = uff[static_cast<size_t>(k) * nu_ + i];
        for (int j = 0; j < nx_; ++j) gain[k](i, j) = K[(static_cast<size_t>(k) * nu_ + i) * nx_ + j];
      }
    }
    sol.controllerPtr_.reset(new ocs2::LinearController(sol.timeTrajectory_, std::move(bias), std::move(gain)));
    primalSolution_ = std::move(sol);
    ocs2::PerformanceIndex before, after;   // baseline and accepted step, as SqpSolver logs them
    before.cost = perf[0]; before.dynamicsViolationSSE = perf[1]; before.equalityConstraintsSSE = perf[2];
    after.cost = perf[3]; after.dynamicsViolationSSE = perf[4]; after.equalityConstraintsSSE = perf[5];
    before.merit = before.cost; after.merit = after.cost;
    performanceLog_.clear(); performanceLog_.push_back(before); performanceLog_.push_back(after);
  }

  ocs2::OptimalControlProblem ocp_;
  bmpc_handle* h_ = nullptr;
  int nx_ = 0, nu_ = 0, batch_ = 0, maxNodes_ = 0;
  std::vector<double> offsets_;
  ocs2::PrimalSolution primalSolution_;
  std::vector<ocs2::PerformanceIndex> performanceLog_;
  size_t numIterations_ = 0;
};

// Replaces ocs2::SqpMpc [UPSTREAM] at BipedalController.cpp:303.
class BmpcMpc final : public ocs2::MPC_BASE {
 public:
  BmpcMpc(ocs2::mpc::Settings mpcSettings, const ocs2::OptimalControlProblem& ocp, const BmpcSolver::Files& files, int replicas = 1, int device = 0)
      : ocs2::MPC_BASE(std::move(mpcSettings)), solverPtr_(new BmpcSolver(ocp, files, replicas, device)) {}
  ~BmpcMpc() override = default;
  BmpcSolver* getSolverPtr() override { return solverPtr_.get(); }
  const BmpcSolver* getSolverPtr() const override { return solverPtr_.get(); }

 protected:
  void calculateController(ocs2::scalar_t initTime, const ocs2::vector_t& initState, ocs2::scalar_t finalTime) override {
    if (settings().coldStart_) solverPtr_->reset();
    solverPtr_->run(initTime, initState, finalTime);
  }

 private:
  std::unique_ptr<BmpcSolver> solverPtr_;
};

}  // namespace bmpc

#endif  // BMPC_OCS2_ADAPTER_DISABLED
