// Drives include/bmpc_ocs2_adapter.hpp exactly as BipedalController drives SqpMpc (BipedalController.cpp:303-308, 332-351, 191-206):
//   mpc = BmpcMpc(...); solver->setReferenceManager(refManager); mpc.run(t, x); solver->primalSolution(tf); controller->computeInput(t, x)
// against the OCS2 stand-ins of tests/stubs/.  Prints the solution as text; tests/test_adapter.py compares it with the C ABI's Python mirror.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <bmpc_ocs2_adapter.hpp>

namespace {
class FixedReferenceManager final : public ocs2::ReferenceManagerInterface {
 public:
  void preSolverRun(ocs2::scalar_t, ocs2::scalar_t, const ocs2::vector_t&) override { ++preRuns; }
  const ocs2::ModeSchedule& getModeSchedule() const override { return ms; }
  void setModeSchedule(const ocs2::ModeSchedule& m) override { ms = m; }
  const ocs2::TargetTrajectories& getTargetTrajectories() const override { return tt; }
  void setTargetTrajectories(const ocs2::TargetTrajectories& t) override { tt = t; }
  ocs2::ModeSchedule ms; ocs2::TargetTrajectories tt; int preRuns = 0;
};
}  // namespace

// usage: adapter_main <model file> <input file>; input: nx, x0[nx], n_events, events, modes[n_events + 1], n_target, (time, state[nx]) x n_target
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[2]);
  int nx; in >> nx;
  ocs2::vector_t x0(nx); for (int i = 0; i < nx; ++i) in >> x0(i);
  int ne; in >> ne;
  std::vector<double> ev(ne); for (auto& e : ev) in >> e;
  std::vector<size_t> modes(ne + 1); for (auto& m : modes) in >> m;
  int nt; in >> nt;
  ocs2::TargetTrajectories tt;
  for (int k = 0; k < nt; ++k) { double t; in >> t; ocs2::vector_t s(nx); for (int i = 0; i < nx; ++i) in >> s(i); tt.timeTrajectory.push_back(t); tt.stateTrajectory.push_back(s); }
  try {
    ocs2::mpc::Settings mpcSettings; mpcSettings.timeHorizon_ = 1.0;
    bmpc::BmpcSolver::Files files; files.model = argv[1];
    bmpc::BmpcMpc mpc(mpcSettings, ocs2::OptimalControlProblem{}, files, /*replicas=*/2);
    auto ref = std::make_shared<FixedReferenceManager>();
    ref->setModeSchedule(ocs2::ModeSchedule(ev, modes)); ref->setTargetTrajectories(tt);
    mpc.getSolverPtr()->setReferenceManager(ref);
    for (int tick = 0; tick < 2; ++tick) mpc.run(0.0, x0);          // cold tick, warm tick (MPC thread, BipedalController.cpp:339)
    const ocs2::PrimalSolution sol = mpc.getSolverPtr()->primalSolution(mpc.getTimeHorizon());
    const size_t n = sol.timeTrajectory_.size();
    std::printf("n %zu preRuns %d iterations %zu finalTime %.17g events", n, ref->preRuns, mpc.getSolverPtr()->getNumIterations(), mpc.getSolverPtr()->getFinalTime());
    for (size_t i : sol.postEventIndices_) std::printf(" %zu", i);
    std::printf("\n");
    const ocs2::PerformanceIndex& pi = mpc.getSolverPtr()->getPerformanceIndeces();
    std::printf("perf %.17g %.17g %.17g\n", pi.cost, pi.dynamicsViolationSSE, pi.equalityConstraintsSSE);
    for (size_t k = 0; k < n; ++k) {
      std::printf("t %.17g x", sol.timeTrajectory_[k]);
      for (int i = 0; i < nx; ++i) std::printf(" %.17g", sol.stateTrajectory_[k](i));
      std::printf(" u");
      for (long i = 0; i < sol.inputTrajectory_[k].size(); ++i) std::printf(" %.17g", sol.inputTrajectory_[k](i));
      std::printf("\n");
    }
    // MRT side: evaluatePolicy = state interpolation + LinearController (BipedalController.cpp:200)
    ocs2::vector_t xq(nx); for (int i = 0; i < nx; ++i) xq(i) = x0(i) + 0.01;
    const ocs2::vector_t u = sol.controllerPtr_->computeInput(0.013, xq);
    std::printf("uq");
    for (long i = 0; i < u.size(); ++i) std::printf(" %.17g", u(i));
    std::printf("\n");
    mpc.reset();
    std::printf("reset ok %zu\n", mpc.getSolverPtr()->getNumIterations());
  } catch (const std::exception& e) { std::fprintf(stderr, "exception: %s\n", e.what()); return 1; }
  return 0;
}
