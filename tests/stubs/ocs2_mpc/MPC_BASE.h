// test stub of ocs2_mpc/{MPC_Settings,MPC_BASE}.h
#pragma once
#include <utility>
#include <ocs2_core/Types.h>
#include <ocs2_oc/oc_solver/SolverBase.h>
namespace ocs2 {
namespace mpc {
struct Settings {
  scalar_t timeHorizon_ = 1.0;
  scalar_t solutionTimeWindow_ = -1;
  bool coldStart_ = false;
  bool debugPrint_ = false;
  scalar_t mpcDesiredFrequency_ = -1;
  scalar_t mrtDesiredFrequency_ = 100;
};
}  // namespace mpc
class MPC_BASE {
 public:
  explicit MPC_BASE(mpc::Settings mpcSettings) : mpcSettings_(std::move(mpcSettings)) {}
  virtual ~MPC_BASE() = default;
  virtual void reset() { initRun_ = true; getSolverPtr()->reset(); }
  virtual bool run(scalar_t currentTime, const vector_t& currentState) {
    const scalar_t finalTime = currentTime + mpcSettings_.timeHorizon_;
    calculateController(currentTime, currentState, finalTime);
    initRun_ = false;
    return true;
  }
  virtual SolverBase* getSolverPtr() = 0;
  virtual const SolverBase* getSolverPtr() const = 0;
  scalar_t getTimeHorizon() const { return mpcSettings_.timeHorizon_; }
  const mpc::Settings& settings() const { return mpcSettings_; }
 protected:
  virtual void calculateController(scalar_t initTime, const vector_t& initState, scalar_t finalTime) = 0;
  bool initRun_ = true;
 private:
  mpc::Settings mpcSettings_;
};
}  // namespace ocs2
