// test stub of ocs2_oc/oc_data/PrimalSolution.h
#pragma once
#include <memory>
#include <ocs2_core/Types.h>
#include <ocs2_core/control/LinearController.h>
#include <ocs2_core/reference/ModeSchedule.h>
namespace ocs2 {
struct PrimalSolution {
  PrimalSolution() = default;
  PrimalSolution(const PrimalSolution& o)
      : timeTrajectory_(o.timeTrajectory_), postEventIndices_(o.postEventIndices_), stateTrajectory_(o.stateTrajectory_), inputTrajectory_(o.inputTrajectory_),
        modeSchedule_(o.modeSchedule_), controllerPtr_(o.controllerPtr_ ? o.controllerPtr_->clone() : nullptr) {}
  PrimalSolution(PrimalSolution&&) = default;
  PrimalSolution& operator=(PrimalSolution&&) = default;
  PrimalSolution& operator=(const PrimalSolution& o) { PrimalSolution t(o); *this = std::move(t); return *this; }
  scalar_array_t timeTrajectory_;
  size_array_t postEventIndices_;
  vector_array_t stateTrajectory_;
  vector_array_t inputTrajectory_;
  ModeSchedule modeSchedule_;
  std::unique_ptr<ControllerBase> controllerPtr_;
};
}  // namespace ocs2
