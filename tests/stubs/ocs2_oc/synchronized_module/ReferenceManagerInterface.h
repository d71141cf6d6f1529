// test stub of ocs2_oc/synchronized_module/ReferenceManagerInterface.h
#pragma once
#include <ocs2_core/Types.h>
#include <ocs2_core/reference/ModeSchedule.h>
#include <ocs2_core/reference/TargetTrajectories.h>
namespace ocs2 {
class ReferenceManagerInterface {
 public:
  virtual ~ReferenceManagerInterface() = default;
  virtual void preSolverRun(scalar_t initTime, scalar_t finalTime, const vector_t& initState) = 0;
  virtual const ModeSchedule& getModeSchedule() const = 0;
  virtual void setModeSchedule(const ModeSchedule& modeSchedule) = 0;
  virtual const TargetTrajectories& getTargetTrajectories() const = 0;
  virtual void setTargetTrajectories(const TargetTrajectories& targetTrajectories) = 0;
};
}  // namespace ocs2
