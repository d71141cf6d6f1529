// test stub of ocs2_oc/oc_solver/SolverBase.h: the virtual interface and the run() = preRun -> runImpl -> postRun wrapper
#pragma once
#include <memory>
#include <stdexcept>
#include <vector>
#include <ocs2_core/Types.h>
#include <ocs2_core/control/LinearController.h>
#include <ocs2_oc/oc_data/PerformanceIndex.h>
#include <ocs2_oc/oc_data/PrimalSolution.h>
#include <ocs2_oc/oc_problem/OptimalControlProblem.h>
#include <ocs2_oc/synchronized_module/ReferenceManagerInterface.h>
namespace ocs2 {
class SolverBase {
 public:
  SolverBase() = default;
  virtual ~SolverBase() = default;
  virtual void reset() = 0;
  void run(scalar_t initTime, const vector_t& initState, scalar_t finalTime) {
    if (referenceManagerPtr_) referenceManagerPtr_->preSolverRun(initTime, finalTime, initState);
    runImpl(initTime, initState, finalTime);
  }
  void run(scalar_t initTime, const vector_t& initState, scalar_t finalTime, const ControllerBase* externalControllerPtr) {
    if (referenceManagerPtr_) referenceManagerPtr_->preSolverRun(initTime, finalTime, initState);
    runImpl(initTime, initState, finalTime, externalControllerPtr);
  }
  void setReferenceManager(std::shared_ptr<ReferenceManagerInterface> referenceManagerPtr) {
    if (referenceManagerPtr == nullptr) throw std::runtime_error("[SolverBase] ReferenceManager pointer cannot be a nullptr!");
    referenceManagerPtr_ = std::move(referenceManagerPtr);
  }
  const ReferenceManagerInterface& getReferenceManager() const { return *referenceManagerPtr_; }
  ReferenceManagerInterface& getReferenceManager() { return *referenceManagerPtr_; }
  virtual const PerformanceIndex& getPerformanceIndeces() const = 0;
  virtual size_t getNumIterations() const = 0;
  virtual const OptimalControlProblem& getOptimalControlProblem() const = 0;
  virtual const std::vector<PerformanceIndex>& getIterationsLog() const = 0;
  virtual scalar_t getFinalTime() const = 0;
  virtual void getPrimalSolution(scalar_t finalTime, PrimalSolution* primalSolutionPtr) const = 0;
  PrimalSolution primalSolution(scalar_t finalTime) const { PrimalSolution p; getPrimalSolution(finalTime, &p); return p; }
  virtual ScalarFunctionQuadraticApproximation getValueFunction(scalar_t time, const vector_t& state) const = 0;
  virtual ScalarFunctionQuadraticApproximation getHamiltonian(scalar_t time, const vector_t& state, const vector_t& input) = 0;
  virtual vector_t getStateInputEqualityConstraintLagrangian(scalar_t time, const vector_t& state) const = 0;
 private:
  virtual void runImpl(scalar_t initTime, const vector_t& initState, scalar_t finalTime) = 0;
  virtual void runImpl(scalar_t initTime, const vector_t& initState, scalar_t finalTime, const ControllerBase* externalControllerPtr) = 0;
  std::shared_ptr<ReferenceManagerInterface> referenceManagerPtr_;
};
}  // namespace ocs2
