// test stub of ocs2_oc/oc_problem/OptimalControlProblem.h (opaque for the adapter: only copied and handed back)
#pragma once
namespace ocs2 {
struct OptimalControlProblem { int stubTag = 0; };
}  // namespace ocs2
