// test stub of ocs2_core/control/{ControllerBase,LinearController}.h
#pragma once
#include <utility>
#include <ocs2_core/Types.h>
namespace ocs2 {
class ControllerBase {
 public:
  virtual ~ControllerBase() = default;
  virtual vector_t computeInput(scalar_t t, const vector_t& x) = 0;
  virtual ControllerBase* clone() const = 0;
};
class LinearController final : public ControllerBase {
 public:
  LinearController() = default;
  LinearController(scalar_array_t controllerTime, vector_array_t controllerBias, matrix_array_t controllerGain)
      : timeStamp_(std::move(controllerTime)), biasArray_(std::move(controllerBias)), gainArray_(std::move(controllerGain)) {}
  // u = uff(t) + K(t) x, linear interpolation in time (LinearInterpolation::timeSegment semantics)
  vector_t computeInput(scalar_t t, const vector_t& x) override {
    const size_t n = timeStamp_.size();
    size_t i = 0; scalar_t a = 1.0;
    if (n > 1) {
      size_t hi = 0; while (hi < n && timeStamp_[hi] < t) ++hi;
      if (hi == 0) { i = 0; a = 1.0; } else if (hi >= n) { i = n - 2; a = 0.0; } else { i = hi - 1; a = (timeStamp_[hi] - t) / (timeStamp_[hi] - timeStamp_[i]); }
    }
    const size_t j = n > 1 ? i + 1 : i;
    vector_t u(biasArray_[i].size());
    for (long r = 0; r < u.size(); ++r) {
      scalar_t s = a * biasArray_[i](r) + (1.0 - a) * biasArray_[j](r);
      for (long c = 0; c < x.size(); ++c) s += (a * gainArray_[i](r, c) + (1.0 - a) * gainArray_[j](r, c)) * x(c);
      u(r) = s;
    }
    return u;
  }
  LinearController* clone() const override { return new LinearController(*this); }
  scalar_array_t timeStamp_;
  vector_array_t biasArray_;
  matrix_array_t gainArray_;
};
}  // namespace ocs2
