// test stub of ocs2_core/Types.h: scalar_t = double; vector_t / matrix_t with the subset of the Eigen dynamic-size API the adapter uses
#pragma once
#include <cstddef>
#include <vector>
namespace ocs2 {
using scalar_t = double;
using size_array_t = std::vector<size_t>;
using scalar_array_t = std::vector<scalar_t>;
class vector_t {
 public:
  vector_t() = default;
  explicit vector_t(long n) : d_(static_cast<size_t>(n), 0.0) {}
  long size() const { return static_cast<long>(d_.size()); }
  void resize(long n) { d_.assign(static_cast<size_t>(n), 0.0); }
  scalar_t& operator()(long i) { return d_[static_cast<size_t>(i)]; }
  const scalar_t& operator()(long i) const { return d_[static_cast<size_t>(i)]; }
  scalar_t* data() { return d_.data(); }
  const scalar_t* data() const { return d_.data(); }
 private:
  std::vector<scalar_t> d_;
};
class matrix_t {   // column major, as Eigen's default
 public:
  matrix_t() = default;
  matrix_t(long r, long c) : r_(r), c_(c), d_(static_cast<size_t>(r * c), 0.0) {}
  long rows() const { return r_; }
  long cols() const { return c_; }
  void resize(long r, long c) { r_ = r; c_ = c; d_.assign(static_cast<size_t>(r * c), 0.0); }
  scalar_t& operator()(long i, long j) { return d_[static_cast<size_t>(j * r_ + i)]; }
  const scalar_t& operator()(long i, long j) const { return d_[static_cast<size_t>(j * r_ + i)]; }
 private:
  long r_ = 0, c_ = 0;
  std::vector<scalar_t> d_;
};
using vector_array_t = std::vector<vector_t>;
using matrix_array_t = std::vector<matrix_t>;
struct ScalarFunctionQuadraticApproximation { matrix_t dfdxx, dfdux, dfduu; vector_t dfdx, dfdu; scalar_t f = 0.0; };
}  // namespace ocs2
