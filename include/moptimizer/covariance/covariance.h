// Mirrors include/moptimizer/covariance/covariance.h:10-13.  The reference aliases Eigen::MatrixX; Eigen is
// not a dependency here, so Matrix<Scalar> is a small column-major dense matrix exposing the members the
// reference's callers use (resize, setIdentity, operator(), data, rows, cols, *=).  With Eigen available,
// define MOPTIMIZER_USE_EIGEN to get the reference's alias instead.
#pragma once

#include <cstddef>
#include <memory>
#include <vector>

#include "moptimizer/types.h"

#if defined(MOPTIMIZER_USE_EIGEN) && __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
namespace moptimizer::covariance {
template <class Scalar>
using Matrix = Eigen::Matrix<Scalar, Eigen::Dynamic, Eigen::Dynamic>;
}
#else
namespace moptimizer::covariance {
template <class Scalar>
class Matrix {
 public:
  Matrix() = default;
  Matrix(int rows, int cols) { resize(rows, cols); }
  void resize(int rows, int cols) {
    rows_ = rows;
    cols_ = cols;
    v_.assign(static_cast<size_t>(rows) * cols, Scalar(0));
  }
  void setZero() { v_.assign(v_.size(), Scalar(0)); }
  void setIdentity() {
    setZero();
    for (int i = 0; i < rows_ && i < cols_; ++i) (*this)(i, i) = Scalar(1);
  }
  Scalar& operator()(int r, int c) { return v_[static_cast<size_t>(r) + static_cast<size_t>(c) * rows_]; }
  const Scalar& operator()(int r, int c) const { return v_[static_cast<size_t>(r) + static_cast<size_t>(c) * rows_]; }
  Matrix& operator*=(Scalar s) {
    for (auto& x : v_) x *= s;
    return *this;
  }
  Scalar* data() { return v_.data(); }
  const Scalar* data() const { return v_.data(); }
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  size_t size() const { return v_.size(); }

 private:
  int rows_ = 0, cols_ = 0;
  std::vector<Scalar> v_;
};
}  // namespace moptimizer::covariance
#endif

namespace moptimizer::covariance {
template <class Scalar>
using MatrixPtr = std::shared_ptr<Matrix<Scalar>>;
}  // namespace moptimizer::covariance
