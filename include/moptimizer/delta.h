// Mirrors include/moptimizer/delta.h:11-16.  The test itself takes a raw pointer (no dependency); with
// MOPTIMIZER_USE_EIGEN defined and <Eigen/Dense> on the include path the reference's own signature
// `isDeltaSmall(const Eigen::Matrix<Scalar, Dim, 1>&)` is declared as well and forwards to it.
#pragma once

#include <cmath>
#include <limits>

namespace moptimizer {

/// max |delta_i| < sqrt(eps)
template <class Scalar>
inline bool isDeltaSmall(const Scalar* delta, int n) {
  Scalar m = 0;
  for (int i = 0; i < n; ++i) m = std::fabs(delta[i]) > m ? Scalar(std::fabs(delta[i])) : m;
  return m < std::sqrt(std::numeric_limits<Scalar>::epsilon());
}

}  // namespace moptimizer

#if defined(MOPTIMIZER_USE_EIGEN) && __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
namespace moptimizer {
/// include/moptimizer/delta.h:11-16
template <class Scalar, int Dim>
inline bool isDeltaSmall(const Eigen::Matrix<Scalar, Dim, 1>& delta) {
  return isDeltaSmall<Scalar>(delta.data(), int(delta.size()));
}
}  // namespace moptimizer
#endif
