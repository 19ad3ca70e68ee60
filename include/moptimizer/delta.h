// Mirrors include/moptimizer/delta.h:11-16 with a raw-pointer signature (no Eigen).
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
