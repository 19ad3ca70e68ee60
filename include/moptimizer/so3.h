// Mirrors the on-path part of include/moptimizer/so3.h + src/so3.cpp:7-19,43-57 with raw-array signatures
// (row-major), no Eigen.  The device copies used by the kernels live in csrc/mopt_setup.cuh.
#pragma once

#include <cmath>
#include <limits>

namespace so3 {

/// Rodrigues, guard `norm > 10 eps` (src/so3.cpp:43-57).  R is 3x3 row-major.
template <typename Scalar>
inline void Exp(const Scalar* delta, Scalar* R) {
  const Scalar n = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? Scalar(1) : Scalar(0);
  if (n > Scalar(10.0) * std::numeric_limits<Scalar>::epsilon()) {
    const Scalar a[3] = {delta[0] / n, delta[1] / n, delta[2] / n};
    const Scalar K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    const Scalar s = std::sin(n), c = std::cos(n);
    for (int r = 0; r < 3; ++r)
      for (int col = 0; col < 3; ++col) {
        Scalar kk = 0;
        for (int k = 0; k < 3; ++k) kk += K[r * 3 + k] * K[k * 3 + col];
        R[r * 3 + col] += s * K[r * 3 + col] + (Scalar(1.0) - c) * kk;
      }
  }
}

/// x = [t(3), omega(3)] -> 4x4 homogeneous transform, row-major (src/so3.cpp:7-19).
template <typename Scalar>
inline void convert6DOFParameterToMatrix(const Scalar* x, Scalar* T16) {
  Scalar R[9];
  Exp<Scalar>(x + 3, R);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T16[r * 4 + c] = R[r * 3 + c];
    T16[r * 4 + 3] = x[r];
  }
  T16[12] = T16[13] = T16[14] = Scalar(0);
  T16[15] = Scalar(1);
}

}  // namespace so3
