// Mirrors include/moptimizer/so3.h + src/so3.cpp.  The arithmetic lives in raw-array functions (3x3 / 4x4 row-major,
// no dependency); with MOPTIMIZER_USE_EIGEN defined and <Eigen/Dense> on the include path the reference's own
// Eigen-typed signatures (include/moptimizer/so3.h:8-41) are declared as well and forward to them, so that reference
// code such as tst/point2point.cpp:33 `so3::convert6DOFParameterToMatrix(x, transform_)` compiles unchanged.
// convert6DOFParameterToMatrix and Exp are on the hot path (the device copies the kernels use live in
// csrc/mopt_setup.cuh); the rest of src/so3.cpp:21-155 is the kept API surface around the manifold update, restated
// with the reference's own thresholds and formulas — including its first-order right/left Jacobians.
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

/// x = omega(3) -> 4x4 homogeneous transform with zero translation (src/so3.cpp:21-31).
template <typename Scalar>
inline void convert3DOFParameterToMatrix(const Scalar* x, Scalar* T16) {
  Scalar R[9];
  Exp<Scalar>(x, R);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T16[r * 4 + c] = R[r * 3 + c];
    T16[r * 4 + 3] = Scalar(0);
  }
  T16[12] = T16[13] = T16[14] = Scalar(0);
  T16[15] = Scalar(1);
}

/// x = omega(3) -> 3x3 rotation (src/so3.cpp:33-40).
template <typename Scalar>
inline void convert3DOFParameterToMatrix3(const Scalar* x, Scalar* R9) {
  Exp<Scalar>(x, R9);
}

/// `Matrix3 Exp(const Vector3& ang_vel, const Scalar& dt)` (src/so3.cpp:76-94): rotation by |ang_vel| dt about
/// ang_vel, identity below |ang_vel| = 1e-7.  `Matrix3 Exp(const Vector3& ang)` (:60-74) is the dt = 1 case.
template <typename Scalar>
inline void Exp(const Scalar* ang_vel, const Scalar& dt, Scalar* R) {
  const Scalar n = std::sqrt(ang_vel[0] * ang_vel[0] + ang_vel[1] * ang_vel[1] + ang_vel[2] * ang_vel[2]);
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? Scalar(1) : Scalar(0);
  if (n > Scalar(0.0000001)) {
    const Scalar a[3] = {ang_vel[0] / n, ang_vel[1] / n, ang_vel[2] / n};
    const Scalar K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    const Scalar ang = n * dt;
    const Scalar s = std::sin(ang), c = std::cos(ang);
    for (int r = 0; r < 3; ++r)
      for (int col = 0; col < 3; ++col) {
        Scalar kk = 0;
        for (int k = 0; k < 3; ++k) kk += K[r * 3 + k] * K[k * 3 + col];
        R[r * 3 + col] += s * K[r * 3 + col] + (Scalar(1.0) - c) * kk;
      }
  }
}
template <typename Scalar>
inline void ExpAng(const Scalar* ang, Scalar* R) {
  Exp<Scalar>(ang, Scalar(1), R);
}

/// delta = Log(R) (src/so3.cpp:96-105): theta = 0 when trace(R) > 3 - 1e-6, first-order branch below theta = 1e-3.
template <typename Scalar>
inline void Log(const Scalar* R, Scalar* delta) {
  const Scalar tr = R[0] + R[4] + R[8];
  const Scalar theta = (tr > Scalar(3.0 - 1e-6)) ? Scalar(0) : std::acos(Scalar(0.5) * (tr - Scalar(1)));
  const Scalar K[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  const Scalar f = (std::abs(theta) < Scalar(0.001)) ? Scalar(0.5) : Scalar(0.5) * theta / std::sin(theta);
  for (int i = 0; i < 3; ++i) delta[i] = f * K[i];
}

namespace detail {
template <typename Scalar>
inline void skew(const Scalar* v, Scalar* K) {
  const Scalar k[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
  for (int i = 0; i < 9; ++i) K[i] = k[i];
}
}  // namespace detail

/// src/so3.cpp:107-121: I + [r]x / 2 + (1/|r|^2 - (1 + cos|r|) / (2 |r| sin|r|)) [r]x^2, identity below |r|^2 = 1e-5.
template <typename Scalar>
inline void inverseRightJacobian(const Scalar* r, Scalar* J) {
  const double theta_sq = double(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  for (int i = 0; i < 9; ++i) J[i] = (i % 4 == 0) ? Scalar(1) : Scalar(0);
  if (theta_sq < 1e-5) return;
  Scalar K[9];
  detail::skew(r, K);
  const Scalar n = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  const Scalar factor = Scalar(1) / (n * n) - (Scalar(1) + std::cos(n)) / (Scalar(2) * n * std::sin(n));
  for (int row = 0; row < 3; ++row)
    for (int col = 0; col < 3; ++col) {
      Scalar kk = 0;
      for (int k = 0; k < 3; ++k) kk += K[row * 3 + k] * K[k * 3 + col];
      J[row * 3 + col] += Scalar(0.5) * K[row * 3 + col] + factor * kk;
    }
}

/// src/so3.cpp:123-139: the reference's right Jacobian, I - (1 - cos|r|) / |r|^2 [r]x (no second-order term).
template <typename Scalar>
inline void rightJacobian(const Scalar* r, Scalar* J) {
  const double theta_sq = double(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  for (int i = 0; i < 9; ++i) J[i] = (i % 4 == 0) ? Scalar(1) : Scalar(0);
  if (theta_sq < 1e-5) return;
  Scalar K[9];
  detail::skew(r, K);
  const Scalar factor = Scalar((1.0 - std::cos(std::sqrt(theta_sq))) / theta_sq);
  for (int i = 0; i < 9; ++i) J[i] -= factor * K[i];
}

/// src/so3.cpp:141-155: the reference's left Jacobian, I + (1 - cos|r|) / |r|^2 [r]x (no second-order term; the
/// device path's exact point2point Jacobian uses the full closed form, csrc/mopt_setup.cuh so3_left_jacobian_dev).
template <typename Scalar>
inline void leftJacobian(const Scalar* r, Scalar* J) {
  const double theta_sq = double(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  for (int i = 0; i < 9; ++i) J[i] = (i % 4 == 0) ? Scalar(1) : Scalar(0);
  if (theta_sq < 1e-5) return;
  Scalar K[9];
  detail::skew(r, K);
  const Scalar factor = Scalar((1.0 - std::cos(std::sqrt(theta_sq))) / theta_sq);
  for (int i = 0; i < 9; ++i) J[i] += factor * K[i];
}

}  // namespace so3

#if defined(MOPTIMIZER_USE_EIGEN) && __has_include(<Eigen/Dense>)
#include <Eigen/Dense>

#ifndef SKEW_SYMMETRIC_FROM  // include/moptimizer/so3.h:4
#define SKEW_SYMMETRIC_FROM(v) 0.0, -v[2], v[1], v[2], 0.0, -v[0], -v[1], v[0], 0.0
#endif

namespace so3 {
namespace detail {
template <typename Scalar, typename M>
inline void store3(const Scalar* R9, M&& out) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out(r, c) = R9[r * 3 + c];
}
template <typename Scalar, typename M>
inline void store4(const Scalar* T16, M&& out) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out(r, c) = T16[r * 4 + c];
}
}  // namespace detail

/// include/moptimizer/so3.h:8-9 (src/so3.cpp:7-19)
template <typename Scalar>
inline void convert6DOFParameterToMatrix(const Scalar* x, Eigen::Matrix<Scalar, 4, 4>& transform_matrix_) {
  Scalar T[16];
  convert6DOFParameterToMatrix<Scalar>(x, T);
  detail::store4(T, transform_matrix_);
}
/// :11-12 (src/so3.cpp:21-31)
template <typename Scalar>
inline void convert3DOFParameterToMatrix(const Scalar* x, Eigen::Matrix<Scalar, 4, 4>& transform_matrix_) {
  Scalar T[16];
  convert3DOFParameterToMatrix<Scalar>(x, T);
  detail::store4(T, transform_matrix_);
}
/// :14-15 (src/so3.cpp:33-40)
template <typename Scalar>
inline void convert3DOFParameterToMatrix3(const Scalar* x, Eigen::Matrix<Scalar, 3, 3>& transform_matrix_) {
  Scalar R[9];
  convert3DOFParameterToMatrix3<Scalar>(x, R);
  detail::store3(R, transform_matrix_);
}
/// :18-20 (src/so3.cpp:43-57): R = exp(delta)
template <typename Scalar>
inline void Exp(const Eigen::Ref<const Eigen::Matrix<Scalar, 3, 1>>& delta, Eigen::Ref<Eigen::Matrix<Scalar, 3, 3>> R) {
  const Scalar d[3] = {delta[0], delta[1], delta[2]};
  Scalar R9[9];
  Exp<Scalar>(d, R9);
  detail::store3(R9, R);
}
/// :22-23 (src/so3.cpp:60-74)
template <typename Scalar>
inline Eigen::Matrix<Scalar, 3, 3> Exp(const Eigen::Matrix<Scalar, 3, 1>& ang) {
  const Scalar d[3] = {ang[0], ang[1], ang[2]};
  Scalar R9[9];
  ExpAng<Scalar>(d, R9);
  Eigen::Matrix<Scalar, 3, 3> R;
  detail::store3(R9, R);
  return R;
}
/// :25-26 (src/so3.cpp:76-94)
template <typename Scalar>
inline Eigen::Matrix<Scalar, 3, 3> Exp(const Eigen::Matrix<Scalar, 3, 1>& ang_vel, const Scalar& dt) {
  const Scalar d[3] = {ang_vel[0], ang_vel[1], ang_vel[2]};
  Scalar R9[9];
  Exp<Scalar>(d, dt, R9);
  Eigen::Matrix<Scalar, 3, 3> R;
  detail::store3(R9, R);
  return R;
}
/// :29-30 (src/so3.cpp:96-105): delta = Log(R)
template <typename Scalar>
inline void Log(const Eigen::Ref<Eigen::Matrix<Scalar, 3, 3>>& R, Eigen::Matrix<Scalar, 3, 1>& delta) {
  Scalar R9[9], d[3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R9[r * 3 + c] = R(r, c);
  Log<Scalar>(R9, d);
  for (int i = 0; i < 3; ++i) delta[i] = d[i];
}
/// :32-34 (src/so3.cpp:107-121)
template <typename Scalar>
inline void inverseRightJacobian(const Eigen::Matrix<Scalar, 3, 1>& r, Eigen::Ref<Eigen::Matrix<Scalar, 3, 3>> inv_jacobian) {
  const Scalar v[3] = {r[0], r[1], r[2]};
  Scalar J[9];
  inverseRightJacobian<Scalar>(v, J);
  detail::store3(J, inv_jacobian);
}
/// :36-38 (src/so3.cpp:123-139)
template <typename Scalar>
inline void rightJacobian(const Eigen::Ref<const Eigen::Matrix<Scalar, 3, 1>>& r, Eigen::Ref<Eigen::Matrix<Scalar, 3, 3>> jacobian) {
  const Scalar v[3] = {r[0], r[1], r[2]};
  Scalar J[9];
  rightJacobian<Scalar>(v, J);
  detail::store3(J, jacobian);
}
/// :40-42 (src/so3.cpp:141-155)
template <typename Scalar>
inline void leftJacobian(const Eigen::Ref<const Eigen::Matrix<Scalar, 3, 1>>& r, Eigen::Ref<Eigen::Matrix<Scalar, 3, 3>> left_jacobian) {
  const Scalar v[3] = {r[0], r[1], r[2]};
  Scalar J[9];
  leftJacobian<Scalar>(v, J);
  detail::store3(J, left_jacobian);
}
}  // namespace so3
#endif  // MOPTIMIZER_USE_EIGEN
