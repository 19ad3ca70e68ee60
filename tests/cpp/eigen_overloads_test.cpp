// The reference's Eigen-typed call forms of so3::* and isDeltaSmall (include/moptimizer/so3.h:8-41, delta.h:11-16),
// as its tests and models use them (tst/point2point.cpp:33, tst/camera_calibration.cpp:33, tst/state_model.cpp:17-46,
// tst/manifold.cpp:41-45, src/levenberg_marquadt_dyn.cpp:98), compiled against the mirror headers with
// MOPTIMIZER_USE_EIGEN and checked against the raw-array functions they forward to.  Host only.
// Build: g++ -std=c++17 -DMOPTIMIZER_USE_EIGEN -I include -I tests/cpp/eigen_stub tests/cpp/eigen_overloads_test.cpp
#include <cmath>
#include <cstdio>

#include "moptimizer/delta.h"
#include "moptimizer/so3.h"

static int failures = 0;
#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) {                                                       \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);        \
      ++failures;                                                        \
    }                                                                    \
  } while (0)

template <typename Scalar>
static void run() {
  using M3 = Eigen::Matrix<Scalar, 3, 3>;
  using M4 = Eigen::Matrix<Scalar, 4, 4>;
  using V3 = Eigen::Matrix<Scalar, 3, 1>;
  const Scalar x[6] = {Scalar(10.5), Scalar(10.2), Scalar(0.1), Scalar(0.39), Scalar(0.315), Scalar(0.55)};
  // tst/point2point.cpp:33 / tst/camera_calibration.cpp:33
  M4 transform_;
  so3::convert6DOFParameterToMatrix(x, transform_);
  Scalar T[16];
  so3::convert6DOFParameterToMatrix<Scalar>(x, T);
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) CHECK(transform_(r, c) == T[r * 4 + c]);
  M4 t3;
  so3::convert3DOFParameterToMatrix(x + 3, t3);
  M3 r3;
  so3::convert3DOFParameterToMatrix3(x + 3, r3);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) CHECK(t3(r, c) == T[r * 4 + c] && r3(r, c) == T[r * 4 + c]);
  CHECK(t3(0, 3) == Scalar(0) && t3(3, 3) == Scalar(1));
  // src/so3.cpp:43-57 through Ref<>, as tst/state_model.cpp:28-31 calls it
  V3 w;
  w[0] = x[3]; w[1] = x[4]; w[2] = x[5];
  M3 R;
  so3::Exp<Scalar>(w, R);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) CHECK(R(r, c) == T[r * 4 + c]);
  // the value-returning forms (tst/manifold.cpp:41-45) agree with the Rodrigues form to rounding
  const M3 Ra = so3::Exp<Scalar>(w);
  const M3 Rh = so3::Exp<Scalar>(w, Scalar(0.5));
  V3 half;
  for (int i = 0; i < 3; ++i) half[i] = w[i] * Scalar(0.5);
  M3 Rhalf;
  so3::Exp<Scalar>(half, Rhalf);
  const Scalar tol = sizeof(Scalar) == 4 ? Scalar(1e-6) : Scalar(1e-14);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) CHECK(std::fabs(Ra(r, c) - R(r, c)) <= tol && std::fabs(Rh(r, c) - Rhalf(r, c)) <= tol);
  // Log(Exp(w)) = w (tst/state_model.cpp:33-46)
  V3 back;
  so3::Log<Scalar>(R, back);
  for (int i = 0; i < 3; ++i) CHECK(std::fabs(back[i] - w[i]) <= Scalar(20) * tol);
  // the three Jacobians forward to the raw-array forms
  M3 Jr, Jl, Jir;
  so3::rightJacobian<Scalar>(w, Jr);
  so3::leftJacobian<Scalar>(w, Jl);
  so3::inverseRightJacobian<Scalar>(w, Jir);
  Scalar jr[9], jl[9], jir[9];
  so3::rightJacobian<Scalar>(x + 3, jr);
  so3::leftJacobian<Scalar>(x + 3, jl);
  so3::inverseRightJacobian<Scalar>(x + 3, jir);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) CHECK(Jr(r, c) == jr[r * 3 + c] && Jl(r, c) == jl[r * 3 + c] && Jir(r, c) == jir[r * 3 + c]);
  // delta.h:11-16 as src/levenberg_marquadt_dyn.cpp:98 calls it
  Eigen::Matrix<Scalar, 6, 1> d;
  for (int i = 0; i < 6; ++i) d[i] = Scalar(1e-9) * Scalar(i - 3);
  CHECK(moptimizer::isDeltaSmall(d));
  d[4] = Scalar(1e-2);
  CHECK(!moptimizer::isDeltaSmall(d));
  const double skew[9] = {SKEW_SYMMETRIC_FROM(w)};
  CHECK(skew[1] == -double(w[2]) && skew[5] == -double(w[0]) && skew[6] == -double(w[1]));
}

int main() {
  run<double>();
  run<float>();
  if (failures) {
    std::printf("%d failure(s)\n", failures);
    return 1;
  }
  std::printf("eigen overloads ok\n");
  return 0;
}
