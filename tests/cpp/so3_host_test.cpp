// Host-only checks of the so3 mirror (include/moptimizer/so3.h) against closed forms and the reference's own use
// (tst/state_model.cpp:17-46: Exp / Log round trip; tst/manifold.cpp:41-45: R <- R Exp(delta)).  No GPU needed.
#include <cmath>
#include <cstdio>

#include "moptimizer/so3.h"

static int failures = 0;
#define CHECK_NEAR(a, b, tol)                                                                  \
  do {                                                                                         \
    if (!(std::fabs((a) - (b)) <= (tol))) {                                                    \
      std::printf("FAIL %s:%d  %s = %.17g vs %s = %.17g\n", __FILE__, __LINE__, #a, double(a), #b, double(b)); \
      ++failures;                                                                              \
    }                                                                                          \
  } while (0)

static void matmul(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      C[r * 3 + c] = 0;
      for (int k = 0; k < 3; ++k) C[r * 3 + c] += A[r * 3 + k] * B[k * 3 + c];
    }
}

int main() {
  // rotation about z by 0.3: closed form
  double w[3] = {0, 0, 0.3}, R[9];
  so3::Exp<double>(w, R);
  CHECK_NEAR(R[0], std::cos(0.3), 1e-15); CHECK_NEAR(R[1], -std::sin(0.3), 1e-15); CHECK_NEAR(R[8], 1.0, 1e-15);
  // orthonormality + Log(Exp(w)) = w for a generic vector (tst/state_model.cpp Plus/Minus round trip)
  double v[3] = {0.1, 0.2, 0.3}, Rv[9], RtR[9], Rt[9], back[3];
  so3::Exp<double>(v, Rv);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Rt[r * 3 + c] = Rv[c * 3 + r];
  matmul(Rt, Rv, RtR);
  for (int i = 0; i < 9; ++i) CHECK_NEAR(RtR[i], (i % 4 == 0) ? 1.0 : 0.0, 1e-15);
  so3::Log<double>(Rv, back);
  for (int i = 0; i < 3; ++i) CHECK_NEAR(back[i], v[i], 1e-14);
  // the three Exp flavours agree; the dt form scales the angle
  double R2[9], R3[9], half[3] = {0.05, 0.1, 0.15};
  so3::ExpAng<double>(v, R2);
  so3::Exp<double>(half, 2.0, R3);
  for (int i = 0; i < 9; ++i) { CHECK_NEAR(R2[i], Rv[i], 1e-15); CHECK_NEAR(R3[i], Rv[i], 1e-15); }
  // guards: identity below the thresholds, first-order Log branch
  double tiny[3] = {1e-16, 0, 0}, Rtiny[9];
  so3::Exp<double>(tiny, Rtiny);
  for (int i = 0; i < 9; ++i) CHECK_NEAR(Rtiny[i], (i % 4 == 0) ? 1.0 : 0.0, 0.0);
  double small[3] = {2e-4, -1e-4, 3e-4}, Rs[9], ls[3];
  so3::Exp<double>(small, Rs);
  so3::Log<double>(Rs, ls);
  for (int i = 0; i < 3; ++i) CHECK_NEAR(ls[i], small[i], 1e-10);
  // convert6DOF / convert3DOF layouts
  double x6[6] = {1, 2, 3, 0.1, 0.2, 0.3}, T[16], T3[16], R33[9];
  so3::convert6DOFParameterToMatrix<double>(x6, T);
  so3::convert3DOFParameterToMatrix<double>(v, T3);
  so3::convert3DOFParameterToMatrix3<double>(v, R33);
  for (int r = 0; r < 3; ++r) {
    CHECK_NEAR(T[r * 4 + 3], x6[r], 0.0);
    CHECK_NEAR(T3[r * 4 + 3], 0.0, 0.0);
    for (int c = 0; c < 3; ++c) { CHECK_NEAR(T[r * 4 + c], Rv[r * 3 + c], 0.0); CHECK_NEAR(T3[r * 4 + c], Rv[r * 3 + c], 0.0); CHECK_NEAR(R33[r * 3 + c], Rv[r * 3 + c], 0.0); }
  }
  CHECK_NEAR(T[15], 1.0, 0.0); CHECK_NEAR(T[12] + T[13] + T[14], 0.0, 0.0);
  // Jacobians as the reference defines them: J_r = I - a [r]x, J_l = I + a [r]x with a = (1 - cos t) / t^2, so
  // J_l + J_r = 2 I; the inverse right Jacobian inverts the EXACT right Jacobian I - a [r]x + b [r]x^2
  double Jr[9], Jl[9], Jinv[9];
  so3::rightJacobian<double>(v, Jr);
  so3::leftJacobian<double>(v, Jl);
  for (int i = 0; i < 9; ++i) CHECK_NEAR(Jr[i] + Jl[i], (i % 4 == 0) ? 2.0 : 0.0, 1e-15);
  const double t = std::sqrt(0.14), a = (1 - std::cos(t)) / (t * t), b = (t - std::sin(t)) / (t * t * t);
  CHECK_NEAR(Jr[1], a * v[2], 1e-15);  // -a * K(0,1) = a * r_z
  const double K[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
  double KK[9], Jexact[9], prod[9];
  matmul(K, K, KK);
  for (int i = 0; i < 9; ++i) Jexact[i] = ((i % 4 == 0) ? 1.0 : 0.0) - a * K[i] + b * KK[i];
  so3::inverseRightJacobian<double>(v, Jinv);
  matmul(Jinv, Jexact, prod);
  for (int i = 0; i < 9; ++i) CHECK_NEAR(prod[i], (i % 4 == 0) ? 1.0 : 0.0, 1e-12);
  double z[3] = {1e-3, 0, 0}, Jz[9];
  so3::inverseRightJacobian<double>(z, Jz);
  for (int i = 0; i < 9; ++i) CHECK_NEAR(Jz[i], (i % 4 == 0) ? 1.0 : 0.0, 0.0);  // |r|^2 < 1e-5 -> identity
  // float instantiation
  float vf[3] = {0.1f, 0.2f, 0.3f}, Rf[9];
  so3::Exp<float>(vf, Rf);
  for (int i = 0; i < 9; ++i) CHECK_NEAR(Rf[i], Rv[i], 2e-7);
  std::printf("%s (%d failures)\n", failures ? "so3 host test FAILED" : "so3 host test ok", failures);
  return failures ? 1 : 0;
}
