// Host-only check of the AFFINE_FD model hooks (csrc/mopt_models.cuh: affine / finish / finish_diff / tail_partials)
// that the fp32 finite-difference pass kernels use for the two camera models: compiled with nvcc, run on the CPU.
//   * finish_diff(ua, du, H) must equal the literal difference quotient (f(ua + H du) - f(ua)) / H, which is what
//     the reference forms per residual (linearization.h:97-111); the literal value is taken in long double so that
//     its own cancellation does not pollute the comparison;
//   * tail_partials must equal the difference quotient in each second-stage parameter (the residual is affine in
//     each of them alone, so the quotient does not depend on the step);
//   * the float instantiation must agree with the double one to a few float ulp of the column's scale — the point
//     of the mode: the literal float quotient is off by eps_f32 |pixel| / H instead.
#include <cmath>
#include <cstdio>
#include <random>

#include "mopt_models.cuh"

using namespace mopt;

static int failures = 0;
#define CHECK(cond, ...)                          \
  do {                                            \
    if (!(cond)) {                                \
      std::printf("FAIL %s:%d  ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);                   \
      std::printf("\n");                          \
      ++failures;                                 \
    }                                             \
  } while (0)

typedef long double ld;

// literal residuals in long double, same formulas as the device models
static void pinhole_ld(const ld* M, const ld* e, ld* r) {
  ld u[3];
  for (int k = 0; k < 3; ++k) u[k] = M[k * 4] * e[0] + M[k * 4 + 1] * e[1] + M[k * 4 + 2] * e[2] + M[k * 4 + 3];
  r[0] = e[3] - u[0] / u[2];
  r[1] = e[4] - u[1] / u[2];
}
static void distort_ld(const ld* s, const ld* e, ld* r) {
  ld p[3];
  for (int k = 0; k < 3; ++k) p[k] = s[k * 4] * e[0] + s[k * 4 + 1] * e[1] + s[k * 4 + 2] * e[2] + s[k * 4 + 3];
  const ld xn = p[0] / p[2], yn = p[1] / p[2], r2 = xn * xn + yn * yn;
  const ld radial = 1 + r2 * (s[16] + r2 * (s[17] + r2 * s[20]));
  const ld xy2 = 2 * xn * yn;
  const ld xd = xn * radial + s[18] * xy2 + s[19] * (2 * xn * xn + r2);
  const ld yd = yn * radial + s[18] * (2 * yn * yn + r2) + s[19] * xy2;
  r[0] = e[3] - (s[12] * xd + s[14]);
  r[1] = e[4] - (s[13] * yd + s[15]);
}

// sum of the magnitudes of the terms of one row of the affine stage
static double absaff(const double* row, const double* e) {
  return std::fabs(row[0] * e[0]) + std::fabs(row[1] * e[1]) + std::fabs(row[2] * e[2]) + std::fabs(row[3]);
}

template <class M, int SETN, typename CT>
static void quotient_hook(const double* set_ref, const double* D, double H, const double* e5, double* d_out) {
  CT s[SETN], Dj[SETN], e[5], ua[3], du[3], d[2];
  for (int i = 0; i < SETN; ++i) { s[i] = CT(set_ref[i]); Dj[i] = CT(D[i]); }
  for (int i = 0; i < 5; ++i) e[i] = CT(e5[i]);
  M::template affine<CT>(s, e, ua);
  M::template affine<CT>(Dj, e, du);
  M::template finish_diff<CT>(s, ua, du, CT(H), e, d);
  d_out[0] = double(d[0]);
  d_out[1] = double(d[1]);
}

int main() {
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  // a camera like tst/camera_calibration.cpp:22-30 looking down +Z at points 2..5 m away
  const double fx = 586.0, fy = 722.0, cx = 638.0, cy = 323.0;
  double worstP64 = 0, worstP32 = 0, worstD64 = 0, worstD32 = 0, worstT = 0, worstLit32 = 0;
  for (int trial = 0; trial < 20000; ++trial) {
    // rigid part: small rotation + translation; pinhole set M = K [R | t], distort set = [R | t], intrinsics
    double R[9] = {1, 0.02 * U(rng), 0.02 * U(rng), 0.02 * U(rng), 1, 0.02 * U(rng), 0.02 * U(rng), 0.02 * U(rng), 1};
    double t[3] = {0.1 * U(rng), 0.1 * U(rng), 0.1 * U(rng)};
    double TC[12], Mk[12];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) TC[r * 4 + c] = R[r * 3 + c];
      TC[r * 4 + 3] = t[r];
    }
    for (int c = 0; c < 4; ++c) {
      Mk[c] = fx * TC[c] + cx * TC[8 + c];
      Mk[4 + c] = fy * TC[4 + c] + cy * TC[8 + c];
      Mk[8 + c] = TC[8 + c];
    }
    double e[5] = {U(rng), 0.5 * U(rng), 3.5 + 1.5 * U(rng), 0, 0};
    {  // observed pixel = projection + ~0.5 px, as in the synthetic workloads (the residual itself is small)
      double u[3];
      PinholeModel::affine<double>(Mk, e, u);
      e[3] = u[0] / u[2] + 0.5 * U(rng);
      e[4] = u[1] / u[2] + 0.5 * U(rng);
    }
    // a perturbation direction of the set (what setup yields for x + h e_j minus the reference set, over H), and a
    // step of the size the reference's rule gives for float (sqrt(eps_f32) |x_j|, x_j ~ 1e-2) or double
    double Dm[12], Dt[24] = {0};
    for (int i = 0; i < 12; ++i) { Dm[i] = 50.0 * U(rng); Dt[i] = U(rng); }
    // (with the fp64 rule's step, 1.5e-8 |x_j|, even the long double literal quotient is rounding-limited at ~1e-9)
    const double H = (trial & 1) ? 3.45e-6 : 3.45e-4;
    // ---- pinhole -------------------------------------------------------------------------------------------
    {
      ld Ma[12], Mb[12], el[5], ra[2], rb[2];
      for (int i = 0; i < 12; ++i) { Ma[i] = Mk[i]; Mb[i] = ld(Mk[i]) + ld(H) * ld(Dm[i]); }
      for (int i = 0; i < 5; ++i) el[i] = e[i];
      pinhole_ld(Ma, el, ra);
      pinhole_ld(Mb, el, rb);
      double d64[2], d32[2];
      quotient_hook<PinholeModel, 12, double>(Mk, Dm, H, e, d64);
      quotient_hook<PinholeModel, 12, float>(Mk, Dm, H, e, d32);
      double ua[3], du[3];
      PinholeModel::affine<double>(Mk, e, ua);
      PinholeModel::affine<double>(Dm, e, du);
      for (int o = 0; o < 2; ++o) {
        const double lit = double((rb[o] - ra[o]) / ld(H));
        // natural scale of the column entry: the two products of the common-denominator numerator, with the
        // perturbation's first stage measured by its terms (a random direction may cancel inside affine())
        const double scale = (std::fabs(ua[o]) * absaff(Dm + 8, e) + absaff(Dm + 4 * o, e) * std::fabs(ua[2])) / (ua[2] * ua[2]);
        (void)du;
        worstP64 = std::fmax(worstP64, std::fabs(d64[o] - lit) / scale);
        worstP32 = std::fmax(worstP32, std::fabs(d32[o] - lit) / scale);
        // what the per-residual float form gives: two projections rounded to float (~640 px), subtracted
        const double pa = double(el[3 + o] - ra[o]), pb = double(el[3 + o] - rb[o]);
        const double litf = -double((float(pb) - float(pa)) / float(H));
        if (trial & 1) worstLit32 = std::fmax(worstLit32, std::fabs(litf - lit) / scale);
      }
    }
    // ---- pinhole + distortion ------------------------------------------------------------------------------
    {
      double s[24] = {0};
      for (int i = 0; i < 12; ++i) s[i] = TC[i];
      s[12] = fx; s[13] = fy; s[14] = cx; s[15] = cy;
      s[16] = -0.12; s[17] = 0.05; s[18] = 0.001; s[19] = -0.0007; s[20] = 0.01;
      ld sa[24], sb[24], el[5], ra[2], rb[2];
      for (int i = 0; i < 24; ++i) { sa[i] = s[i]; sb[i] = ld(s[i]) + ld(H) * ld(Dt[i]); }
      for (int i = 0; i < 5; ++i) el[i] = e[i];
      distort_ld(sa, el, ra);
      distort_ld(sb, el, rb);
      double d64[2], d32[2];
      quotient_hook<PinholeDistortModel, 24, double>(s, Dt, H, e, d64);
      quotient_hook<PinholeDistortModel, 24, float>(s, Dt, H, e, d32);
      double ua[3], du[3];
      PinholeDistortModel::affine<double>(s, e, ua);
      PinholeDistortModel::affine<double>(Dt, e, du);
      for (int o = 0; o < 2; ++o) {
        const double lit = double((rb[o] - ra[o]) / ld(H));
        const double scale = s[12 + o] * (std::fabs(ua[o]) * absaff(Dt + 8, e) + absaff(Dt + 4 * o, e) * std::fabs(ua[2])) / (ua[2] * ua[2]);
        (void)du;
        worstD64 = std::fmax(worstD64, std::fabs(d64[o] - lit) / scale);
        worstD32 = std::fmax(worstD32, std::fabs(d32[o] - lit) / scale);
      }
      // second-stage parameters: quotient with a large and a small step, both equal the partial derivative
      double sd[24], ed[5], td[3], Jt[18];
      for (int i = 0; i < 24; ++i) sd[i] = s[i];
      for (int i = 0; i < 5; ++i) ed[i] = e[i];
      {
        double u[3];
        PinholeDistortModel::affine<double>(sd, ed, u);
        td[0] = u[0] / u[2]; td[1] = u[1] / u[2]; td[2] = td[0] * td[0] + td[1] * td[1];
      }
      PinholeDistortModel::tail_partials<double>(sd, td, Jt);
      for (int k = 0; k < 9; ++k)
        for (int rep = 0; rep < 2; ++rep) {
          const ld h = (rep ? 1e-2L : 1e-5L) * std::fabs(ld(s[12 + k]));
          ld sp[24], sm[24], rp[2], rm[2];
          for (int i = 0; i < 24; ++i) sp[i] = sm[i] = s[i];
          sp[12 + k] += h; sm[12 + k] -= h;
          distort_ld(sp, el, rp);
          distort_ld(sm, el, rm);
          for (int o = 0; o < 2; ++o) {
            const double lit = double((rp[o] - rm[o]) / (2 * h));
            worstT = std::fmax(worstT, std::fabs(Jt[o * 9 + k] - lit) / (std::fabs(lit) + 1.0));
          }
        }
    }
  }
  std::printf("pinhole   finish_diff vs literal quotient (long double): fp64 %.2e  fp32 %.2e  (per-residual float form: %.2e)\n",
              worstP64, worstP32, worstLit32);
  std::printf("distorted finish_diff vs literal quotient (long double): fp64 %.2e  fp32 %.2e\n", worstD64, worstD32);
  std::printf("distorted tail_partials vs central quotient           : fp64 %.2e\n", worstT);
  CHECK(worstP64 < 1e-10, "pinhole fp64 %.3e", worstP64);
  CHECK(worstP32 < 2e-6, "pinhole fp32 %.3e", worstP32);
  CHECK(worstD64 < 1e-10, "distort fp64 %.3e", worstD64);
  CHECK(worstD32 < 2e-6, "distort fp32 %.3e", worstD32);
  CHECK(worstT < 1e-8, "tail partials %.3e", worstT);
  CHECK(worstLit32 > 100 * worstP32, "the per-residual float quotient should be far noisier (%.3e vs %.3e)", worstLit32, worstP32);
  return failures ? 1 : 0;
}
