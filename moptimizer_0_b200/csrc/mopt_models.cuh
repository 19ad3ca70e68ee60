// Builtin device models: the workloads the reference defines as IBaseModel subclasses in its tests,
// restated as static device functions over one parameter set (the result of setup(x), see
// mopt_setup.cuh) and one residual's stream values.
//   P      parameters            O   residual dimension
//   NS     planar data streams   NA  streams in data group A (the rest are group B)
//   SETN   doubles per parameter set
#pragma once

#include "mopt_common.cuh"

namespace mopt {

#ifdef __CUDACC__

// hooks that tests/cpp/affine_fd_host_test.cu also runs on the host (nvcc -x cu, no GPU needed)
#define MOPT_HD __host__ __device__ __forceinline__

// tst/point2point.cpp:24-84.  streams: src x,y,z | tgt x,y,z.  set = R (9, row-major), t (3).
struct P2PModel {
  static constexpr int P = 6, O = 3, NS = 6, NA = 3, SETN = 12;
  static constexpr bool HAS_JAC = false;  // analytical linearization uses p2p_moment_kernel
  template <typename CT>
  static __device__ __forceinline__ void residual(const CT* s, const CT (&e)[6], CT (&r)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
      r[k] = (fma(s[k * 3 + 0], e[0], fma(s[k * 3 + 1], e[1], s[k * 3 + 2] * e[2])) + s[9 + k]) - e[3 + k];
  }
  template <typename CT>
  static __device__ __forceinline__ void residual_jacobian(const CT*, const CT (&)[6], CT (&)[3], CT (&)[18]) {}
};

// tst/parallel.cpp:12-32.  r = src - tgt, no parameters.
struct PointDistModel {
  static constexpr int P = 0, O = 3, NS = 6, NA = 3, SETN = 0;
  static constexpr bool HAS_JAC = false;
  template <typename CT>
  static __device__ __forceinline__ void residual(const CT*, const CT (&e)[6], CT (&r)[3]) {
    r[0] = e[0] - e[3]; r[1] = e[1] - e[4]; r[2] = e[2] - e[5];
  }
  template <typename CT>
  static __device__ __forceinline__ void residual_jacobian(const CT*, const CT (&)[6], CT (&)[3], CT (&)[3]) {}
};

// tst/curve_fitting.cpp:86-93.  streams: t | y.  set = x.
struct ExpCurveModel {
  static constexpr int P = 2, O = 1, NS = 2, NA = 1, SETN = 2;
  static constexpr bool HAS_JAC = true;
  // residual = finish(affine(set, e)) with u = m t + c linear in the set: the finite-difference quotient of the pass
  // kernels' AFFINE_FD mode is
  //   (finish(ua + H du) - finish(ua)) / H = -exp(ua) du g(H du),   g(z) = (e^z - 1) / z,
  // and exp(ua) = exp(u0) exp(ua - u0) with u0 the unperturbed stage (ua = u0 for forward differences, u0 - h t or
  // u0 - h for central ones).  Both arguments are of the order of the step: two short series and the residual's own
  // exponential replace the four (two) further exponentials of the literal quotient and their cancellation; a
  // perturbation too large for the series (|arg| >= 2^-5: huge |x_j| or |t|) takes the literal form.
  static constexpr int NAFF = 1;
  template <typename CT>
  static MOPT_HD void affine(const CT* s, const CT (&e)[2], CT (&u)[1]) { u[0] = fma(s[0], e[0], s[1]); }
  template <typename CT>
  static MOPT_HD void finish(const CT*, const CT (&u)[1], const CT (&e)[2], CT (&r)[1]) { r[0] = e[1] - exp(u[0]); }
  template <typename CT>
  static MOPT_HD void finish_diff(const CT* s0, const CT (&ua)[1], const CT (&du)[1], CT H, const CT (&e)[2], CT (&d)[1]) {
    CT u0[1];
    affine<CT>(s0, e, u0);
    const CT dl = ua[0] - u0[0];
    const CT z = H * du[0];
    if (all_abs_below(dl, 0.03125f) && all_abs_below(z, 0.03125f)) {
      const CT E0 = exp(u0[0]);  // the residual's own exponential (one evaluation after inlining)
      // e^dl to dl^5 / 120 (next term 1.3e-12) and g(z) to z^5 / 720 (next term 1.8e-13)
      const CT ed = fma(dl, fma(dl, fma(dl, fma(dl, fma(dl, CT(1.0 / 120), CT(1.0 / 24)), CT(1.0 / 6)), CT(0.5)), CT(1)), CT(1));
      const CT g = fma(z, fma(z, fma(z, fma(z, fma(z, CT(1.0 / 720), CT(1.0 / 120)), CT(1.0 / 24)), CT(1.0 / 6)), CT(0.5)), CT(1));
      d[0] = -((E0 * ed) * (du[0] * g));
    } else {
      d[0] = -((exp(fma(H, du[0], ua[0])) - exp(ua[0])) / H);
    }
  }
  template <typename CT>
  static __device__ __forceinline__ void residual(const CT* s, const CT (&e)[2], CT (&r)[1]) {
    r[0] = e[1] - exp(fma(s[0], e[0], s[1]));
  }
  template <typename CT>
  static __device__ __forceinline__ void residual_jacobian(const CT* s, const CT (&e)[2], CT (&r)[1], CT (&J)[2]) {
    const CT ex = exp(fma(s[0], e[0], s[1]));
    r[0] = e[1] - ex;
    J[0] = -e[0] * ex;
    J[1] = -ex;
  }
};

// tst/test_models.h:14-17, tst/differentiation.cpp:20-37.  streams: t | y.  set = x.
struct MichaelisMentenModel {
  static constexpr int P = 2, O = 1, NS = 2, NA = 1, SETN = 2;
  static constexpr bool HAS_JAC = true;
  template <typename CT>
  static __device__ __forceinline__ void residual(const CT* s, const CT (&e)[2], CT (&r)[1]) {
    r[0] = e[1] - (s[0] * e[0]) / (s[1] + e[0]);
  }
  template <typename CT>
  static __device__ __forceinline__ void residual_jacobian(const CT* s, const CT (&e)[2], CT (&r)[1], CT (&J)[2]) {
    const CT den = s[1] + e[0];
    r[0] = e[1] - (s[0] * e[0]) / den;
    J[0] = -e[0] / den;
    J[1] = (s[0] * e[0]) / (den * den);
  }
};

// tst/camera_calibration.cpp:35-41.  streams: X,Y,Z | u,v.  set = M = K T(x) C (3x4 row-major).
struct PinholeModel {
  static constexpr int P = 6, O = 2, NS = 5, NA = 3, SETN = 12;
  static constexpr bool HAS_JAC = false;
  // The residual is finish(affine(set, e)) with a first stage that is LINEAR in the set (u = M [X Y Z 1]^T), so
  // the difference of two first stages is the first stage of the difference of the sets.  The AFFINE_FD mode of
  // the pass kernels uses that to form the finite-difference quotient without subtracting two rounded quotients.
  static constexpr int NAFF = 3;
  template <typename CT>
  static MOPT_HD void affine(const CT* s, const CT (&e)[5], CT (&u)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) u[k] = fma(s[k * 4 + 0], e[0], fma(s[k * 4 + 1], e[1], fma(s[k * 4 + 2], e[2], s[k * 4 + 3])));
  }
  template <typename CT>
  static MOPT_HD void finish(const CT*, const CT (&u)[3], const CT (&e)[5], CT (&r)[2]) {
    r[0] = e[3] - (u[0] / u[2]);
    r[1] = e[4] - (u[1] / u[2]);
  }
  // (finish(ua + H du) - finish(ua)) / H for the perspective division, over the common denominator:
  //   ((ua0 + H du0) / (ua2 + H du2) - ua0 / ua2) / H = (du0 ua2 - ua0 du2) / (ua2 (ua2 + H du2))
  template <typename CT>
  static MOPT_HD void finish_diff(const CT*, const CT (&ua)[3], const CT (&du)[3], CT H, const CT (&)[5], CT (&d)[2]) {
    const CT inv = fast_rcp(ua[2] * fma(H, du[2], ua[2]));
    d[0] = fma(ua[0], du[2], -(du[0] * ua[2])) * inv;  // r = pix - u0/u2: the sign is flipped
    d[1] = fma(ua[1], du[2], -(du[1] * ua[2])) * inv;
  }
  template <typename CT>
  static MOPT_HD void residual(const CT* s, const CT (&e)[5], CT (&r)[2]) {
    CT u[3];
    affine<CT>(s, e, u);
    finish<CT>(s, u, e, r);
  }
  template <typename CT>
  static __device__ __forceinline__ void residual_jacobian(const CT*, const CT (&)[5], CT (&)[2], CT (&)[12]) {}
};

// Pinhole + OpenCV distortion with free intrinsics (new; BASELINE.json configs[4]).  streams: X,Y,Z | u,v.
// set = TC (3x4 row-major, 12), fx, fy, cx, cy, k1, k2, p1, p2, k3.
struct PinholeDistortModel {
  static constexpr int P = 15, O = 2, NS = 5, NA = 3, SETN = 21;
  static constexpr bool HAS_JAC = false;
  // Two-stage residual for the wide kernel: the extrinsics (x[0..6) -> set[0..12)) only enter the rigid transform
  // and the perspective division; perturbing an intrinsic or distortion parameter re-runs stage 2 alone.
  static constexpr int STAGE1_PARAMS = 6, STAGE1_VALUES = 3;
  template <typename CT>
  static __device__ __forceinline__ void stage1(const CT* s, const CT (&e)[5], CT (&t)[3]) {
    CT p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = fma(s[k * 4 + 0], e[0], fma(s[k * 4 + 1], e[1], fma(s[k * 4 + 2], e[2], s[k * 4 + 3])));
    const CT iz = CT(1) / p[2];
    t[0] = p[0] * iz;
    t[1] = p[1] * iz;
    t[2] = fma(t[0], t[0], t[1] * t[1]);
  }
  template <typename CT>
  static __device__ __forceinline__ void stage2(const CT* s, const CT (&e)[5], const CT (&t)[3], CT (&r)[2]) {
    const CT xn = t[0], yn = t[1], r2 = t[2];
    const CT radial = fma(r2, fma(r2, fma(r2, s[20], s[17]), s[16]), CT(1));  // 1 + k1 r2 + k2 r2^2 + k3 r2^3
    const CT xy2 = CT(2) * xn * yn;
    const CT xd = fma(xn, radial, fma(s[18], xy2, s[19] * fma(CT(2) * xn, xn, r2)));
    const CT yd = fma(yn, radial, fma(s[18], fma(CT(2) * yn, yn, r2), s[19] * xy2));
    r[0] = e[3] - fma(s[12], xd, s[14]);
    r[1] = e[4] - fma(s[13], yd, s[15]);
  }
  template <typename CT>
  static __device__ __forceinline__ void residual(const CT* s, const CT (&e)[5], CT (&r)[2]) {
    CT t[3];
    stage1<CT>(s, e, t);
    stage2<CT>(s, e, t, r);
  }
  // ---- hooks of the AFFINE_FD mode (see PinholeModel): u = (T C) [X Y Z 1]^T is linear in set[0..12) ----------
  static constexpr int NAFF = 3;
  template <typename CT>
  static MOPT_HD void affine(const CT* s, const CT (&e)[5], CT (&u)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) u[k] = fma(s[k * 4 + 0], e[0], fma(s[k * 4 + 1], e[1], fma(s[k * 4 + 2], e[2], s[k * 4 + 3])));
  }
  // (f(ua + H du) - f(ua)) / H for the parameters that enter through u (the extrinsics), every difference taken
  // in product form: D(a b) = Da b_b + a_a Db, D(a^2) = (a_b + a_a) Da, D(a^3) = (a_b^2 + a_b a_a + a_a^2) Da,
  // so nothing is obtained by subtracting two nearly equal rounded values.  `s` is the base set (fx .. k3).
  template <typename CT>
  static MOPT_HD void finish_diff(const CT* s, const CT (&ua)[3], const CT (&du)[3], CT H, const CT (&)[5], CT (&d)[2]) {
    const CT ia = fast_rcp(ua[2]), ib = fast_rcp(fma(H, du[2], ua[2]));
    const CT xa = ua[0] * ia, ya = ua[1] * ia;
    const CT iab = ia * ib;
    const CT Dx = fma(du[0], ua[2], -(ua[0] * du[2])) * iab;  // (x_b - x_a) / H over the common denominator
    const CT Dy = fma(du[1], ua[2], -(ua[1] * du[2])) * iab;
    const CT xb = fma(H, Dx, xa), yb = fma(H, Dy, ya);
    const CT sx = xb + xa, sy = yb + ya;
    const CT r2a = fma(xa, xa, ya * ya);
    const CT Dr2 = fma(sx, Dx, sy * Dy);
    const CT r2b = fma(H, Dr2, r2a);
    const CT k1 = s[16], k2 = s[17], p1 = s[18], p2 = s[19], k3 = s[20];
    const CT radial_a = fma(r2a, fma(r2a, fma(r2a, k3, k2), k1), CT(1));
    const CT Dradial = Dr2 * fma(k3, fma(r2b, r2b, fma(r2b, r2a, r2a * r2a)), fma(k2, r2b + r2a, k1));
    const CT radial_b = fma(H, Dradial, radial_a);
    const CT Dxy2 = CT(2) * fma(Dx, yb, xa * Dy);
    const CT Dxx = fma(CT(2) * sx, Dx, Dr2);  // D(2 x^2 + r2)
    const CT Dyy = fma(CT(2) * sy, Dy, Dr2);  // D(2 y^2 + r2)
    const CT Dxd = fma(Dx, radial_b, fma(xa, Dradial, fma(p1, Dxy2, p2 * Dxx)));
    const CT Dyd = fma(Dy, radial_b, fma(ya, Dradial, fma(p1, Dyy, p2 * Dxy2)));
    d[0] = -(s[12] * Dxd);
    d[1] = -(s[13] * Dyd);
  }
  // Columns STAGE1_PARAMS .. P-1 of J (fx fy cx cy k1 k2 p1 p2 k3): the residual is affine in each of them taken
  // alone, so the reference's difference quotient in that parameter (linearization.h:105; central likewise) IS the
  // partial derivative, whatever the step.  t = stage1 values at the base set.  Jt row-major O x (P - STAGE1_PARAMS).
  template <typename CT>
  static MOPT_HD void tail_partials(const CT* s, const CT (&t)[3], CT (&Jt)[18]) {
    const CT xn = t[0], yn = t[1], r2 = t[2];
    const CT fx = s[12], fy = s[13], k1 = s[16], k2 = s[17], p1 = s[18], p2 = s[19], k3 = s[20];
    const CT radial = fma(r2, fma(r2, fma(r2, k3, k2), k1), CT(1));
    const CT xy2 = CT(2) * xn * yn;
    const CT xx = fma(CT(2) * xn, xn, r2), yy = fma(CT(2) * yn, yn, r2);
    const CT xd = fma(xn, radial, fma(p1, xy2, p2 * xx));
    const CT yd = fma(yn, radial, fma(p1, yy, p2 * xy2));
    const CT fxx = fx * xn, fyy = fy * yn, r4 = r2 * r2, r6 = r4 * r2;
    Jt[0] = -xd;        Jt[1] = CT(0);      Jt[2] = CT(-1);     Jt[3] = CT(0);
    Jt[4] = -(fxx * r2); Jt[5] = -(fxx * r4); Jt[6] = -(fx * xy2); Jt[7] = -(fx * xx); Jt[8] = -(fxx * r6);
    Jt[9] = CT(0);      Jt[10] = -yd;       Jt[11] = CT(0);     Jt[12] = CT(-1);
    Jt[13] = -(fyy * r2); Jt[14] = -(fyy * r4); Jt[15] = -(fy * yy); Jt[16] = -(fy * xy2); Jt[17] = -(fyy * r6);
  }
  template <typename CT>
  static __device__ __forceinline__ void residual_jacobian(const CT*, const CT (&)[5], CT (&)[2], CT (&)[30]) {}
};

// tst/powell.cpp:22-59.  No data.  set = x.
struct PowellModel {
  static constexpr int P = 4, O = 4, NS = 0, NA = 0, SETN = 4;
  static constexpr bool HAS_JAC = true;
  template <typename CT>
  static __device__ __forceinline__ void residual(const CT* x, const CT (&)[1], CT (&r)[4]) {
    r[0] = x[0] + CT(10) * x[1];
    r[1] = sqrt(CT(5)) * (x[2] - x[3]);
    r[2] = (x[1] - CT(2) * x[2]) * (x[1] - CT(2) * x[2]);
    r[3] = sqrt(CT(10)) * (x[0] - x[3]) * (x[0] - x[3]);
  }
  template <typename CT>
  static __device__ __forceinline__ void residual_jacobian(const CT* x, const CT (&e)[1], CT (&r)[4], CT (&J)[16]) {
    residual<CT>(x, e, r);
#pragma unroll
    for (int i = 0; i < 16; ++i) J[i] = CT(0);
    // row-major O x P; the entries (and the sign of d f2/d x1) are those of tst/powell.cpp:35-56
    J[0] = CT(1);
    J[12] = sqrt(CT(10)) * CT(2) * (x[0] - x[3]);
    J[1] = CT(10);
    J[9] = CT(2) * (x[1] + CT(2) * x[2]);
    J[6] = sqrt(CT(5));
    J[10] = CT(2) * (x[1] + CT(2) * x[2]) * CT(-2);
    J[7] = -sqrt(CT(5));
    J[15] = sqrt(CT(10)) * CT(2) * (x[0] - x[3]) * CT(-1);
  }
};

#endif  // __CUDACC__

// Host-visible shape table (same numbers as the structs above).
struct ModelShape {
  int P, O, NS, NA;
  int ncomp_a, ncomp_b;  // components of the host AoS groups A and B
  bool has_analytical;
};
ModelShape user_model_shape(int model);  // mopt_rtc.cu: run-time compiled user models (ids >= MOPT_MODEL_USER_BASE)
inline ModelShape model_shape(int model) {
  if (model >= MOPT_MODEL_USER_BASE) return user_model_shape(model);
  switch (model) {
    case MOPT_MODEL_POINT2POINT: return {6, 3, 6, 3, 3, 3, true};
    case MOPT_MODEL_EXP_CURVE: return {2, 1, 2, 1, 1, 1, true};
    case MOPT_MODEL_MICHAELIS_MENTEN: return {2, 1, 2, 1, 1, 1, true};
    case MOPT_MODEL_PINHOLE: return {6, 2, 5, 3, 3, 2, false};
    case MOPT_MODEL_POWELL: return {4, 4, 0, 0, 0, 0, true};
    case MOPT_MODEL_POINT_DIST: return {0, 3, 6, 3, 3, 3, false};
    case MOPT_MODEL_PINHOLE_DISTORT: return {15, 2, 5, 3, 3, 2, false};
    default: return {-1, -1, -1, -1, 0, 0, false};
  }
}

}  // namespace mopt
