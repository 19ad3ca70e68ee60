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
  template <typename CT>
  static __device__ __forceinline__ void residual(const CT* s, const CT (&e)[5], CT (&r)[2]) {
    CT u[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) u[k] = fma(s[k * 4 + 0], e[0], fma(s[k * 4 + 1], e[1], fma(s[k * 4 + 2], e[2], s[k * 4 + 3])));
    r[0] = e[3] - (u[0] / u[2]);
    r[1] = e[4] - (u[1] / u[2]);
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
