// ============================================================================
// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Plain C++17 restatement (no Eigen, no TBB) of the reference's linearization
// hot path and the Levenberg-Marquardt step that consumes it.  Only `tests/`,
// `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
// leg may build, load or call anything under `oracle/`.  The product (the CUDA
// library under moptimizer_0_b200/csrc) never links or calls this code.
//
// Parity pinning: the reference cannot be compiled in this image (Eigen3, oneTBB
// and GoogleTest are absent, no network), so this restatement is pinned against
// every known-answer test the reference holds for this path (see
// tests/test_oracle_golden.py and SURVEY.md §8c): curve fitting, camera
// calibration, Powell, Michaelis-Menten (+Geman-McClure), covariance scaling,
// split-cost equality, analytical-vs-numerical Hessians, serial==parallel cost.
// Absolute H/b values are pinned by this restatement only.
//
// Third-party arithmetic that is NOT under /root/reference and is restated here
// from its published algorithm:
//   * Eigen 3.4.0 (libeigen3-dev of ubuntu-22.04, unpinned by the reference):
//     dense products J^T C J (trivial) and Eigen::LDLT (robust Cholesky with
//     diagonal pivoting, Eigen/src/Cholesky/LDLT.h, `ldlt_inplace<Lower>::unblocked`
//     and `_solve_impl_transposed`) used at src/levenberg_marquadt_dyn.cpp:78-80.
//   * oneTBB parallel_reduce (linearization.h:52) — a plain chunked sum.
//
// All `file:line` citations are relative to /root/reference.
// ============================================================================
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace oracle {

// ---------------------------------------------------------------- status ----
// include/moptimizer/types.h:6-12
enum Status {
  CONVERGED = 0,
  MAXIMUM_ITERATIONS_REACHED = 1,
  SMALL_DELTA = 2,
  NUMERIC_ERROR = 3,
  FATAL_ERROR = 4,
};

// ------------------------------------------------------------------- so3 ----
// src/so3.cpp:43-57 — Rodrigues with the `norm > 10*eps` guard (Ref overload).
template <class S>
inline void so3_exp(const S w[3], S R[9] /*row-major 3x3*/) {
  const S n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? S(1) : S(0);
  if (n > S(10.0) * std::numeric_limits<S>::epsilon()) {
    const S a[3] = {w[0] / n, w[1] / n, w[2] / n};
    // K = [a]x ; K*K = a a^T - I (|a| = 1 up to rounding; computed explicitly as the
    // reference does, by a 3x3 product).
    const S K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    S KK[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        S s = 0;
        for (int k = 0; k < 3; ++k) s += K[r * 3 + k] * K[k * 3 + c];
        KK[r * 3 + c] = s;
      }
    const S sn = std::sin(n), cs = std::cos(n);
    for (int i = 0; i < 9; ++i) R[i] += sn * K[i] + (S(1.0) - cs) * KK[i];
  }
}

// src/so3.cpp:7-19 — x = [t(3), omega(3)] -> 4x4 homogeneous transform (row-major here).
template <class S>
inline void so3_convert6dof(const S* x, S T[16]) {
  S R[9];
  const S w[3] = {x[3], x[4], x[5]};
  so3_exp(w, R);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[r * 4 + c] = R[r * 3 + c];
    T[r * 4 + 3] = x[r];
  }
  T[12] = T[13] = T[14] = 0;
  T[15] = 1;
}

// src/so3.cpp:96-105 — so3::Log exactly as the reference defines it.  R row-major.
template <class S>
inline void so3_log(const S R[9], S w[3]) {
  const S tr = R[0] + R[4] + R[8];
  const S theta = (tr > S(3.0) - S(1e-6)) ? S(0.0) : std::acos(S(0.5) * (tr - S(1)));
  const S K[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  if (std::fabs(theta) < S(0.001)) {
    for (int i = 0; i < 3; ++i) w[i] = S(0.5) * K[i];
  } else {
    for (int i = 0; i < 3; ++i) w[i] = S(0.5) * theta / std::sin(theta) * K[i];
  }
}

// x (+) delta: additive (levenberg_marquadt_dyn.cpp:82-83) or, as the opt-in that finishes the reference's
// "TODO Manifold operation" (SURVEY.md §8f-3), a left SO(3) perturbation of the rotation-vector block x[3..5]:
// omega <- Log(Exp(delta_omega) Exp(omega)).
enum Manifold { MANIFOLD_ADDITIVE = 0, MANIFOLD_SO3_LEFT = 1 };
template <class S>
inline void retract(int manifold, int P, const S* x, const S* delta, S* out) {
  for (int i = 0; i < P; ++i) out[i] = x[i] + delta[i];
  if (manifold == MANIFOLD_SO3_LEFT && P >= 6) {
    S Rx[9], Rd[9], Rn[9], w[3];
    so3_exp(x + 3, Rx);
    so3_exp(delta + 3, Rd);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        S s = 0;
        for (int k = 0; k < 3; ++k) s += Rd[r * 3 + k] * Rx[k * 3 + c];
        Rn[r * 3 + c] = s;
      }
    so3_log(Rn, w);
    for (int i = 0; i < 3; ++i) out[3 + i] = w[i];
  }
}

// Full closed-form left Jacobian of SO(3):  I + (1-cos)/th^2 [w]x + (th-sin)/th^3 [w]x^2.
// (The reference's so3::leftJacobian, src/so3.cpp:141-155, omits the [w]x^2 term; the
//  exact form is needed for an analytical Jacobian that agrees with finite differences.)
template <class S>
inline void so3_left_jacobian_full(const S w[3], S Jl[9]) {
  const S th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const S th = std::sqrt(th2);
  S A, B;
  if (th2 < S(1e-8)) {
    A = S(0.5) - th2 / S(24);
    B = S(1) / S(6) - th2 / S(120);
  } else {
    A = (S(1) - std::cos(th)) / th2;
    B = (th - std::sin(th)) / (th2 * th);
  }
  const S K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      S kk = 0;
      for (int k = 0; k < 3; ++k) kk += K[r * 3 + k] * K[k * 3 + c];
      Jl[r * 3 + c] = (r == c ? S(1) : S(0)) + A * K[r * 3 + c] + B * kk;
    }
}

// ------------------------------------------------------------------ loss ----
// include/moptimizer/loss_function/loss_function.h:7-23
template <class S>
struct ILoss {
  virtual ~ILoss() = default;
  virtual S weight(S e2) = 0;
};
template <class S>
struct NoLoss : ILoss<S> {
  S weight(S) override { return S(1.0); }
};
// include/moptimizer/loss_function/geman_mcclure.h:11-13
template <class S>
struct GemanMcClure : ILoss<S> {
  explicit GemanMcClure(S th) : th_(th) {}
  S weight(S e2) override {
    const S num = th_ * th_;
    const S den = (e2 + th_) * (e2 + th_);
    return num / den;
  }
  S th_;
};
// NOT in the reference: Huber IRLS weight in terms of e2 = |r|^2,
//   w = 1 if e2 <= k^2 else k / sqrt(e2).
template <class S>
struct Huber : ILoss<S> {
  explicit Huber(S k) : k_(k) {}
  S weight(S e2) override { return (e2 <= k_ * k_) ? S(1) : k_ / std::sqrt(e2); }
  S k_;
};

// ----------------------------------------------------------------- model ----
// include/moptimizer/model.h:11-47
template <class S>
struct IModel {
  virtual ~IModel() = default;
  virtual void setup(const S* x) = 0;
  virtual void update(const S*) {}
  virtual bool f(const S* x, S* r, unsigned i) const = 0;
  virtual bool f_df(const S* x, S* r, S* J /*row-major OxP*/, unsigned i) const = 0;
  virtual std::shared_ptr<IModel<S>> clone() const = 0;
  virtual bool has_jacobian() const { return true; }
};

enum P2PJacobian {
  P2P_EXACT = 0,             // [I | -[R p]x J_l(w)]  — exact for the additive rot-vector update
  P2P_REFTEST = 1,           // [I | -[p]x] stored row-major (tst/point2point.cpp:72-75, layout fixed)
  P2P_REFTEST_COLMAJOR = 2,  // same values written column-major into the row-major buffer,
                             // bit-faithful to tst/point2point.cpp:18,71 (scrambled H)
  P2P_LEFT = 3,              // [I | -[R p]x]: derivative w.r.t. a left SO(3) perturbation (MANIFOLD_SO3_LEFT)
};

// tst/point2point.cpp:24-84 — r = T p - q;  data as AoS xyz.
template <class S>
struct Point2Point : IModel<S> {
  Point2Point(const S* src, const S* tgt, int jac) : src_(src), tgt_(tgt), jac_(jac) {
    for (int i = 0; i < 16; ++i) T_[i] = (i % 5 == 0) ? S(1) : S(0);
    for (int i = 0; i < 9; ++i) Jl_[i] = (i % 4 == 0) ? S(1) : S(0);
  }
  void setup(const S* x) override {
    so3_convert6dof(x, T_);
    const S w[3] = {x[3], x[4], x[5]};
    so3_left_jacobian_full(w, Jl_);
  }
  inline void residual(unsigned i, S* r, S* wp) const {
    const S* p = src_ + 3 * size_t(i);
    const S* q = tgt_ + 3 * size_t(i);
    for (int k = 0; k < 3; ++k) {
      // (T * [p,1])[k] - q[k]   (tst/point2point.cpp:41-44)
      const S rot = T_[k * 4 + 0] * p[0] + T_[k * 4 + 1] * p[1] + T_[k * 4 + 2] * p[2];
      wp[k] = rot;  // R p (without translation)
      r[k] = (rot + T_[k * 4 + 3] * S(1)) - q[k];
    }
  }
  bool f(const S*, S* r, unsigned i) const override {
    S wp[3];
    residual(i, r, wp);
    return true;
  }
  bool f_df(const S*, S* r, S* J, unsigned i) const override {
    S wp[3];
    residual(i, r, wp);
    const S* p = src_ + 3 * size_t(i);
    const S* v = (jac_ == P2P_EXACT || jac_ == P2P_LEFT) ? wp : p;
    // -[v]x
    const S nsk[9] = {0, v[2], -v[1], -v[2], 0, v[0], v[1], -v[0], 0};
    S right[9];
    if (jac_ == P2P_EXACT) {
      for (int r_ = 0; r_ < 3; ++r_)
        for (int c = 0; c < 3; ++c) {
          S s = 0;
          for (int k = 0; k < 3; ++k) s += nsk[r_ * 3 + k] * Jl_[k * 3 + c];
          right[r_ * 3 + c] = s;
        }
    } else {
      for (int k = 0; k < 9; ++k) right[k] = nsk[k];
    }
    if (jac_ == P2P_REFTEST_COLMAJOR) {
      // Eigen::Map<Matrix<S,3,6>> is column-major: element (r,c) lands at J[c*3+r].
      for (int r_ = 0; r_ < 3; ++r_)
        for (int c = 0; c < 3; ++c) {
          J[c * 3 + r_] = (r_ == c) ? S(1) : S(0);
          J[(c + 3) * 3 + r_] = right[r_ * 3 + c];
        }
    } else {
      for (int r_ = 0; r_ < 3; ++r_)
        for (int c = 0; c < 3; ++c) {
          J[r_ * 6 + c] = (r_ == c) ? S(1) : S(0);
          J[r_ * 6 + 3 + c] = right[r_ * 3 + c];
        }
    }
    return true;
  }
  std::shared_ptr<IModel<S>> clone() const override { return std::make_shared<Point2Point>(*this); }
  const S* src_;
  const S* tgt_;
  int jac_;
  S T_[16];
  S Jl_[9];
};

// Point-to-point with correspondence re-association in update(x) — the hook the reference declares
// (model.h:24-26 "i.e registration correspondences", called at levenberg_marquadt_dyn.cpp:54) and never
// implements.  Brute-force exact nearest neighbour in the fixed target cloud (ties -> lowest index), within
// max_dist; a source point without a match makes f / f_df return false, so it is skipped (linearization.h:102,144).
template <class S>
struct Point2PointICP : IModel<S> {
  Point2PointICP(const S* src, int n, const S* target, int m, S max_dist, int jac)
      : src_(src), n_(n), target_(target), m_(m), max_dist_(max_dist), jac_(jac), tgt_(size_t(n) * 3, S(0)),
        valid_(size_t(n), 0), inner_(src, tgt_.data(), jac) {}
  Point2PointICP(const Point2PointICP& o)
      : src_(o.src_), n_(o.n_), target_(o.target_), m_(o.m_), max_dist_(o.max_dist_), jac_(o.jac_), tgt_(o.tgt_),
        valid_(o.valid_), inner_(o.inner_) {
    inner_.tgt_ = tgt_.data();
  }
  void setup(const S* x) override { inner_.setup(x); }
  void update(const S* x) override {
    S T[16];
    so3_convert6dof(x, T);
    const S r2max = max_dist_ * max_dist_;
    for (int i = 0; i < n_; ++i) {
      const S* p = src_ + 3 * size_t(i);
      S q[3];
      for (int k = 0; k < 3; ++k) q[k] = T[k * 4] * p[0] + T[k * 4 + 1] * p[1] + T[k * 4 + 2] * p[2] + T[k * 4 + 3];
      S best = r2max;
      int arg = -1;
      for (int j = 0; j < m_; ++j) {
        const S* t = target_ + 3 * size_t(j);
        const S d0 = t[0] - q[0], d1 = t[1] - q[1], d2 = t[2] - q[2];
        const S dd = d0 * d0 + d1 * d1 + d2 * d2;
        if (dd < best || (arg < 0 && dd <= r2max)) {
          best = dd;
          arg = j;
        }
      }
      valid_[i] = arg >= 0;
      if (arg >= 0)
        for (int k = 0; k < 3; ++k) tgt_[3 * size_t(i) + k] = target_[3 * size_t(arg) + k];
    }
  }
  bool f(const S* x, S* r, unsigned i) const override { return valid_[i] ? inner_.f(x, r, i) : false; }
  bool f_df(const S* x, S* r, S* J, unsigned i) const override { return valid_[i] ? inner_.f_df(x, r, J, i) : false; }
  std::shared_ptr<IModel<S>> clone() const override { return std::make_shared<Point2PointICP>(*this); }
  int matched() const {
    int c = 0;
    for (char v : valid_) c += v;
    return c;
  }
  const S* src_;
  int n_;
  const S* target_;
  int m_;
  S max_dist_;
  int jac_;
  std::vector<S> tgt_;
  std::vector<char> valid_;
  Point2Point<S> inner_;
};

// tst/parallel.cpp:12-32 — r = src - tgt, no parameters.
template <class S>
struct PointDist : IModel<S> {
  PointDist(const S* src, const S* tgt) : src_(src), tgt_(tgt) {}
  void setup(const S*) override {}
  bool f(const S*, S* r, unsigned i) const override {
    for (int k = 0; k < 3; ++k) r[k] = src_[3 * size_t(i) + k] - tgt_[3 * size_t(i) + k];
    return true;
  }
  bool f_df(const S*, S*, S*, unsigned) const override { return false; }
  bool has_jacobian() const override { return false; }
  std::shared_ptr<IModel<S>> clone() const override { return std::make_shared<PointDist>(*this); }
  const S* src_;
  const S* tgt_;
};

// tst/curve_fitting.cpp:81-98 — r = y - exp(x0 t + x1).
template <class S>
struct ExpCurve : IModel<S> {
  ExpCurve(const S* t, const S* y) : t_(t), y_(y) {}
  void setup(const S*) override {}
  bool f(const S* x, S* r, unsigned i) const override {
    r[0] = y_[i] - std::exp(x[0] * t_[i] + x[1]);
    return true;
  }
  bool f_df(const S* x, S* r, S* J, unsigned i) const override {  // analytical J: addition
    const S e = std::exp(x[0] * t_[i] + x[1]);
    r[0] = y_[i] - e;
    J[0] = -t_[i] * e;
    J[1] = -e;
    return true;
  }
  std::shared_ptr<IModel<S>> clone() const override { return std::make_shared<ExpCurve>(*this); }
  const S* t_;
  const S* y_;
};

// tst/test_models.h:8-19 and tst/differentiation.cpp:16-41 — r = y - x0 t / (x1 + t).
template <class S>
struct MichaelisMenten : IModel<S> {
  MichaelisMenten(const S* t, const S* y) : t_(t), y_(y) {}
  void setup(const S*) override {}
  bool f(const S* x, S* r, unsigned i) const override {
    r[0] = y_[i] - (x[0] * t_[i]) / (x[1] + t_[i]);
    return true;
  }
  bool f_df(const S* x, S* r, S* J, unsigned i) const override {
    const S den = x[1] + t_[i];
    r[0] = y_[i] - (x[0] * t_[i]) / (x[1] + t_[i]);
    J[0] = -t_[i] / den;
    J[1] = (x[0] * t_[i]) / (den * den);
    return true;
  }
  std::shared_ptr<IModel<S>> clone() const override {
    return std::make_shared<MichaelisMenten>(*this);
  }
  const S* t_;
  const S* y_;
};

// tst/powell.cpp:22-59 — Powell's singular function, 4 outputs, 4 parameters, no data.
template <class S>
struct Powell : IModel<S> {
  void setup(const S*) override {}
  bool f(const S* x, S* r, unsigned) const override {
    r[0] = x[0] + 10 * x[1];
    r[1] = std::sqrt(S(5)) * (x[2] - x[3]);
    r[2] = (x[1] - 2 * x[2]) * (x[1] - 2 * x[2]);
    r[3] = std::sqrt(S(10)) * (x[0] - x[3]) * (x[0] - x[3]);
    return true;
  }
  bool f_df(const S* x, S* r, S* J, unsigned i) const override {
    f(x, r, i);
    for (int k = 0; k < 16; ++k) J[k] = 0;
    // row-major: J[row*4 + col]; values as written at tst/powell.cpp:35-56 (including the
    // reference's sign on d f2 / d x1, `2 (x1 + 2 x2)`).
    J[0] = 1;
    J[12] = std::sqrt(S(10)) * 2 * (x[0] - x[3]);
    J[1] = 10;
    J[9] = 2 * (x[1] + 2 * x[2]);
    J[6] = std::sqrt(S(5));
    J[10] = 2 * (x[1] + 2 * x[2]) * (-2);
    J[7] = -std::sqrt(S(5));
    J[15] = std::sqrt(S(10)) * 2 * (x[0] - x[3]) * (-1);
    return true;
  }
  std::shared_ptr<IModel<S>> clone() const override { return std::make_shared<Powell>(*this); }
};

// tst/camera_calibration.cpp:12-57 — r = pixel - proj(K * T(x) * C * P), 6-DoF extrinsics.
// `K` 3x4 row-major, `C` 4x4 row-major (camera<-laser frame conversion) are model constants.
template <class S>
struct Pinhole : IModel<S> {
  Pinhole(const S* pts /*xyz AoS, w=1*/, const S* pix /*uv AoS*/, const S* K34, const S* C44)
      : pts_(pts), pix_(pix) {
    for (int i = 0; i < 12; ++i) K_[i] = K34[i];
    for (int i = 0; i < 16; ++i) C_[i] = C44[i];
    for (int i = 0; i < 12; ++i) M_[i] = 0;
  }
  void setup(const S* x) override {
    S T[16];
    so3_convert6dof(x, T);
    // M = (K * T) * C, evaluated left to right as Eigen does for the chained product
    // at tst/camera_calibration.cpp:37.
    S KT[12];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        S s = 0;
        for (int k = 0; k < 4; ++k) s += K_[r * 4 + k] * T[k * 4 + c];
        KT[r * 4 + c] = s;
      }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        S s = 0;
        for (int k = 0; k < 4; ++k) s += KT[r * 4 + k] * C_[k * 4 + c];
        M_[r * 4 + c] = s;
      }
  }
  bool f(const S*, S* r, unsigned i) const override {
    const S* P = pts_ + 3 * size_t(i);
    S u[3];
    for (int k = 0; k < 3; ++k)
      u[k] = M_[k * 4 + 0] * P[0] + M_[k * 4 + 1] * P[1] + M_[k * 4 + 2] * P[2] + M_[k * 4 + 3];
    r[0] = pix_[2 * size_t(i) + 0] - (u[0] / u[2]);
    r[1] = pix_[2 * size_t(i) + 1] - (u[1] / u[2]);
    return true;
  }
  bool f_df(const S*, S*, S*, unsigned) const override { return false; }
  bool has_jacobian() const override { return false; }
  std::shared_ptr<IModel<S>> clone() const override { return std::make_shared<Pinhole>(*this); }
  const S* pts_;
  const S* pix_;
  S K_[12], C_[16], M_[12];
};

// NOT in the reference (its camera test has fixed intrinsics and no distortion): pinhole
// with free intrinsics and OpenCV-style distortion, x = [t, omega, fx, fy, cx, cy, k1, k2, p1, p2, k3],
//   Pc = T(x) C P;  (xn, yn) = Pc.xy / Pc.z;  r2 = xn^2 + yn^2;  radial = 1 + k1 r2 + k2 r2^2 + k3 r2^3
//   xd = xn radial + 2 p1 xn yn + p2 (r2 + 2 xn^2);  yd = yn radial + p1 (r2 + 2 yn^2) + 2 p2 xn yn
//   r = pixel - (fx xd + cx, fy yd + cy).
template <class S>
struct PinholeDistort : IModel<S> {
  PinholeDistort(const S* pts, const S* pix, const S* C44) : pts_(pts), pix_(pix) {
    for (int i = 0; i < 16; ++i) C_[i] = C44[i];
    for (int i = 0; i < 12; ++i) TC_[i] = 0;
    for (int i = 0; i < 9; ++i) in_[i] = 0;
  }
  void setup(const S* x) override {
    S T[16];
    so3_convert6dof(x, T);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        S s = 0;
        for (int k = 0; k < 4; ++k) s += T[r * 4 + k] * C_[k * 4 + c];
        TC_[r * 4 + c] = s;
      }
    for (int i = 0; i < 9; ++i) in_[i] = x[6 + i];
  }
  bool f(const S*, S* r, unsigned i) const override {
    const S* P = pts_ + 3 * size_t(i);
    S p[3];
    for (int k = 0; k < 3; ++k)
      p[k] = TC_[k * 4 + 0] * P[0] + TC_[k * 4 + 1] * P[1] + TC_[k * 4 + 2] * P[2] + TC_[k * 4 + 3];
    const S xn = p[0] / p[2], yn = p[1] / p[2];
    const S r2 = xn * xn + yn * yn;
    const S radial = S(1) + in_[4] * r2 + in_[5] * r2 * r2 + in_[8] * r2 * r2 * r2;
    const S xd = xn * radial + S(2) * in_[6] * xn * yn + in_[7] * (r2 + S(2) * xn * xn);
    const S yd = yn * radial + in_[6] * (r2 + S(2) * yn * yn) + S(2) * in_[7] * xn * yn;
    r[0] = pix_[2 * size_t(i) + 0] - (in_[0] * xd + in_[2]);
    r[1] = pix_[2 * size_t(i) + 1] - (in_[1] * yd + in_[3]);
    return true;
  }
  bool f_df(const S*, S*, S*, unsigned) const override { return false; }
  bool has_jacobian() const override { return false; }
  std::shared_ptr<IModel<S>> clone() const override { return std::make_shared<PinholeDistort>(*this); }
  const S* pts_;
  const S* pix_;
  S C_[16], TC_[12], in_[9];
};

// ---------------------------------------------------------- linearization ----
enum JacobianMode { JAC_ANALYTICAL = 0, JAC_FORWARD = 1, JAC_CENTRAL = 2 };

// include/moptimizer/linearization.h:12-167.  H is P x P (symmetric, so the reference's
// column-major Map and a row-major view coincide), C is O x O column-major.
template <class S>
struct CostComputation {
  CostComputation(int P, int O) : P_(P), O_(O), r_(O), rp_(O), J_(size_t(O) * P) {}

  // linearization.h:36-47
  S computeCost(const S* x, IModel<S>& m, int n) {
    m.setup(x);
    S sum = 0;
    for (int i = 0; i < n; ++i)
      if (m.f(x, r_.data(), i)) sum += dot(r_.data(), r_.data(), O_);
    return sum;
  }

  // linearization.h:49-63, restated WITHOUT its two defects: per-thread
  // residual buffers instead of the shared `residuals_`, and partial sums carried in S.
  // `float_carry=true` re-introduces the oneTBB `0.0f` identity quirk (partials and joins
  // carried as float) for the hard-part-5 experiment; chunking is fixed => deterministic.
  S parallelComputeCost(const S* x, IModel<S>& m, int n, int nthreads, bool float_carry = false) {
    m.setup(x);
    if (nthreads < 1) nthreads = 1;
    std::vector<double> part(nthreads, 0.0);
    auto body = [&](int t) {
      const int64_t lo = int64_t(n) * t / nthreads, hi = int64_t(n) * (t + 1) / nthreads;
      std::vector<S> r(O_);
      if (float_carry) {
        float acc = 0.0f;
        for (int64_t i = lo; i < hi; ++i)
          if (m.f(x, r.data(), unsigned(i))) acc = float(S(acc) + dot(r.data(), r.data(), O_));
        part[t] = acc;
      } else {
        S acc = 0;
        for (int64_t i = lo; i < hi; ++i)
          if (m.f(x, r.data(), unsigned(i))) acc += dot(r.data(), r.data(), O_);
        part[t] = double(acc);
      }
    };
    run_threads(nthreads, body);
    if (float_carry) {
      float s = 0.0f;
      for (int t = 0; t < nthreads; ++t) s = s + float(part[t]);
      return S(s);
    }
    S s = 0;
    for (int t = 0; t < nthreads; ++t) s += S(part[t]);
    return s;
  }

  // linearization.h:126-158 — analytical.
  S computeHessian(const S* x, const S* C, ILoss<S>& loss, S* H, S* b, IModel<S>& m, int n) {
    m.setup(x);
    zero(H, b);
    S sum = 0;
    for (int i = 0; i < n; ++i)
      if (m.f_df(x, r_.data(), J_.data(), i)) accumulate(C, loss, H, b, sum);
    return sum;
  }

  // linearization.h:65-124 — numerical.  `mode` JAC_FORWARD is the reference (step
  // sqrt(eps)*|x_j|, or sqrt(eps) when that is 0; one-sided difference, :78,85-87,105);
  // JAC_CENTRAL is the north-star addition: same step, (r(x+h) - r(x-h)) / 2h.
  S computeHessianNumerical(const S* x, const S* C, ILoss<S>& loss, S* H, S* b, IModel<S>& m,
                            int n, int mode = JAC_FORWARD) {
    const S min_step = std::sqrt(std::numeric_limits<S>::epsilon());
    std::vector<S> h(P_);
    std::vector<std::vector<S>> xp(P_, std::vector<S>(x, x + P_)), xm;
    std::vector<std::shared_ptr<IModel<S>>> mp(P_), mm;
    if (mode == JAC_CENTRAL) {
      xm.assign(P_, std::vector<S>(x, x + P_));
      mm.resize(P_);
    }
    (void)min_step;
    for (int j = 0; j < P_; ++j) {
      perturbed(x, j, h[j], xp[j], mode == JAC_CENTRAL ? &xm[j] : nullptr);
      mp[j] = m.clone();
      mp[j]->setup(xp[j].data());
      if (mode == JAC_CENTRAL) {
        mm[j] = m.clone();
        mm[j]->setup(xm[j].data());
      }
    }
    m.setup(x);
    zero(H, b);
    S sum = 0;
    std::vector<S> rm(O_);
    for (int i = 0; i < n; ++i) {
      if (!m.f(x, r_.data(), i)) continue;
      for (int j = 0; j < P_; ++j) {
        mp[j]->f(xp[j].data(), rp_.data(), i);  // returned bool ignored (:104)
        if (mode == JAC_CENTRAL) {
          mm[j]->f(xm[j].data(), rm.data(), i);
          for (int o = 0; o < O_; ++o) J_[size_t(o) * P_ + j] = (rp_[o] - rm[o]) / (S(2) * h[j]);
        } else {
          for (int o = 0; o < O_; ++o) J_[size_t(o) * P_ + j] = (rp_[o] - r_[o]) / h[j];
        }
      }
      accumulate(C, loss, H, b, sum);
    }
    return sum;
  }

  // Threaded linearization with thread-local H/b — the "generous upper bound" CPU baseline the
  // reference does not have.  Fixed chunking => deterministic.
  S parallelLinearize(const S* x, const S* C, ILoss<S>& loss, S* H, S* b, IModel<S>& m, int n,
                      int mode, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    m.setup(x);
    // Perturbed models are built once (same construction as computeHessianNumerical).
    const S min_step = std::sqrt(std::numeric_limits<S>::epsilon());
    std::vector<S> h(P_);
    std::vector<std::vector<S>> xp(P_, std::vector<S>(x, x + P_)), xm(P_, std::vector<S>(x, x + P_));
    std::vector<std::shared_ptr<IModel<S>>> mp(P_), mm(P_);
    if (mode != JAC_ANALYTICAL)
      for (int j = 0; j < P_; ++j) {
        (void)min_step;
        perturbed(x, j, h[j], xp[j], &xm[j]);
        mp[j] = m.clone();
        mp[j]->setup(xp[j].data());
        mm[j] = m.clone();
        mm[j]->setup(xm[j].data());
      }
    std::vector<std::vector<S>> Hs(nthreads, std::vector<S>(size_t(P_) * P_, 0)),
        bs(nthreads, std::vector<S>(P_, 0));
    std::vector<S> sums(nthreads, 0);
    auto body = [&](int t) {
      CostComputation<S> cc(P_, O_);
      std::vector<S> rm(O_);
      const int64_t lo = int64_t(n) * t / nthreads, hi = int64_t(n) * (t + 1) / nthreads;
      S sum = 0;
      for (int64_t i = lo; i < hi; ++i) {
        if (mode == JAC_ANALYTICAL) {
          if (!m.f_df(x, cc.r_.data(), cc.J_.data(), unsigned(i))) continue;
        } else {
          if (!m.f(x, cc.r_.data(), unsigned(i))) continue;
          for (int j = 0; j < P_; ++j) {
            mp[j]->f(xp[j].data(), cc.rp_.data(), unsigned(i));
            if (mode == JAC_CENTRAL) {
              mm[j]->f(xm[j].data(), rm.data(), unsigned(i));
              for (int o = 0; o < O_; ++o)
                cc.J_[size_t(o) * P_ + j] = (cc.rp_[o] - rm[o]) / (S(2) * h[j]);
            } else {
              for (int o = 0; o < O_; ++o)
                cc.J_[size_t(o) * P_ + j] = (cc.rp_[o] - cc.r_[o]) / h[j];
            }
          }
        }
        cc.accumulate(C, loss, Hs[t].data(), bs[t].data(), sum);
      }
      sums[t] = sum;
    };
    run_threads(nthreads, body);
    zero(H, b);
    S sum = 0;
    for (int t = 0; t < nthreads; ++t) {
      for (int k = 0; k < P_ * P_; ++k) H[k] += Hs[t][k];
      for (int k = 0; k < P_; ++k) b[k] += bs[t][k];
      sum += sums[t];
    }
    return sum;
  }

  // H += w J^T C J ; b += w J^T C r ; sum += r^T r   (linearization.h:108-115,145-152)
  inline void accumulate(const S* C, ILoss<S>& loss, S* H, S* b, S& sum) {
    const S e2 = dot(r_.data(), r_.data(), O_);
    const S w = loss.weight(e2);
    // CJ = C * J (O x P), Cr = C * r (O)
    S CJ[kMaxO * kMaxP], Cr[kMaxO];
    for (int o = 0; o < O_; ++o) {
      for (int p = 0; p < P_; ++p) {
        S s = 0;
        for (int k = 0; k < O_; ++k) s += C[o + size_t(k) * O_] * J_[size_t(k) * P_ + p];
        CJ[o * P_ + p] = s;
      }
      S s = 0;
      for (int k = 0; k < O_; ++k) s += C[o + size_t(k) * O_] * r_[k];
      Cr[o] = s;
    }
    for (int a = 0; a < P_; ++a) {
      for (int c = 0; c < P_; ++c) {
        S s = 0;
        for (int o = 0; o < O_; ++o) s += J_[size_t(o) * P_ + a] * CJ[o * P_ + c];
        H[a + size_t(c) * P_] += w * s;
      }
      S s = 0;
      for (int o = 0; o < O_; ++o) s += J_[size_t(o) * P_ + a] * Cr[o];
      b[a] += w * s;
    }
    sum += e2;
  }

  // Step and perturbed copies of x along parameter j: h_j = sqrt(eps) |x_j| (sqrt(eps) if that is 0),
  // linearization.h:78,85-89.  With MANIFOLD_SO3_LEFT the rotation block is perturbed on the manifold and
  // its tangent coordinate is 0, hence h_j = sqrt(eps).
  void perturbed(const S* x, int j, S& h, std::vector<S>& plus, std::vector<S>* minus) const {
    const S min_step = std::sqrt(std::numeric_limits<S>::epsilon());
    h = min_step * std::fabs(x[j]);  // see SURVEY.md §3.3 hazard: must be fabs
    if (h == S(0)) h = min_step;
    if (manifold_ == MANIFOLD_SO3_LEFT && P_ >= 6 && j >= 3 && j < 6) {
      h = min_step;
      std::vector<S> d(P_, S(0));
      d[j] = h;
      retract(manifold_, P_, x, d.data(), plus.data());
      if (minus) {
        d[j] = -h;
        retract(manifold_, P_, x, d.data(), minus->data());
      }
    } else {
      plus[j] += h;
      if (minus) (*minus)[j] -= h;
    }
  }

  static constexpr int kMaxP = 16, kMaxO = 4;
  int manifold_ = MANIFOLD_ADDITIVE;
  int P_, O_;
  std::vector<S> r_, rp_, J_;

 private:
  static inline S dot(const S* a, const S* b, int n) {
    S s = 0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
  }
  inline void zero(S* H, S* b) const {
    for (int k = 0; k < P_ * P_; ++k) H[k] = 0;
    for (int k = 0; k < P_; ++k) b[k] = 0;
  }
  template <class F>
  static void run_threads(int nthreads, F&& body) {
    if (nthreads == 1) {
      body(0);
      return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(body, t);
    for (auto& t : th) t.join();
  }
};

// ------------------------------------------------------------------ LDLT ----
// Eigen 3.4.0 LDLT (Eigen/src/Cholesky/LDLT.h): in-place L D L^T of the LOWER triangle with
// symmetric diagonal pivoting (largest |a_kk| first), left-looking column update, and the
// solve  x = P^T L^-T D^+ L^-1 P b  with D^+ the pseudo-inverse at tolerance
// numeric_limits::min().  A is column-major n x n, only its lower triangle is read.
template <class S>
inline void ldlt_solve(int n, const S* A_in, const S* rhs, S* out) {
  std::vector<S> A(A_in, A_in + size_t(n) * n);
  std::vector<int> tr(n);
  std::vector<S> tmp(n);
  auto a = [&](int r, int c) -> S& { return A[r + size_t(c) * n]; };
  for (int k = 0; k < n; ++k) {
    int piv = k;
    S best = std::fabs(a(k, k));
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(a(i, i)) > best) {
        best = std::fabs(a(i, i));
        piv = i;
      }
    tr[k] = piv;
    if (piv != k) {
      for (int c = 0; c < k; ++c) std::swap(a(k, c), a(piv, c));
      for (int r = piv + 1; r < n; ++r) std::swap(a(r, k), a(r, piv));
      std::swap(a(k, k), a(piv, piv));
      for (int i = k + 1; i < piv; ++i) std::swap(a(i, k), a(piv, i));
    }
    if (k > 0) {
      for (int c = 0; c < k; ++c) tmp[c] = a(c, c) * a(k, c);
      S s = 0;
      for (int c = 0; c < k; ++c) s += a(k, c) * tmp[c];
      a(k, k) -= s;
      for (int r = k + 1; r < n; ++r) {
        S t = 0;
        for (int c = 0; c < k; ++c) t += a(r, c) * tmp[c];
        a(r, k) -= t;
      }
    }
    const S akk = a(k, k);
    const bool valid = std::fabs(akk) > S(0);
    if (k == 0 && !valid) {
      for (int j = 0; j < n; ++j) tr[j] = j;
      break;
    }
    if (valid)
      for (int r = k + 1; r < n; ++r) a(r, k) /= akk;
  }
  std::vector<S> y(rhs, rhs + n);
  for (int k = 0; k < n; ++k) std::swap(y[k], y[tr[k]]);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < r; ++c) y[r] -= a(r, c) * y[c];
  const S tol = std::numeric_limits<S>::min();
  for (int i = 0; i < n; ++i) y[i] = (std::fabs(a(i, i)) > tol) ? y[i] / a(i, i) : S(0);
  for (int r = n - 1; r >= 0; --r)
    for (int c = r + 1; c < n; ++c) y[r] -= a(c, r) * y[c];
  for (int k = n - 1; k >= 0; --k) std::swap(y[k], y[tr[k]]);
  for (int i = 0; i < n; ++i) out[i] = y[i];
}

// -------------------------------------------------------------------- LM ----
template <class S>
struct Cost {  // include/moptimizer/cost_function.h:15-59 + the *_dyn wrappers
  std::shared_ptr<IModel<S>> model;
  std::shared_ptr<ILoss<S>> loss;
  std::vector<S> C;  // O x O, identity by default (src/cost_function_*_dyn.cpp:14-15)
  int P, O, n;
  int jac_mode;      // JAC_ANALYTICAL => computeHessian, else computeHessianNumerical
  int cost_threads = 1;
  bool float_carry = false;
  int manifold = MANIFOLD_ADDITIVE;
  int lin_threads = 1;  // > 1: thread-local H/b over fixed chunks (not in the reference; lets a 1e7-residual LM
                        // comparison finish in seconds — the sums differ from the serial loop by rounding only)
  S linearize(const S* x, S* H, S* b) {
    CostComputation<S> cc(P, O);
    cc.manifold_ = manifold;
    if (lin_threads > 1) return cc.parallelLinearize(x, C.data(), *loss, H, b, *model, n, jac_mode, lin_threads);
    if (jac_mode == JAC_ANALYTICAL) return cc.computeHessian(x, C.data(), *loss, H, b, *model, n);
    return cc.computeHessianNumerical(x, C.data(), *loss, H, b, *model, n, jac_mode);
  }
  S computeCost(const S* x) {
    CostComputation<S> cc(P, O);
    return cc.parallelComputeCost(x, *model, n, cost_threads, float_carry);
  }
};

struct TraceEntry {
  int outer_it, k;
  double y0, yi, rho, lambda, nu;
  int accepted;
};

// include/moptimizer/delta.h:11-16
template <class S>
inline bool is_delta_small(const S* d, int n) {
  S m = 0;
  for (int i = 0; i < n; ++i) m = std::max(m, S(std::fabs(d[i])));
  return m < std::sqrt(std::numeric_limits<S>::epsilon());
}
// include/moptimizer/optimizer.h:26-29
template <class S>
inline bool is_cost_small(S c) {
  return std::fabs(c) < 8 * std::numeric_limits<S>::epsilon();
}

// src/levenberg_marquadt_dyn.cpp:15-26,34-119
template <class S>
inline Status lm_minimize(std::vector<Cost<S>*>& costs, int P, int max_iterations,
                          int lm_max_iterations, S* x0, int* executed_iterations,
                          std::vector<TraceEntry>* trace, bool stagnation_stop = false) {
  S lambda = S(-1.0);
  const S lambda_factor = S(1e-9);
  std::vector<S> H(size_t(P) * P), Hc(size_t(P) * P), b(P), bc(P), A(size_t(P) * P), nb(P),
      delta(P), xi(P);
  int it = 0;
  *executed_iterations = 0;
  for (it = 0; it < max_iterations; ++it) {
    *executed_iterations = it;
    S y0 = 0;
    std::fill(H.begin(), H.end(), S(0));
    std::fill(b.begin(), b.end(), S(0));
    for (auto* c : costs) {
      c->model->update(x0);
      y0 += c->linearize(x0, Hc.data(), bc.data());
      for (size_t k = 0; k < H.size(); ++k) H[k] += Hc[k];
      for (int k = 0; k < P; ++k) b[k] += bc[k];
    }
    if (is_cost_small(y0)) return CONVERGED;
    if (lambda < S(0)) {
      S mx = 0;
      for (int k = 0; k < P; ++k) mx = std::max(mx, S(std::fabs(H[k + size_t(k) * P])));
      lambda = lambda_factor * mx;
    }
    S nu = S(2.0);
    for (int k = 0; k < lm_max_iterations; ++k) {
      A = H;
      for (int d = 0; d < P; ++d) A[d + size_t(d) * P] += lambda * H[d + size_t(d) * P];
      for (int d = 0; d < P; ++d) nb[d] = -b[d];
      ldlt_solve(P, A.data(), nb.data(), delta.data());
      // additive, no manifold (:82-83) unless the opt-in MANIFOLD_SO3_LEFT is set on the costs
      retract(costs[0]->manifold, P, x0, delta.data(), xi.data());
      S yi = 0;
      for (auto* c : costs) yi += c->computeCost(xi.data());
      if (std::isnan(yi)) return NUMERIC_ERROR;
      S den = 0;
      for (int d = 0; d < P; ++d) den += delta[d] * (lambda * delta[d] - b[d]);
      const S rho = (y0 - yi) / den;
      if (trace) trace->push_back({it, k, double(y0), double(yi), double(rho), double(lambda),
                                   double(nu), !(rho < 0)});
      if (rho < 0) {
        if (is_delta_small(delta.data(), P)) return is_cost_small(yi) ? CONVERGED : SMALL_DELTA;
        lambda = nu * lambda;
        nu = 2 * nu;
        continue;
      }
      for (int d = 0; d < P; ++d) x0[d] = xi[d];
      // NOT in the reference (restates the product's MOPT_LM_STAGNATION_STOP, include/mopt_capi.h): an accepted
      // step below isDeltaSmall's threshold that leaves the cost bit-for-bit unchanged ends the run
      if (stagnation_stop && yi == y0 && is_delta_small(delta.data(), P)) {
        *executed_iterations = it;
        return is_cost_small(yi) ? CONVERGED : SMALL_DELTA;
      }
      // std::max(1.0/3.0, 1 - std::pow(2*rho-1, 3)) is evaluated in double (:113)
      lambda = S(double(lambda) *
                 std::max(1.0 / 3.0, 1.0 - std::pow(double(2 * rho - 1), 3)));
      break;
    }
  }
  *executed_iterations = it;
  return MAXIMUM_ITERATIONS_REACHED;
}

}  // namespace oracle
