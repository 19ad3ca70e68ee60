// Builtin device models.  Each one IS an IBaseModel<Scalar> (so it plugs into the reference's cost-function
// constructors unchanged) and names the kernel-side model the pass kernels run (mopt_model in mopt_capi.h).
// They restate, for the device, the models the reference defines in its tests:
//   Point2Point       tst/point2point.cpp:24-84        ExpCurve          tst/curve_fitting.cpp:81-98
//   MichaelisMenten   tst/test_models.h:8-19           PinholeCamera     tst/camera_calibration.cpp:12-57
//   Powell            tst/powell.cpp:22-59             Point2PointDist   tst/parallel.cpp:12-32
// f / f_df are host entry points of the reference interface; residuals of these models are only ever
// evaluated on the device, so calling them throws.
#pragma once

#include <cstring>
#include <stdexcept>
#include <vector>

#include "moptimizer/device/context.h"
#include "moptimizer/model.h"

namespace moptimizer::device {

/// What a cost function needs to know about a device-backed model.
class IDeviceModel {
 public:
  virtual ~IDeviceModel() = default;
  virtual int kind() const = 0;               // mopt_model
  virtual int variant() const { return 0; }   // mopt_p2p_variant
  virtual const Store::Ptr& store() const = 0;
  virtual void fillConsts(double* consts32) const { std::memset(consts32, 0, sizeof(double) * 32); }
};

template <typename Scalar, class Derived>
class DeviceModel : public IBaseModel<Scalar>, public IDeviceModel {
 public:
  void setup(const Scalar*) override {}   // runs on the device (csrc/mopt_setup.cuh) inside every pass
  void update(const Scalar*) override {}  // correspondences are fixed
  bool f(const Scalar*, Scalar*, unsigned int) const override {
    throw moptimizer::Exception("device model: residuals are evaluated on the GPU only (no host `f`)");
  }
  bool f_df(const Scalar*, Scalar*, Scalar*, unsigned int) const override {
    throw moptimizer::Exception("device model: residuals are evaluated on the GPU only (no host `f_df`)");
  }
  std::shared_ptr<IBaseModel<Scalar>> clone() const override {
    return std::make_shared<Derived>(*static_cast<const Derived*>(this));  // shares the device store
  }
  const Store::Ptr& store() const override { return store_; }

 protected:
  Store::Ptr store_;
};

/// r = T(x) p - q, x = [t, omega].  `src`/`tgt` are AoS xyz arrays of HostScalar (e.g. the memory of a
/// std::vector<Eigen::Vector3d>), copied once into a planar device store of `store_dtype`.
template <typename Scalar>
class Point2Point : public DeviceModel<Scalar, Point2Point<Scalar>> {
 public:
  using Ptr = std::shared_ptr<Point2Point>;
  template <class HostScalar>
  Point2Point(Context::Ptr ctx, const HostScalar* src, const HostScalar* tgt, int64_t n,
              int jacobian_variant = MOPT_P2P_EXACT, int store_dtype = dtypeOf<Scalar>())
      : variant_(jacobian_variant) {
    this->store_ = std::make_shared<Store>(std::move(ctx), MOPT_MODEL_POINT2POINT, store_dtype, n);
    this->store_->upload(0, src, n);
    this->store_->upload(1, tgt, n);
  }
  /// Adopt an existing (e.g. device-generated or sharded) store.
  explicit Point2Point(Store::Ptr store, int jacobian_variant = MOPT_P2P_EXACT) : variant_(jacobian_variant) {
    this->store_ = std::move(store);
  }
  /// Source cloud only: correspondences come from setTarget() + update(x) (registration with unknown matches).
  template <class HostScalar>
  Point2Point(Context::Ptr ctx, const HostScalar* src, int64_t n, int jacobian_variant = MOPT_P2P_EXACT,
              int store_dtype = dtypeOf<Scalar>())
      : variant_(jacobian_variant) {
    this->store_ = std::make_shared<Store>(std::move(ctx), MOPT_MODEL_POINT2POINT, store_dtype, n);
    this->store_->upload(0, src, n);
  }
  int kind() const override { return MOPT_MODEL_POINT2POINT; }
  int variant() const override { return variant_; }

  /// The fixed cloud update(x) searches: tgt_i <- nearest target point of T(x) src_i within max_distance;
  /// source points without one are skipped (`f` returning false in the reference, linearization.h:102,144).
  template <class HostScalar>
  void setTarget(const HostScalar* target_xyz, int64_t m, double max_distance) {
    index_ = std::make_shared<NNIndex>(this->store_->context(), target_xyz, m, max_distance);
    check(mopt_store_set_target(this->store_->get(), index_->get()), "mopt_store_set_target");
  }
  /// model->update(x) (model.h:24-26): re-associate correspondences on the device.  The optimizer calls it at
  /// the start of every outer iteration (levenberg_marquadt_dyn.cpp:54); on the device path that happens
  /// inside mopt_lm_minimize without a host round trip.
  void update(const Scalar* x) override {
    if (!index_) return;
    double xd[6];
    for (int i = 0; i < 6; ++i) xd[i] = double(x[i]);
    check(mopt_store_reassociate(this->store_->get(), xd, &matched_), "mopt_store_reassociate");
  }
  int64_t matched() const { return matched_; }

 private:
  int variant_;
  NNIndex::Ptr index_;
  int64_t matched_ = 0;
};

/// r = src - tgt (no parameters); cost only.
template <typename Scalar>
class Point2PointDist : public DeviceModel<Scalar, Point2PointDist<Scalar>> {
 public:
  using Ptr = std::shared_ptr<Point2PointDist>;
  template <class HostScalar>
  Point2PointDist(Context::Ptr ctx, const HostScalar* src, const HostScalar* tgt, int64_t n,
                  int store_dtype = dtypeOf<Scalar>()) {
    this->store_ = std::make_shared<Store>(std::move(ctx), MOPT_MODEL_POINT_DIST, store_dtype, n);
    this->store_->upload(0, src, n);
    this->store_->upload(1, tgt, n);
  }
  int kind() const override { return MOPT_MODEL_POINT_DIST; }
};

/// r = y - exp(x0 t + x1).  `dataset` is the reference's interleaved {t0,y0,t1,y1,...} array.
template <typename Scalar>
class ExpCurve : public DeviceModel<Scalar, ExpCurve<Scalar>> {
 public:
  using Ptr = std::shared_ptr<ExpCurve>;
  template <class HostScalar>
  ExpCurve(Context::Ptr ctx, const HostScalar* dataset, int64_t n, int store_dtype = dtypeOf<Scalar>()) {
    this->store_ = std::make_shared<Store>(std::move(ctx), MOPT_MODEL_EXP_CURVE, store_dtype, n);
    this->store_->upload(0, dataset, n, 2);
    this->store_->upload(1, dataset + 1, n, 2);
  }
  explicit ExpCurve(Store::Ptr store) { this->store_ = std::move(store); }
  int kind() const override { return MOPT_MODEL_EXP_CURVE; }
};

/// r = y - x0 t / (x1 + t).
template <typename Scalar>
class MichaelisMenten : public DeviceModel<Scalar, MichaelisMenten<Scalar>> {
 public:
  using Ptr = std::shared_ptr<MichaelisMenten>;
  template <class HostScalar>
  MichaelisMenten(Context::Ptr ctx, const HostScalar* t, const HostScalar* y, int64_t n,
                  int store_dtype = dtypeOf<Scalar>()) {
    this->store_ = std::make_shared<Store>(std::move(ctx), MOPT_MODEL_MICHAELIS_MENTEN, store_dtype, n);
    this->store_->upload(0, t, n);
    this->store_->upload(1, y, n);
  }
  int kind() const override { return MOPT_MODEL_MICHAELIS_MENTEN; }
};

/// r = pixel - proj(K T(x) C P).  `points` AoS with `point_stride` scalars per point (4 for Vector4d),
/// `pixels` AoS uv; K 3x4 and C 4x4 row-major.
template <typename Scalar>
class PinholeCamera : public DeviceModel<Scalar, PinholeCamera<Scalar>> {
 public:
  using Ptr = std::shared_ptr<PinholeCamera>;
  template <class HostScalar>
  PinholeCamera(Context::Ptr ctx, const HostScalar* points, int64_t point_stride, const HostScalar* pixels, int64_t n,
                const double* K34, const double* C44, int store_dtype = dtypeOf<Scalar>()) {
    if (n <= 0) throw std::runtime_error("Empty point list");  // tst/camera_calibration.cpp:16-18
    this->store_ = std::make_shared<Store>(std::move(ctx), MOPT_MODEL_PINHOLE, store_dtype, n);
    this->store_->upload(0, points, n, point_stride);
    this->store_->upload(1, pixels, n);
    consts_.assign(32, 0.0);
    for (int i = 0; i < 12; ++i) consts_[i] = K34[i];
    for (int i = 0; i < 16; ++i) consts_[12 + i] = C44[i];
  }
  int kind() const override { return MOPT_MODEL_PINHOLE; }
  void fillConsts(double* c) const override { std::memcpy(c, consts_.data(), sizeof(double) * 32); }

 private:
  std::vector<double> consts_;
};

/// Pinhole with free intrinsics and OpenCV distortion — the "n x n calibration case" (not in the reference):
/// x = [t, omega, fx, fy, cx, cy, k1, k2, p1, p2, k3], 15 parameters, 2 outputs.  C 4x4 row-major.
template <typename Scalar>
class PinholeDistortCamera : public DeviceModel<Scalar, PinholeDistortCamera<Scalar>> {
 public:
  using Ptr = std::shared_ptr<PinholeDistortCamera>;
  template <class HostScalar>
  PinholeDistortCamera(Context::Ptr ctx, const HostScalar* points, int64_t point_stride, const HostScalar* pixels,
                       int64_t n, const double* C44, int store_dtype = dtypeOf<Scalar>()) {
    if (n <= 0) throw std::runtime_error("Empty point list");
    this->store_ = std::make_shared<Store>(std::move(ctx), MOPT_MODEL_PINHOLE_DISTORT, store_dtype, n);
    this->store_->upload(0, points, n, point_stride);
    this->store_->upload(1, pixels, n);
    consts_.assign(32, 0.0);
    for (int i = 0; i < 16; ++i) consts_[i] = C44[i];
  }
  int kind() const override { return MOPT_MODEL_PINHOLE_DISTORT; }
  void fillConsts(double* c) const override { std::memcpy(c, consts_.data(), sizeof(double) * 32); }

 private:
  std::vector<double> consts_;
};

/// Powell's singular function (no data).
template <typename Scalar>
class Powell : public DeviceModel<Scalar, Powell<Scalar>> {
 public:
  using Ptr = std::shared_ptr<Powell>;
  explicit Powell(Context::Ptr ctx) {
    this->store_ = std::make_shared<Store>(std::move(ctx), MOPT_MODEL_POWELL, dtypeOf<Scalar>(), 1);
  }
  int kind() const override { return MOPT_MODEL_POWELL; }
};

/// A model defined by the user as CUDA C++ source (mopt_capi.h, "user-defined device models"): the device
/// counterpart of deriving from BaseModel / BaseModelJacobian (model.h:50-104).  The source defines
///   template <typename T> __device__ void mopt_f(const T* s, const T* a, const T* b, T* r);
///   template <typename T> __device__ void mopt_f_df(const T* s, const T* a, const T* b, T* r, T* J);   // optional
///   __device__ void mopt_setup(const double* x, const double* consts, double* s);                      // optional
/// and is compiled at run time into the library's pass kernels.  `a` / `b` are AoS arrays with ncomp_a / ncomp_b
/// scalars per residual.
class UserModelSource {
 public:
  using Ptr = std::shared_ptr<UserModelSource>;
  UserModelSource(const std::string& cuda_source, const mopt_user_model_desc& desc) : desc_(desc) {
    check(mopt_user_model_compile(cuda_source.c_str(), &desc, &id_), "mopt_user_model_compile");
  }
  ~UserModelSource() { mopt_user_model_release(id_); }
  UserModelSource(const UserModelSource&) = delete;
  UserModelSource& operator=(const UserModelSource&) = delete;
  int id() const { return id_; }
  const mopt_user_model_desc& desc() const { return desc_; }
  std::string log() const { return mopt_user_model_log(id_); }

 private:
  mopt_user_model_desc desc_;
  int id_ = 0;
};

template <typename Scalar>
class UserModel : public DeviceModel<Scalar, UserModel<Scalar>> {
 public:
  using Ptr = std::shared_ptr<UserModel>;
  template <class HostScalar>
  UserModel(Context::Ptr ctx, UserModelSource::Ptr source, const HostScalar* a, const HostScalar* b, int64_t n,
            int store_dtype = dtypeOf<Scalar>())
      : source_(std::move(source)) {
    this->store_ = std::make_shared<Store>(std::move(ctx), source_->id(), store_dtype, n);
    this->store_->upload(0, a, n);
    if (source_->desc().ncomp_b > 0) this->store_->upload(1, b, n);
    consts_.assign(32, 0.0);
  }
  int kind() const override { return source_->id(); }
  /// Constants handed to the source's mopt_setup (mopt_problem.consts, up to 32 doubles).
  void setConsts(const std::vector<double>& c) {
    for (size_t i = 0; i < c.size() && i < 32; ++i) consts_[i] = c[i];
  }
  void fillConsts(double* c) const override { std::memcpy(c, consts_.data(), sizeof(double) * 32); }

 private:
  UserModelSource::Ptr source_;
  std::vector<double> consts_;
};

}  // namespace moptimizer::device
