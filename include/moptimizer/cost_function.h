// Mirrors include/moptimizer/cost_function.h:15-59 (same operator API, ownership and defaults).  New: a
// device hook so the optimizer can run every pass and its own state machine on the GPU.
#pragma once

#include <cstring>
#include <memory>

#include "mopt_capi.h"
#include "moptimizer/covariance/covariance.h"
#include "moptimizer/device/models.h"
#include "moptimizer/loss_function/geman_mcclure.h"
#include "moptimizer/loss_function/huber.h"
#include "moptimizer/loss_function/loss_function.h"
#include "moptimizer/model.h"
#include "moptimizer/types.h"

namespace moptimizer {

template <class Scalar = double>
class CostFunctionBase {
 public:
  using Model = IBaseModel<Scalar>;
  using ModelPtr = typename Model::Ptr;
  using ModelConstPtr = typename Model::ConstPtr;
  using LossFunctionPtr = typename loss::ILossFunction<Scalar>::Ptr;

  CostFunctionBase(ModelPtr model, int num_residuals) : num_residuals_(num_residuals), model_(model) {
    loss_function_.reset(new loss::NoLoss<Scalar>());
    covariance_.reset(new covariance::Matrix<Scalar>());
  }
  CostFunctionBase() = delete;
  CostFunctionBase(const CostFunctionBase&) = delete;
  CostFunctionBase& operator=(const CostFunctionBase&) = delete;
  virtual ~CostFunctionBase() = default;

  inline void setLossFunction(LossFunctionPtr loss_function) { loss_function_ = loss_function; }
  inline void setCovariance(const covariance::MatrixPtr<Scalar> covariance) { covariance_ = covariance; }
  /// New (opt-in): MOPT_MANIFOLD_SO3_LEFT makes the optimizer update the rotation block of x
  /// with so3::Exp/Log instead of adding delta (the reference's "TODO Manifold operation"); Jacobians are then
  /// taken in the tangent space (use the MOPT_P2P_LEFT variant for analytical point2point).
  inline void setManifold(int manifold) { manifold_ = manifold; }
  /// New (opt-in): mopt_problem_flags for the device pass, e.g. MOPT_FLAG_STABLE_FD (double: finite differences of
  /// the camera models over a common denominator) or MOPT_FLAG_GENERIC_KERNEL (always the per-residual quotient of
  /// linearization.h:97-111).  0 by default.
  inline void setKernelFlags(int flags) { kernel_flags_ = flags; }

  virtual void update(const Scalar* x) { model_->update(x); }
  virtual Scalar computeCost(const Scalar* x) = 0;
  virtual Scalar linearize(const Scalar* x, Scalar* hessian, Scalar* b) = 0;

  /// Device description of this cost term (C-ABI problem + store).  Throws moptimizer::Exception if the model
  /// is not a builtin device model or the loss is not one the kernels implement: there is no host fallback.
  virtual void deviceProblem(mopt_problem* problem, device::Store::Ptr* store) const = 0;

 protected:
  void fillDeviceProblem(int num_parameters, int num_outputs, int jacobian, mopt_problem* p,
                         device::Store::Ptr* store) const {
    const auto* dm = dynamic_cast<const device::IDeviceModel*>(model_.get());
    if (!dm)
      throw Exception("cost function: the model is not a device model (moptimizer/device/models.h); user-defined "
                      "host models cannot run on the GPU path and there is no CPU fallback");
    std::memset(p, 0, sizeof(*p));
    p->model = dm->kind();
    p->variant = dm->variant();
    p->num_parameters = num_parameters;
    p->num_outputs = num_outputs;
    p->jacobian = jacobian;
    p->compute_dtype = device::dtypeOf<Scalar>();
    p->manifold = manifold_;
    p->flags = kernel_flags_;
    int loss_kind = MOPT_LOSS_NONE;
    double loss_parameter = 0.0;
    if (!loss_function_ || !loss_function_->deviceLoss(&loss_kind, &loss_parameter))
      throw Exception("cost function: this loss has no device implementation (ILossFunction::deviceLoss); NoLoss, "
                      "GemmanMCClure and Huber are implemented by the kernels");
    p->loss = loss_kind;
    p->loss_param = loss_parameter;
    const int O = num_outputs;
    if (covariance_ && covariance_->rows() == O && covariance_->cols() == O) {
      bool identity = true;
      for (int r = 0; r < O; ++r)
        for (int c = 0; c < O; ++c) {
          const double v = double((*covariance_)(r, c));
          p->covariance[r + c * O] = v;
          if (v != (r == c ? 1.0 : 0.0)) identity = false;
        }
      p->has_covariance = identity ? 0 : 1;
    } else if (covariance_ && covariance_->size() != 0) {
      throw Exception("cost function: covariance must be num_outputs x num_outputs");
    }
    dm->fillConsts(p->consts);
    *store = dm->store();
  }

  int num_residuals_;
  int manifold_ = MOPT_MANIFOLD_ADDITIVE;
  int kernel_flags_ = 0;
  ModelPtr model_;
  LossFunctionPtr loss_function_;
  covariance::MatrixPtr<Scalar> covariance_;
};

namespace detail {
// Shared body of the four cost-function classes: one device pass through the C ABI.
template <class Scalar>
inline Scalar deviceLinearize(const mopt_problem& p, const device::Store::Ptr& st, const Scalar* x, Scalar* hessian,
                              Scalar* b) {
  const int P = p.num_parameters;
  double xd[MOPT_MAX_PARAMETERS] = {0}, Hd[MOPT_MAX_PARAMETERS * MOPT_MAX_PARAMETERS], bd[MOPT_MAX_PARAMETERS], sum = 0;
  for (int i = 0; i < P; ++i) xd[i] = double(x[i]);
  device::check(mopt_linearize(st->context()->get(), st->get(), &p, xd, Hd, bd, &sum), "mopt_linearize");
  for (int i = 0; i < P * P; ++i) hessian[i] = Scalar(Hd[i]);
  for (int i = 0; i < P; ++i) b[i] = Scalar(bd[i]);
  return Scalar(sum);
}
template <class Scalar>
inline Scalar deviceCost(const mopt_problem& p, const device::Store::Ptr& st, const Scalar* x) {
  double xd[MOPT_MAX_PARAMETERS] = {0}, sum = 0;
  for (int i = 0; i < p.num_parameters && x; ++i) xd[i] = double(x[i]);
  device::check(mopt_compute_cost(st->context()->get(), st->get(), &p, xd, &sum), "mopt_compute_cost");
  return Scalar(sum);
}
}  // namespace detail

}  // namespace moptimizer
