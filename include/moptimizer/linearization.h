// Mirrors include/moptimizer/linearization.h:12-167: CostComputation<Scalar, P, O> with the reference's four
// entry points and argument order.  Each call is one fused device pass (csrc/mopt_pass.cuh).
#pragma once

#include <cstring>

#include "mopt_capi.h"
#include "moptimizer/cost_function.h"

namespace moptimizer {

constexpr int Dynamic = -1;  // stands in for Eigen::Dynamic

template <class Scalar, int model_parameter_dim = Dynamic, int model_output_dim = Dynamic>
class CostComputation {
 public:
  using ModelPtr = typename IBaseModel<Scalar>::Ptr;
  using LossPtr = typename loss::ILossFunction<Scalar>::Ptr;

  CostComputation() : actual_parameter_dim_(model_parameter_dim), actual_output_dim_(model_output_dim) {}
  explicit CostComputation(int dyn_model_parameter_dim, int dyn_model_output_dim)
      : actual_parameter_dim_(dyn_model_parameter_dim), actual_output_dim_(dyn_model_output_dim) {}

  /// linearization.h:36-47
  Scalar computeCost(const Scalar* const x, ModelPtr model, int num_elements) { return cost(x, model, num_elements); }
  /// linearization.h:49-63 — on the device every cost evaluation is the parallel one.
  Scalar parallelComputeCost(const Scalar* const x, ModelPtr model, int num_elements) {
    return cost(x, model, num_elements);
  }
  /// linearization.h:65-124
  Scalar computeHessianNumerical(const Scalar* const x, const Scalar* const covariance_data, const LossPtr loss_function,
                                 Scalar* hessian_data, Scalar* b_data, ModelPtr model, int num_elements) {
    return lin(x, covariance_data, loss_function, hessian_data, b_data, model, num_elements, MOPT_JAC_FORWARD);
  }
  /// linearization.h:126-158
  Scalar computeHessian(const Scalar* const x, const Scalar* const covariance_data, const LossPtr loss_function,
                        Scalar* hessian_data, Scalar* b_data, ModelPtr model, int num_elements) {
    return lin(x, covariance_data, loss_function, hessian_data, b_data, model, num_elements, MOPT_JAC_ANALYTICAL);
  }

 private:
  // A throw-away cost function gives us the model/loss/covariance -> mopt_problem translation.
  struct Shim : CostFunctionBase<Scalar> {
    Shim(ModelPtr m, int n, int P, int O, int jac) : CostFunctionBase<Scalar>(m, n), P_(P), O_(O), jac_(jac) {}
    Scalar computeCost(const Scalar*) override { return 0; }
    Scalar linearize(const Scalar*, Scalar*, Scalar*) override { return 0; }
    void deviceProblem(mopt_problem* p, device::Store::Ptr* st) const override {
      this->fillDeviceProblem(P_, O_, jac_, p, st);
    }
    int P_, O_, jac_;
  };
  void checkCount(const device::Store::Ptr& st, int num_elements) const {
    if (num_elements != st->size())
      throw Exception("CostComputation: num_elements must equal the device store size");
  }
  Scalar cost(const Scalar* x, ModelPtr model, int n) {
    Shim shim(model, n, actual_parameter_dim_, actual_output_dim_, MOPT_JAC_FORWARD);
    mopt_problem p;
    device::Store::Ptr st;
    shim.deviceProblem(&p, &st);
    // parameter-free models (tst/parallel.cpp instantiates CostComputation<scalar,3,3> for one) report P = 0
    if (p.model == MOPT_MODEL_POINT_DIST) p.num_parameters = 0;
    checkCount(st, n);
    return detail::deviceCost<Scalar>(p, st, x);
  }
  Scalar lin(const Scalar* x, const Scalar* cov, const LossPtr loss, Scalar* H, Scalar* b, ModelPtr model, int n, int jac) {
    Shim shim(model, n, actual_parameter_dim_, actual_output_dim_, jac);
    if (loss) shim.setLossFunction(loss);
    if (cov) {
      auto c = std::make_shared<covariance::Matrix<Scalar>>();
      c->resize(actual_output_dim_, actual_output_dim_);
      std::memcpy(c->data(), cov, sizeof(Scalar) * size_t(actual_output_dim_) * actual_output_dim_);
      shim.setCovariance(c);
    }
    mopt_problem p;
    device::Store::Ptr st;
    shim.deviceProblem(&p, &st);
    checkCount(st, n);
    return detail::deviceLinearize<Scalar>(p, st, x, H, b);
  }

  int actual_parameter_dim_;
  int actual_output_dim_;
};

}  // namespace moptimizer
