// Mirrors include/moptimizer/cost_function_numerical_dyn.h:13-34 + src/cost_function_numerical_dyn.cpp:7-33.
// linearize -> computeHessianNumerical (linearization.h:65-124: forward difference, step sqrt(eps)|x_j|).
// New: setCentralDifferences(true) selects (r(x+h) - r(x-h)) / 2h with the same steps (north-star addition).
#pragma once

#include "moptimizer/cost_function.h"

namespace moptimizer {

template <class Scalar = double>
class CostFunctionNumericalDynamic : public CostFunctionBase<Scalar> {
 public:
  using typename CostFunctionBase<Scalar>::Model;
  using typename CostFunctionBase<Scalar>::ModelPtr;

  CostFunctionNumericalDynamic(ModelPtr model, int num_parameters, int num_outputs, int num_residuals)
      : CostFunctionBase<Scalar>(model, num_residuals), num_parameters_(num_parameters), num_outputs_(num_outputs) {
    covariance_->resize(num_outputs_, num_outputs_);  // src/cost_function_numerical_dyn.cpp:14-15
    covariance_->setIdentity();
  }
  ~CostFunctionNumericalDynamic() override = default;

  void setCentralDifferences(bool on) { central_ = on; }

  Scalar computeCost(const Scalar* x) override {
    mopt_problem p;
    device::Store::Ptr st;
    deviceProblem(&p, &st);
    return detail::deviceCost<Scalar>(p, st, x);
  }
  Scalar linearize(const Scalar* x, Scalar* hessian, Scalar* b) override {
    mopt_problem p;
    device::Store::Ptr st;
    deviceProblem(&p, &st);
    return detail::deviceLinearize<Scalar>(p, st, x, hessian, b);
  }
  void deviceProblem(mopt_problem* p, device::Store::Ptr* st) const override {
    this->fillDeviceProblem(num_parameters_, num_outputs_, central_ ? MOPT_JAC_CENTRAL : MOPT_JAC_FORWARD, p, st);
  }

 protected:
  using CostFunctionBase<Scalar>::num_residuals_;
  using CostFunctionBase<Scalar>::model_;
  using CostFunctionBase<Scalar>::loss_function_;
  using CostFunctionBase<Scalar>::covariance_;
  int num_parameters_;
  int num_outputs_;
  bool central_ = false;
};

}  // namespace moptimizer
