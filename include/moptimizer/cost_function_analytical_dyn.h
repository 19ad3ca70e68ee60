// Mirrors include/moptimizer/cost_function_analytical_dyn.h:12-32 + src/cost_function_analytical_dyn.cpp:7-33.
// linearize -> computeHessian (linearization.h:126-158), computeCost -> parallelComputeCost (:49-63), both as
// one device pass through the C ABI.
#pragma once

#include "moptimizer/cost_function.h"

namespace moptimizer {

template <class Scalar = double>
class CostFunctionAnalyticalDynamic : public CostFunctionBase<Scalar> {
 public:
  using typename CostFunctionBase<Scalar>::Model;
  using typename CostFunctionBase<Scalar>::ModelPtr;

  CostFunctionAnalyticalDynamic(ModelPtr model, int num_parameters, int num_outputs, int num_residuals)
      : CostFunctionBase<Scalar>(model, num_residuals), num_parameters_(num_parameters), num_outputs_(num_outputs) {
    covariance_->resize(num_outputs_, num_outputs_);  // src/cost_function_analytical_dyn.cpp:14-15
    covariance_->setIdentity();
  }
  ~CostFunctionAnalyticalDynamic() override = default;

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
    this->fillDeviceProblem(num_parameters_, num_outputs_, MOPT_JAC_ANALYTICAL, p, st);
  }

 protected:
  using CostFunctionBase<Scalar>::num_residuals_;
  using CostFunctionBase<Scalar>::model_;
  using CostFunctionBase<Scalar>::loss_function_;
  using CostFunctionBase<Scalar>::covariance_;
  int num_parameters_;
  int num_outputs_;
};

}  // namespace moptimizer
