// Mirrors include/moptimizer/cost_function_numerical.h:16-46: compile-time (P, O) flavour of the numerical
// cost function.
#pragma once

#include "moptimizer/cost_function_numerical_dyn.h"

namespace moptimizer {

template <class Scalar = double, int model_parameter_dim = 1, int model_output_dim = 1>
class CostFunctionNumerical : public CostFunctionNumericalDynamic<Scalar> {
 public:
  using typename CostFunctionBase<Scalar>::ModelPtr;
  CostFunctionNumerical(ModelPtr model, int num_residuals)
      : CostFunctionNumericalDynamic<Scalar>(model, model_parameter_dim, model_output_dim, num_residuals) {}
  CostFunctionNumerical(const CostFunctionNumerical&) = delete;
  CostFunctionNumerical& operator=(const CostFunctionNumerical&) = delete;
};

}  // namespace moptimizer
