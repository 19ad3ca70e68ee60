// Mirrors include/moptimizer/cost_function_analytical.h:16-46: compile-time (P, O) flavour of the analytical
// cost function.  Same device pass as the dynamic class (the kernels are specialised per builtin model).
#pragma once

#include "moptimizer/cost_function_analytical_dyn.h"

namespace moptimizer {

template <class Scalar = double, int model_parameter_dim = 1, int model_output_dim = 1>
class CostFunctionAnalytical : public CostFunctionAnalyticalDynamic<Scalar> {
 public:
  using typename CostFunctionBase<Scalar>::ModelPtr;
  CostFunctionAnalytical(ModelPtr model, int num_residuals)
      : CostFunctionAnalyticalDynamic<Scalar>(model, model_parameter_dim, model_output_dim, num_residuals) {}
  CostFunctionAnalytical(const CostFunctionAnalytical&) = delete;
  CostFunctionAnalytical& operator=(const CostFunctionAnalytical&) = delete;
};

}  // namespace moptimizer
