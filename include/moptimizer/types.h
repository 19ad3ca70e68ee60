// Status codes of an optimization run.  Same enumerator names as include/moptimizer/types.h:6-12 of the reference;
// the numeric values are pinned here because they cross the C ABI (mopt_optimization_status in mopt_capi.h).
#pragma once

#include "mopt_capi.h"

namespace moptimizer {

enum OptimizationStatus {
  CONVERGED = MOPT_CONVERGED,                                    // cost below 8 eps (optimizer.h:26-29)
  MAXIMUM_ITERATIONS_REACHED = MOPT_MAXIMUM_ITERATIONS_REACHED,  // outer loop exhausted
  SMALL_DELTA = MOPT_SMALL_DELTA,                                // a rejected step shorter than sqrt(eps) (delta.h:11-16)
  NUMERIC_ERROR = MOPT_NUMERIC_ERROR,                            // NaN cost at a trial point
  FATAL_ERROR = MOPT_FATAL_ERROR,
};
static_assert(CONVERGED == 0 && MAXIMUM_ITERATIONS_REACHED == 1 && SMALL_DELTA == 2 && NUMERIC_ERROR == 3 && FATAL_ERROR == 4,
              "OptimizationStatus must keep the reference's implicit values");

inline const char* toString(OptimizationStatus s) {
  static const char* const names[] = {"CONVERGED", "MAXIMUM_ITERATIONS_REACHED", "SMALL_DELTA", "NUMERIC_ERROR", "FATAL_ERROR"};
  return (s >= 0 && s <= FATAL_ERROR) ? names[s] : "?";
}

}  // namespace moptimizer
