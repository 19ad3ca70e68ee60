// Mirrors include/moptimizer/types.h:6-12 of the reference (same enumerators, same values).
#pragma once

namespace moptimizer {

enum OptimizationStatus {
  CONVERGED,
  MAXIMUM_ITERATIONS_REACHED,
  SMALL_DELTA,
  NUMERIC_ERROR,
  FATAL_ERROR,
};

}  // namespace moptimizer
