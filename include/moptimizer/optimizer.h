// Mirrors include/moptimizer/optimizer.h:12-89 (same members, defaults, exceptions and ownership rules).
#pragma once

#include <cmath>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <vector>

#include "moptimizer/cost_function.h"
#include "moptimizer/logger.h"

namespace moptimizer {

template <class Scalar = double>
class Optimizer {
 public:
  using Ptr = std::shared_ptr<Optimizer>;
  using ConstPtr = std::shared_ptr<const Optimizer>;
  using CostFunctionType = CostFunctionBase<Scalar>;

  Optimizer() : maximum_iterations_(15), executed_iterations_(0) {
    logger_.reset(new duna::Logger(std::cout, duna::Logger::L_ERROR, "Optimizer"));
  }
  Optimizer(const Optimizer&) = delete;
  Optimizer& operator=(const Optimizer&) = delete;
  virtual ~Optimizer() = default;

  bool isCostSmall(Scalar cost_sum) { return std::abs(cost_sum) < 8 * (std::numeric_limits<Scalar>::epsilon()); }

  inline void setMaximumIterations(int max_iterations) {
    if (max_iterations < 0) throw std::invalid_argument("Optimization::max_iterations cannot be less than 0.");
    maximum_iterations_ = max_iterations;
  }
  inline unsigned int getMaximumIterations() const { return maximum_iterations_; }
  inline unsigned int getExecutedIterations() const { return executed_iterations_; }

  inline bool checkCosts() const {
    if (costs_.size() == 0) {
      std::cerr << "No cost function added!\n";
      throw std::runtime_error("No cost function added!");
    }
    return true;
  }
  inline void addCost(CostFunctionType* cost) { costs_.push_back(cost); }
  inline void clearCosts(bool delete_costs = false) {
    if (delete_costs)
      for (size_t i = 0; i < costs_.size(); ++i) delete costs_[i];
    costs_.clear();
  }

  virtual OptimizationStatus step(Scalar* x0) = 0;
  virtual OptimizationStatus minimize(Scalar* x0) = 0;

 protected:
  virtual bool hasConverged() = 0;
  virtual void prepare(Scalar* x0) = 0;
  std::vector<CostFunctionType*> costs_;
  unsigned int maximum_iterations_;
  unsigned int executed_iterations_;
  std::shared_ptr<duna::Logger> logger_;
};

}  // namespace moptimizer
