// Optimizer base of the reference API (include/moptimizer/optimizer.h:12-89): a list of non-owning cost functions, an
// iteration budget (default 15) and the minimize / step contract.  Same public members, defaults, exceptions and
// ownership rules; the device path adds gatherDeviceProblems(), which turns the cost list into what one
// mopt_lm_minimize call needs.
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

  Optimizer() : logger_(std::make_shared<duna::Logger>(std::cout, duna::Logger::L_ERROR, "Optimizer")) {}
  Optimizer(const Optimizer&) = delete;
  Optimizer& operator=(const Optimizer&) = delete;
  virtual ~Optimizer() = default;

  // ---- iteration budget (optimizer.h:19,33-44) -------------------------------------------------------------
  void setMaximumIterations(int max_iterations) {
    if (max_iterations < 0) throw std::invalid_argument("Optimization::max_iterations cannot be less than 0.");
    maximum_iterations_ = static_cast<unsigned int>(max_iterations);
  }
  unsigned int getMaximumIterations() const { return maximum_iterations_; }
  unsigned int getExecutedIterations() const { return executed_iterations_; }

  // ---- cost list: raw, non-owning pointers unless clearCosts(true) (optimizer.h:46-66) ----------------------
  void addCost(CostFunctionType* cost) { costs_.push_back(cost); }
  void clearCosts(bool delete_costs = false) {
    if (delete_costs)
      for (CostFunctionType* c : costs_) delete c;
    costs_.clear();
  }
  bool checkCosts() const {
    if (!costs_.empty()) return true;
    std::cerr << "No cost function added!\n";
    throw std::runtime_error("No cost function added!");
  }

  /// |cost| < 8 eps of Scalar (optimizer.h:26-29); the device state machine applies the same test (csrc/mopt_lm.cuh).
  bool isCostSmall(Scalar cost_sum) { return std::abs(cost_sum) < Scalar(8) * std::numeric_limits<Scalar>::epsilon(); }

  virtual OptimizationStatus step(Scalar* x0) = 0;
  virtual OptimizationStatus minimize(Scalar* x0) = 0;

 protected:
  virtual bool hasConverged() = 0;
  virtual void prepare(Scalar* x0) = 0;

  /// cost->update(x0) for every cost (levenberg_marquadt_dyn.cpp:54), then their device descriptions.  Throws if a
  /// cost is not device-backed, if there are more than MOPT_MAX_COSTS, or if they live on different contexts.
  void gatherDeviceProblems(const Scalar* x0, std::vector<mopt_problem>* problems, std::vector<device::Store::Ptr>* stores,
                            std::vector<mopt_store*>* raw) const {
    const size_t n = costs_.size();
    if (n > size_t(MOPT_MAX_COSTS)) throw Exception("Optimizer: too many cost functions for the device path");
    problems->resize(n);
    stores->resize(n);
    raw->resize(n);
    for (size_t i = 0; i < n; ++i) {
      costs_[i]->update(x0);
      costs_[i]->deviceProblem(&(*problems)[i], &(*stores)[i]);
      (*raw)[i] = (*stores)[i]->get();
      if ((*stores)[i]->context() != (*stores)[0]->context())
        throw Exception("Optimizer: all cost functions must live on the same device context");
    }
  }

  std::vector<CostFunctionType*> costs_;
  unsigned int maximum_iterations_ = 15;
  unsigned int executed_iterations_ = 0;
  std::shared_ptr<duna::Logger> logger_;
};

}  // namespace moptimizer
