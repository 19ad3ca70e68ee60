// Huber IRLS weight — NOT in the reference; added for the north-star workload.
// In terms of the API's argument e2 = |r|^2 (loss_function.h:16):  w = 1 if e2 <= k^2 else k / sqrt(e2).
#pragma once

#include <cmath>

#include "loss_function.h"

namespace moptimizer::loss {

template <typename T>
class Huber : public ILossFunction<T> {
 public:
  using Ptr = std::shared_ptr<Huber>;
  explicit Huber(T k) : k_(k) {}
  T weight(T errorSquaredNorm) override {
    return (errorSquaredNorm <= k_ * k_) ? T(1) : k_ / std::sqrt(errorSquaredNorm);
  }
  bool deviceLoss(int* kind, double* parameter) const override {
    *kind = MOPT_LOSS_HUBER;
    *parameter = double(k_);
    return true;
  }
  T k() const { return k_; }

 private:
  T k_;
};

}  // namespace moptimizer::loss
