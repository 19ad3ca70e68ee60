// Mirrors include/moptimizer/loss_function/geman_mcclure.h:7-19 (class name as spelled there).
#pragma once

#include "loss_function.h"

namespace moptimizer::loss {

template <typename T>
class GemmanMCClure : public ILossFunction<T> {
 public:
  using Ptr = std::shared_ptr<GemmanMCClure>;
  explicit GemmanMCClure(T threshold) : threshold_(threshold) {}
  T weight(T errorSquaredNorm) override {
    const T d = errorSquaredNorm + threshold_;
    return (threshold_ * threshold_) / (d * d);
  }
  T threshold() const { return threshold_; }

 private:
  T threshold_;
};

}  // namespace moptimizer::loss
