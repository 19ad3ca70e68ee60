// Geman-McClure weight, class name as the reference spells it (include/moptimizer/loss_function/geman_mcclure.h:7-19):
//   w(e2) = th^2 / (e2 + th)^2
#pragma once

#include "loss_function.h"

namespace moptimizer::loss {

template <typename T>
class GemmanMCClure : public ILossFunction<T> {
 public:
  using Ptr = std::shared_ptr<GemmanMCClure>;
  explicit GemmanMCClure(T threshold) : threshold_(threshold) {}

  T weight(T errorSquaredNorm) override {
    const T shifted = errorSquaredNorm + threshold_;
    return (threshold_ * threshold_) / (shifted * shifted);
  }
  bool deviceLoss(int* kind, double* parameter) const override {
    *kind = MOPT_LOSS_GEMAN_MCCLURE;
    *parameter = double(threshold_);
    return true;
  }
  T threshold() const { return threshold_; }

 private:
  T threshold_;
};

}  // namespace moptimizer::loss
