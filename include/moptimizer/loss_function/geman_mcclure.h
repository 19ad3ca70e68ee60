// Geman-McClure weight, class name as the reference spells it (include/moptimizer/loss_function/geman_mcclure.h:7-19):
//   w(e2) = th^2 / (e2 + th)^2
#pragma once

#include "loss_function.h"

namespace moptimizer::loss {

template <typename T> struct GemmanMCClure : ILossFunction<T> {
  typedef std::shared_ptr<GemmanMCClure> Ptr;

  explicit GemmanMCClure(T threshold) : th_(threshold) {}
  T threshold() const { return th_; }

  T weight(T errorSquaredNorm) override {
    const T shifted = errorSquaredNorm + th_;
    return (th_ * th_) / (shifted * shifted);
  }
  bool deviceLoss(int* kind, double* parameter) const override {
    *kind = MOPT_LOSS_GEMAN_MCCLURE, *parameter = double(th_);
    return true;
  }

 private:
  T th_;
};

}  // namespace moptimizer::loss
