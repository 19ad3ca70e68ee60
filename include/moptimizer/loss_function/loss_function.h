// Mirrors include/moptimizer/loss_function/loss_function.h:7-23.  The host `weight` is kept for API
// parity; on the device path the cost function recognises the concrete type and selects the
// corresponding kernel-side weight (mopt_loss in mopt_capi.h).
#pragma once

#include <memory>

namespace moptimizer::loss {

template <typename T>
class ILossFunction {
 public:
  using Ptr = std::shared_ptr<ILossFunction>;
  using ConstPtr = std::shared_ptr<const ILossFunction>;
  ILossFunction() = default;
  virtual ~ILossFunction() = default;
  virtual T weight(T errorSquaredNorm) = 0;
};

template <typename T>
class NoLoss : public ILossFunction<T> {
 public:
  T weight(T) override { return T(1.0); }
};

}  // namespace moptimizer::loss
