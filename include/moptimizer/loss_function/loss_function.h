// IRLS loss interface of the reference (include/moptimizer/loss_function/loss_function.h:7-23): weight(|r|^2).
// On the device path no virtual call can run per residual, so a loss additionally DESCRIBES itself to the cost
// function (deviceLoss): the kernels implement the weight of the named kind (mopt_loss in mopt_capi.h).  A loss
// that does not override deviceLoss() is a host-only loss and is rejected loudly — there is no CPU fallback.
#pragma once

#include <memory>

#include "mopt_capi.h"

namespace moptimizer::loss {

template <typename T> struct ILossFunction {
  typedef std::shared_ptr<ILossFunction> Ptr;
  typedef std::shared_ptr<const ILossFunction> ConstPtr;

  /// w(e2) applied to J^T C J and J^T C r (linearization.h:112-115,149-152); host copy kept for API parity.
  virtual T weight(T errorSquaredNorm) = 0;

  /// Kernel-side kind and parameter of this loss; false = not implemented on the device.
  virtual bool deviceLoss(int* /*kind*/, double* /*parameter*/) const { return false; }

  virtual ~ILossFunction() {}

 protected:
  ILossFunction() {}
};

/// w = 1 (loss_function.h:20-23).
template <typename T> struct NoLoss final : ILossFunction<T> {
  T weight(T) override { return T(1); }
  bool deviceLoss(int* kind, double* parameter) const override {
    *kind = MOPT_LOSS_NONE, *parameter = 0.0;
    return true;
  }
};

}  // namespace moptimizer::loss
