// Mirrors include/moptimizer/model.h:11-104: the user-model interface.  Virtual f / f_df cannot be called
// from a kernel, so the device path accepts the builtin device models of moptimizer/device/models.h (which
// implement this interface); BaseModel / BaseModelJacobian are kept for source compatibility.
#pragma once

#include <memory>

#include "moptimizer/exception.h"

namespace moptimizer {

template <typename Scalar>
class IBaseModel {
 public:
  using Ptr = std::shared_ptr<IBaseModel>;
  using ConstPtr = std::shared_ptr<const IBaseModel>;
  IBaseModel() = default;
  virtual ~IBaseModel() = default;
  virtual void setup(const Scalar* x) = 0;
  virtual void update(const Scalar* x) = 0;
  virtual bool f(const Scalar* x, Scalar* f_x, unsigned int index) const = 0;
  virtual bool f_df(const Scalar* x, Scalar* f_x, Scalar* jacobian, unsigned int index) const = 0;
  virtual Ptr clone() const = 0;
};

template <typename Scalar, class ModelT>
class BaseModel : public IBaseModel<Scalar> {
 public:
  using Ptr = std::shared_ptr<ModelT>;
  using ConstPtr = std::shared_ptr<const ModelT>;
  void setup(const Scalar*) override {}
  void update(const Scalar*) override {}
  bool f(const Scalar* x, Scalar* f_x, unsigned int index) const override = 0;
  bool f_df(const Scalar*, Scalar*, Scalar*, unsigned int) const final {
    throw moptimizer::Exception("Non implemented non-jacobian model function `f_df` being used.");
  }
  std::shared_ptr<IBaseModel<Scalar>> clone() const override {
    return std::make_shared<ModelT>(*static_cast<const ModelT*>(this));
  }
};

template <typename Scalar, class ModelT>
class BaseModelJacobian : public IBaseModel<Scalar> {
 public:
  using Ptr = std::shared_ptr<ModelT>;
  using ConstPtr = std::shared_ptr<const ModelT>;
  void setup(const Scalar*) override {}
  void update(const Scalar*) override {}
  bool f(const Scalar*, Scalar*, unsigned int) const override {
    throw moptimizer::Exception("Non implemented jacobian model function `f` being used.");
  }
  bool f_df(const Scalar* x, Scalar* f_x, Scalar* jacobian, unsigned int index) const override = 0;
  std::shared_ptr<IBaseModel<Scalar>> clone() const override {
    return std::make_shared<ModelT>(*static_cast<const ModelT*>(this));
  }
};

}  // namespace moptimizer
