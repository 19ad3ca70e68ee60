// RAII views of the C ABI objects (mopt_capi.h) used by the C++ mirror of the reference API.
// New relative to the reference: the "device-buffer residual store" the north star asks for.
#pragma once

#include <cstdint>
#include <memory>
#include <string>

#include "mopt_capi.h"
#include "moptimizer/exception.h"

namespace moptimizer::device {

inline void check(int status, const char* what) {
  if (status != MOPT_OK) throw moptimizer::Exception(std::string(what) + ": " + mopt_last_error(), status);
}

/// One GPU.  Fails loudly (exception) when no CUDA device / library is available: there is no CPU path.
class Context {
 public:
  using Ptr = std::shared_ptr<Context>;
  static Ptr create(int device = 0) { return Ptr(new Context(device)); }
  /// One rank of a sharded problem; `unique_id` (MOPT_NCCL_ID_BYTES bytes) comes from uniqueId() on rank 0.
  static Ptr createSharded(int device, int rank, int world, const void* unique_id) {
    return Ptr(new Context(device, rank, world, unique_id));
  }
  static void uniqueId(void* out) { check(mopt_comm_unique_id(out), "mopt_comm_unique_id"); }
  ~Context() { mopt_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  mopt_ctx* get() const { return ctx_; }
  void synchronize() { check(mopt_ctx_synchronize(ctx_), "mopt_ctx_synchronize"); }

 private:
  explicit Context(int device) { check(mopt_ctx_create(device, &ctx_), "mopt_ctx_create"); }
  Context(int device, int rank, int world, const void* id) {
    check(mopt_ctx_create_sharded(device, rank, world, id, &ctx_), "mopt_ctx_create_sharded");
  }
  mopt_ctx* ctx_ = nullptr;
};

template <class Scalar>
constexpr int dtypeOf() { return sizeof(Scalar) == 4 ? MOPT_F32 : MOPT_F64; }

/// Planar fp32/fp64 data streams in HBM holding one model's residual data.
class Store {
 public:
  using Ptr = std::shared_ptr<Store>;
  Store(Context::Ptr ctx, int model, int dtype, int64_t n) : ctx_(std::move(ctx)), n_(n) {
    check(mopt_store_create(ctx_->get(), model, dtype, n, &store_), "mopt_store_create");
  }
  ~Store() { mopt_store_destroy(store_); }
  Store(const Store&) = delete;
  Store& operator=(const Store&) = delete;
  template <class HostScalar>
  void upload(int group, const HostScalar* host, int64_t count, int64_t host_stride = 0, int64_t first = 0) {
    check(mopt_store_upload(store_, group, host, dtypeOf<HostScalar>(), host_stride, first, count), "mopt_store_upload");
  }
  template <class HostScalar>
  void download(int group, HostScalar* host, int64_t count, int64_t first = 0) {
    check(mopt_store_download(store_, group, host, dtypeOf<HostScalar>(), first, count), "mopt_store_download");
  }
  void generate(const mopt_synth& desc) { check(mopt_store_generate(store_, &desc), "mopt_store_generate"); }
  mopt_store* get() const { return store_; }
  const Context::Ptr& context() const { return ctx_; }
  int64_t size() const { return n_; }

 private:
  Context::Ptr ctx_;
  mopt_store* store_ = nullptr;
  int64_t n_ = 0;
};

/// Fixed target cloud in a device uniform grid: the data behind `model->update(x)` (model.h:24-26).
class NNIndex {
 public:
  using Ptr = std::shared_ptr<NNIndex>;
  template <class HostScalar>
  NNIndex(Context::Ptr ctx, const HostScalar* target_xyz, int64_t m, double max_distance, int index_dtype = MOPT_F64)
      : ctx_(std::move(ctx)) {
    check(mopt_nn_index_create(ctx_->get(), target_xyz, dtypeOf<HostScalar>(), index_dtype, m, max_distance, &index_),
          "mopt_nn_index_create");
  }
  ~NNIndex() { mopt_nn_index_destroy(index_); }
  NNIndex(const NNIndex&) = delete;
  NNIndex& operator=(const NNIndex&) = delete;
  mopt_nn_index* get() const { return index_; }

 private:
  Context::Ptr ctx_;
  mopt_nn_index* index_ = nullptr;
};

}  // namespace moptimizer::device
