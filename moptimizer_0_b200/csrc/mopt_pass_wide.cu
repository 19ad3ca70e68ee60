// Instantiations + dispatch of the wide pass kernel (mopt_pass.cuh): models with more than 32 packed sums.
#include <cstdlib>

#include "mopt_internal.h"
#include "mopt_pass_wide_tc.cuh"

namespace mopt {
namespace {

constexpr int kThreads = 256;

template <class M, typename ST, typename CT, bool AFFINE_FD>
int launch_one(const PassLaunch& L, const PassArgs& a) {
  auto kern = wide_pass_kernel<M, ST, CT, kThreads, true, AFFINE_FD>;
  constexpr size_t smem = wide_smem_bytes<M, CT, kThreads>();
  // function attributes are per device: opt in to the large dynamic shared-memory tile once on each
  static bool configured[64] = {false};
  int dev = 0;
  MOPT_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    MOPT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem) != cudaSuccess || occ < 1) occ = 1;
  int64_t grid = int64_t(occ) * L.num_sms;
  const int64_t need = (a.n + kThreads - 1) / kThreads;  // one residual per lane per sweep
  if (need < grid) grid = need;
  if (grid < 1) grid = 1;
  if (grid > kMaxGrid) grid = kMaxGrid;
  kern<<<int(grid), kThreads, smem, L.stream>>>(a);
  MOPT_CUDA_TRY(cudaGetLastError());
  return MOPT_OK;
}

// Second generation, fp32 only: packed-fp32 finite differences + Gram matrix on the tensor cores (mopt_pass_wide_tc.cuh).
template <class M, int FLUSH>
int launch_tc_flush(const PassLaunch& L, const PassArgs& a) {
  constexpr int kTcThreads = 256;
  auto kern = wide_tc_kernel<M, kTcThreads, 2, FLUSH>;
  constexpr size_t smem = wide_tc_smem_bytes(M::O, M::SETN, M::P, kTcThreads);
  static bool configured[64] = {false};
  int dev = 0;
  MOPT_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    MOPT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTcThreads, smem) != cudaSuccess || occ < 1) occ = 1;
  if (L.ctas_per_sm > 0 && L.ctas_per_sm < occ) occ = L.ctas_per_sm;
  int64_t grid = int64_t(occ) * L.num_sms;
  const int64_t need = (a.n + 2 * kTcThreads - 1) / (2 * kTcThreads);  // two observations per lane per sweep
  if (need < grid) grid = need;
  if (grid < 1) grid = 1;
  if (grid > kMaxGrid) grid = kMaxGrid;
  static const bool prefetch = [] {  // MOPT_WIDE_TC_PREFETCH=0: load the observations where they are used (A/B)
    const char* e = getenv("MOPT_WIDE_TC_PREFETCH");
    return !(e && e[0] == '0');
  }();
  PassArgs aa = a;
  aa.no_ring = prefetch ? 0 : 1;  // (the field belongs to the point2point ring kernel; unused by this one otherwise)
  kern<<<int(grid), kTcThreads, smem, L.stream>>>(aa);
  MOPT_CUDA_TRY(cudaGetLastError());
  return MOPT_OK;
}

template <class M>
int launch_tc(const PassLaunch& L, const PassArgs& a) {
  static const int flush = [] {  // experiment knob: MOPT_WIDE_TC_FLUSH=1 / 2 folds the fp32 accumulators into fp64 more often than every 4 groups
    const char* e = getenv("MOPT_WIDE_TC_FLUSH");
    return (e && e[0]) ? atoi(e) : 4;
  }();
  return flush >= 4 ? launch_tc_flush<M, 4>(L, a) : (flush >= 2 ? launch_tc_flush<M, 2>(L, a) : launch_tc_flush<M, 1>(L, a));
}

template <class M>
int launch_types(const PassLaunch& L, int store_dtype, int compute_dtype, const PassArgs& a) {
  // fp32 store + fp32 compute, common-denominator finite differences, C = I: the tensor-core kernel
  // (mopt_ctx_set_launch(.., 1024) keeps the first-generation kernel for A/B on the same box)
  if (store_dtype == MOPT_F32 && compute_dtype == MOPT_F32 && L.affine_fd && L.identity_cov && !a.masked &&
      L.threads != 1024)
    return launch_tc<M>(L, a);
  // L.affine_fd (launch_pass): common-denominator finite differences (wide_pass_kernel AFFINE_FD) — the default
  // with fp32 compute unless MOPT_FLAG_GENERIC_KERNEL asks for the per-residual form; fp64 compute is the literal
  // restatement of the reference unless MOPT_FLAG_STABLE_FD opts in
  if (store_dtype == MOPT_F32 && compute_dtype == MOPT_F32)
    return L.affine_fd ? launch_one<M, float, float, true>(L, a) : launch_one<M, float, float, false>(L, a);
  if (store_dtype == MOPT_F32 && compute_dtype == MOPT_F64)
    return L.affine_fd ? launch_one<M, float, double, true>(L, a) : launch_one<M, float, double, false>(L, a);
  if (store_dtype == MOPT_F64 && compute_dtype == MOPT_F64)
    return L.affine_fd ? launch_one<M, double, double, true>(L, a) : launch_one<M, double, double, false>(L, a);
  set_last_error("store dtype f64 with compute dtype f32 is not supported");
  return MOPT_ERR_UNSUPPORTED;
}

}  // namespace

int launch_wide(const PassLaunch& L, int model, int store_dtype, int compute_dtype, const PassArgs& a) {
  switch (model) {
    case MOPT_MODEL_PINHOLE_DISTORT: return launch_types<PinholeDistortModel>(L, store_dtype, compute_dtype, a);
    default:
      set_last_error("unknown wide model kind");
      return MOPT_ERR_INVALID_ARGUMENT;
  }
}

}  // namespace mopt
