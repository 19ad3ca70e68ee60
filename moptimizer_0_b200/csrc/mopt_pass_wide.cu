// Instantiations + dispatch of the wide pass kernel (mopt_pass.cuh): models with more than 32 packed sums.
#include "mopt_internal.h"

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

template <class M>
int launch_types(const PassLaunch& L, int store_dtype, int compute_dtype, const PassArgs& a) {
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
