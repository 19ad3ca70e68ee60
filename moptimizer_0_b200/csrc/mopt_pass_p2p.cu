// Instantiations + dispatch of the point-to-point moment kernel (mopt_pass.cuh).
#include "mopt_internal.h"

namespace mopt {

int pick_grid(const void* kernel, int threads, const PassLaunch& L, int64_t work_items) {
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0) != cudaSuccess || occ < 1) occ = 1;
  int per_sm = (L.ctas_per_sm > 0) ? (L.ctas_per_sm < occ ? L.ctas_per_sm : occ) : occ;
  int64_t g = int64_t(per_sm) * L.num_sms;
  if (g > kMaxGrid) g = kMaxGrid;
  const int64_t need = (work_items + threads - 1) / threads;
  if (need < g) g = need;
  if (g < 1) g = 1;
  return int(g);
}

namespace {

constexpr int kThreads = 256;

template <typename ST, typename CT, int LOSS, bool QROT, int MINB>
int launch_one(const PassLaunch& L, const PassArgs& a) {
  auto kern = p2p_moment_kernel<ST, CT, LOSS, QROT, kThreads, MINB>;
  const int64_t groups = a.n / VecOf<ST>::N;
  const int grid = pick_grid(reinterpret_cast<const void*>(kern), kThreads, L, groups);
  kern<<<grid, kThreads, 0, L.stream>>>(a);
  MOPT_CUDA_TRY(cudaGetLastError());
  return MOPT_OK;
}

template <typename ST, typename CT, int MINB>
int launch_loss(const PassLaunch& L, int loss, bool qrot, const PassArgs& a) {
  switch (loss) {
    case MOPT_LOSS_NONE:
      return qrot ? launch_one<ST, CT, MOPT_LOSS_NONE, true, MINB>(L, a) : launch_one<ST, CT, MOPT_LOSS_NONE, false, MINB>(L, a);
    case MOPT_LOSS_GEMAN_MCCLURE:
      return qrot ? launch_one<ST, CT, MOPT_LOSS_GEMAN_MCCLURE, true, MINB>(L, a)
                  : launch_one<ST, CT, MOPT_LOSS_GEMAN_MCCLURE, false, MINB>(L, a);
    case MOPT_LOSS_HUBER:
      return qrot ? launch_one<ST, CT, MOPT_LOSS_HUBER, true, MINB>(L, a) : launch_one<ST, CT, MOPT_LOSS_HUBER, false, MINB>(L, a);
    default:
      set_last_error("unknown loss kind");
      return MOPT_ERR_INVALID_ARGUMENT;
  }
}

}  // namespace

int launch_p2p_moment(const PassLaunch& L, int store_dtype, int compute_dtype, int loss, bool qrot, const PassArgs& a) {
  if (store_dtype == MOPT_F32 && compute_dtype == MOPT_F32) return launch_loss<float, float, 2>(L, loss, qrot, a);
  if (store_dtype == MOPT_F32 && compute_dtype == MOPT_F64) return launch_loss<float, double, 1>(L, loss, qrot, a);
  if (store_dtype == MOPT_F64 && compute_dtype == MOPT_F64) return launch_loss<double, double, 1>(L, loss, qrot, a);
  set_last_error("point2point: store dtype f64 with compute dtype f32 is not supported");
  return MOPT_ERR_UNSUPPORTED;
}

}  // namespace mopt
