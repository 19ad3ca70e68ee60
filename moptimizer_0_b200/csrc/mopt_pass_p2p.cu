// Instantiations + dispatch of the point-to-point moment kernel (mopt_pass.cuh).
#include "mopt_internal.h"
#include "mopt_lm_mono.cuh"
#include "mopt_pass_p2p2.cuh"

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

// Launch shapes (profiles/r1_tune_p2p.txt): the fp32 kernel is bound by bytes in flight, not by its ALU
// work, and does best with the whole SM's 32 warps x 6 outstanding 16-byte loads (64 registers/thread).
struct ShapeF32 {
  static constexpr int THREADS = 1024, MINB = 1, UNROLL = 1, FLUSH = 16;
};
// Second-generation fp32 kernel (mopt_pass_p2p2.cuh): TMA bulk-copy ring + packed fp32 arithmetic.  Shapes from the
// same-box sustained sweeps (profiles/r2_tune_gen2_*.txt): 384 threads x 2 stages x 2 float4 groups per thread and
// stream (72 KB stages, 144 KB ring) reads within 1 % of the bare read-and-sum ceiling on both boxes measured;
// 512 x 3 x 1 is as fast on one of them.  mopt_ctx_set_launch(.., 512) selects the second, (.., 1024) the
// first-generation kernel above (A/B on the same box).
struct ShapeGen2 {
  static constexpr int THREADS = 384, STAGES = 2, U = 2, FLUSH = 16;
};
struct ShapeGen2Wide {
  static constexpr int THREADS = 512, STAGES = 3, U = 1, FLUSH = 32;
};
// fp64 compute (the reference's default Scalar): capped at 128 registers so two CTAs stay resident — 16 warps hide the
// DFMA latency that 8 could not.  f32 store 873 -> 622 us per 100 M, f64 store 553 -> 380 us per 50 M (6.3 TB/s)
// (profiles/r1_tune_p2p_f64.txt).
struct ShapeF64 {
  static constexpr int THREADS = 256, MINB = 2, UNROLL = 2, FLUSH = 8;
};

template <typename ST, typename CT, int LOSS, bool QROT, class S, bool FUSED>
int launch_fused(const PassLaunch& L, const PassArgs& a) {
  auto kern = p2p_moment_kernel<ST, CT, LOSS, QROT, S::THREADS, S::MINB, S::UNROLL, S::FLUSH, 0, false, FUSED>;
  const int64_t groups = a.n / VecOf<ST>::N;
  // small problems: keep CTAs at 256 threads' worth of work granularity by capping the grid, never below 1
  const int grid = pick_grid(reinterpret_cast<const void*>(kern), S::THREADS, L, groups);
  kern<<<grid, S::THREADS, 0, L.stream>>>(a);
  MOPT_CUDA_TRY(cudaGetLastError());
  return MOPT_OK;
}

// PassArgs::fused_setup (host-driven analytical passes): the instantiation that runs model->setup(x) itself
template <typename ST, typename CT, int LOSS, bool QROT, class S>
int launch_one(const PassLaunch& L, const PassArgs& a) {
  return a.fused_setup ? launch_fused<ST, CT, LOSS, QROT, S, true>(L, a) : launch_fused<ST, CT, LOSS, QROT, S, false>(L, a);
}

// ---- second generation (fp32 store, fp32 compute) ----------------------------------------------------------
template <int LOSS, bool QROT, bool MASKED, bool FUSED, class S>
int launch_gen2_one(const PassLaunch& L, const PassArgs& a0) {
  auto kern = p2p_moment2_kernel<LOSS, QROT, MASKED, FUSED, S::THREADS, 1, S::STAGES, S::U, S::FLUSH, 0>;
  constexpr size_t smem = p2p2_ring_bytes(S::THREADS, S::STAGES, S::U);
  static bool configured[64] = {false};  // function attributes are per device
  int dev = 0;
  MOPT_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    MOPT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  PassArgs a = a0;
  // the bulk copies need 16-byte aligned sources; anything else streams through the kernel's direct-load path
  for (int k = 0; k < 6; ++k)
    if (reinterpret_cast<uintptr_t>(a.streams.p[k]) & 15u) a.no_ring = 1;
  const int64_t groups = a.n / 4;
  int64_t grid = int64_t(L.ctas_per_sm > 0 ? L.ctas_per_sm : 1) * L.num_sms;  // one CTA per SM owns the whole ring
  if (grid > L.num_sms) grid = L.num_sms;
  const int64_t need = (groups + S::THREADS - 1) / S::THREADS;
  if (need < grid) grid = need;
  if (grid < 1) grid = 1;
  kern<<<int(grid), S::THREADS, smem, L.stream>>>(a);
  MOPT_CUDA_TRY(cudaGetLastError());
  return MOPT_OK;
}

template <int LOSS, bool QROT, class S>
int launch_gen2_flags(const PassLaunch& L, const PassArgs& a) {
  if (a.masked)
    return a.fused_setup ? launch_gen2_one<LOSS, QROT, true, true, S>(L, a) : launch_gen2_one<LOSS, QROT, true, false, S>(L, a);
  return a.fused_setup ? launch_gen2_one<LOSS, QROT, false, true, S>(L, a) : launch_gen2_one<LOSS, QROT, false, false, S>(L, a);
}

template <class S>
int launch_gen2(const PassLaunch& L, int loss, bool qrot, const PassArgs& a) {
  switch (loss) {
    case MOPT_LOSS_NONE:
      return qrot ? launch_gen2_flags<MOPT_LOSS_NONE, true, S>(L, a) : launch_gen2_flags<MOPT_LOSS_NONE, false, S>(L, a);
    case MOPT_LOSS_GEMAN_MCCLURE:
      return qrot ? launch_gen2_flags<MOPT_LOSS_GEMAN_MCCLURE, true, S>(L, a)
                  : launch_gen2_flags<MOPT_LOSS_GEMAN_MCCLURE, false, S>(L, a);
    case MOPT_LOSS_HUBER:
      return qrot ? launch_gen2_flags<MOPT_LOSS_HUBER, true, S>(L, a) : launch_gen2_flags<MOPT_LOSS_HUBER, false, S>(L, a);
    default:
      set_last_error("unknown loss kind");
      return MOPT_ERR_INVALID_ARGUMENT;
  }
}

template <typename ST, typename CT, class S>
int launch_loss(const PassLaunch& L, int loss, bool qrot, const PassArgs& a) {
  switch (loss) {
    case MOPT_LOSS_NONE:
      return qrot ? launch_one<ST, CT, MOPT_LOSS_NONE, true, S>(L, a) : launch_one<ST, CT, MOPT_LOSS_NONE, false, S>(L, a);
    case MOPT_LOSS_GEMAN_MCCLURE:
      return qrot ? launch_one<ST, CT, MOPT_LOSS_GEMAN_MCCLURE, true, S>(L, a)
                  : launch_one<ST, CT, MOPT_LOSS_GEMAN_MCCLURE, false, S>(L, a);
    case MOPT_LOSS_HUBER:
      return qrot ? launch_one<ST, CT, MOPT_LOSS_HUBER, true, S>(L, a) : launch_one<ST, CT, MOPT_LOSS_HUBER, false, S>(L, a);
    default:
      set_last_error("unknown loss kind");
      return MOPT_ERR_INVALID_ARGUMENT;
  }
}

}  // namespace

// ---- persistent LM kernel for small problems (mopt_lm_mono.cuh) ------------------------------------------------
namespace {
struct ShapeMono {  // small CTAs: more of them share a small problem; one per SM so the optimizer step does not spill
  static constexpr int THREADS = 256, MINB = 1, UNROLL = 2, FLUSH = 8;
};

template <typename ST, typename CT, int LOSS, bool QROT>
int launch_mono_one(const PassLaunch& L, const PassArgs& a, const MonoArgs& m) {
  using S = ShapeMono;
  auto kern = p2p_lm_mono_kernel<ST, CT, LOSS, QROT, S::THREADS, S::MINB, S::UNROLL, S::FLUSH>;
  const int64_t groups = a.n / VecOf<ST>::N;
  const int grid = pick_grid(reinterpret_cast<const void*>(kern), S::THREADS, L, groups);  // <= co-resident CTAs
  PassArgs aa = a;
  MonoArgs mm = m;
  void* params[] = {&aa, &mm};
  MOPT_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(S::THREADS), params, 0,
                                            L.stream));
  return MOPT_OK;
}

template <typename ST, typename CT>
int launch_mono_loss(const PassLaunch& L, int loss, bool qrot, const PassArgs& a, const MonoArgs& m) {
  switch (loss) {
    case MOPT_LOSS_NONE:
      return qrot ? launch_mono_one<ST, CT, MOPT_LOSS_NONE, true>(L, a, m) : launch_mono_one<ST, CT, MOPT_LOSS_NONE, false>(L, a, m);
    case MOPT_LOSS_GEMAN_MCCLURE:
      return qrot ? launch_mono_one<ST, CT, MOPT_LOSS_GEMAN_MCCLURE, true>(L, a, m)
                  : launch_mono_one<ST, CT, MOPT_LOSS_GEMAN_MCCLURE, false>(L, a, m);
    case MOPT_LOSS_HUBER:
      return qrot ? launch_mono_one<ST, CT, MOPT_LOSS_HUBER, true>(L, a, m) : launch_mono_one<ST, CT, MOPT_LOSS_HUBER, false>(L, a, m);
    default:
      set_last_error("unknown loss kind");
      return MOPT_ERR_INVALID_ARGUMENT;
  }
}
}  // namespace

int launch_p2p_lm_mono(const PassLaunch& L, int store_dtype, int compute_dtype, int loss, bool qrot, const PassArgs& a,
                       const MonoArgs& m) {
  if (store_dtype == MOPT_F32 && compute_dtype == MOPT_F32) return launch_mono_loss<float, float>(L, loss, qrot, a, m);
  if (store_dtype == MOPT_F32 && compute_dtype == MOPT_F64) return launch_mono_loss<float, double>(L, loss, qrot, a, m);
  if (store_dtype == MOPT_F64 && compute_dtype == MOPT_F64) return launch_mono_loss<double, double>(L, loss, qrot, a, m);
  set_last_error("point2point: store dtype f64 with compute dtype f32 is not supported");
  return MOPT_ERR_UNSUPPORTED;
}

int launch_p2p_moment(const PassLaunch& L, int store_dtype, int compute_dtype, int loss, bool qrot, const PassArgs& a) {
  if (store_dtype == MOPT_F32 && compute_dtype == MOPT_F32) {
    if (L.threads == 1024) return launch_loss<float, float, ShapeF32>(L, loss, qrot, a);  // first generation (A/B)
    return (L.threads == 512) ? launch_gen2<ShapeGen2Wide>(L, loss, qrot, a) : launch_gen2<ShapeGen2>(L, loss, qrot, a);
  }
  if (store_dtype == MOPT_F32 && compute_dtype == MOPT_F64) return launch_loss<float, double, ShapeF64>(L, loss, qrot, a);
  if (store_dtype == MOPT_F64 && compute_dtype == MOPT_F64) return launch_loss<double, double, ShapeF64>(L, loss, qrot, a);
  set_last_error("point2point: store dtype f64 with compute dtype f32 is not supported");
  return MOPT_ERR_UNSUPPORTED;
}

}  // namespace mopt
