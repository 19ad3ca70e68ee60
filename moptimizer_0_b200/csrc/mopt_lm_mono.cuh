// Small problems: the WHOLE Levenberg-Marquardt loop in one persistent kernel.
//
// With a few tens of thousands of residuals a pass takes microseconds and mopt_lm_minimize's per-trial sequence
// [pass kernel] -> [lm_step_kernel] is bound by the two dependent launches and the one-warp optimizer step between
// them (tst/point2point's 29 310-point cloud: pass 8.6 us, step 17.7 us, 22 k LM iterations per second).  Here one
// cooperative launch keeps every CTA resident; each trial is
//     pass over the CTA's residuals -> CTA partials -> arrival ticket ->
//     CTA 0 (always the same SM: the serial tail stays in its instruction cache) reduces, assembles (H, b, sum), runs the optimizer transition (accept / reject, LDL^T solve,
//     x (+) delta, model->setup) and opens the grid barrier -> every CTA reads the new parameter block and goes on,
// until the state machine sets PASS_SKIP.  No launch, no host round trip and no second kernel per trial.
// Restates the loop of LevenbergMarquadtDynamic::minimize (src/levenberg_marquadt_dyn.cpp:34-119) around
// p2p_moment_body (mopt_pass.cuh); same arithmetic and summation order as the multi-launch path with the same grid.
#pragma once

#include "mopt_lm.cuh"
#include "mopt_pass.cuh"

namespace mopt {

struct MonoArgs {
  LmState* st;
  CostSlot* slots;
  unsigned long long* gen;      // grid-barrier generation counter: monotonic over the context's life, never reset
  unsigned long long gen_base;  // its value when this launch starts
  int max_slots;
  int loss, qrot;               // mopt_loss_kind / "the moments are taken in q = R p" of the pass, chosen at run time here
  LmInit init;                  // prepare() arguments: CTA 0 initialises the state before the first pass
  unsigned long long* dbg;  // optional (MOPT_LM_MONO_TRACE=1): 4 globaltimer stamps per trial from the last CTA
};

#ifdef __CUDACC__

// The loss kind and the Jacobian form are run-time switches over the six instantiations of the pass body (a small
// problem does not care), so that there is ONE entry per (store, compute) type pair: ptxas compiles the optimizer
// transition once per entry, and with a kernel per (loss, form) this translation unit took six minutes to build.
template <typename ST, typename CT, int THREADS, int UNROLL, int FLUSH_ROUNDS>
__device__ __forceinline__ bool p2p_mono_pass(const PassArgs& a, int mode, int loss, bool qrot) {
#define MOPT_MONO_BODY(L, Q) p2p_moment_body<ST, CT, L, Q, THREADS, UNROLL, FLUSH_ROUNDS, 0, false, false, true>(a, mode)
  switch (loss) {
    case MOPT_LOSS_NONE: return qrot ? MOPT_MONO_BODY(MOPT_LOSS_NONE, true) : MOPT_MONO_BODY(MOPT_LOSS_NONE, false);
    case MOPT_LOSS_GEMAN_MCCLURE:
      return qrot ? MOPT_MONO_BODY(MOPT_LOSS_GEMAN_MCCLURE, true) : MOPT_MONO_BODY(MOPT_LOSS_GEMAN_MCCLURE, false);
    default: return qrot ? MOPT_MONO_BODY(MOPT_LOSS_HUBER, true) : MOPT_MONO_BODY(MOPT_LOSS_HUBER, false);
  }
#undef MOPT_MONO_BODY
}

template <typename ST, typename CT, int THREADS, int MINB, int UNROLL, int FLUSH_ROUNDS>
__global__ void __launch_bounds__(THREADS, MINB) p2p_lm_mono_kernel(const PassArgs a, const MonoArgs m) {
  __shared__ LmStepShared s_sh;
  auto open_barrier = [&](unsigned long long value) {  // thread 0 of the CTA that did the serial work
    __threadfence();  // state, ParamBlock and control word before the barrier opens
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(m.gen), "l"(value) : "memory");
  };
  auto wait_barrier = [&](unsigned long long value) {  // thread 0 of every other CTA
    unsigned long long g;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(g) : "l"(m.gen) : "memory");
      if (g >= value) break;
      __nanosleep(100);
    }
    __threadfence();
  };
  // prepare() + the first setup(x0): CTA 0, everyone else waits (no separate init kernel, no launch gap)
  if (blockIdx.x == 0) {
    if (threadIdx.x < 32) lm_init_warp(m.st, m.slots, m.init, &s_sh, threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) open_barrier(m.gen_base + 1);
  } else if (threadIdx.x == 0) {
    wait_barrier(m.gen_base + 1);
  }
  __syncthreads();
  for (int slot = 0; slot < m.max_slots; ++slot) {
    // the control word and the ParamBlock were written by another CTA, possibly on another SM: read past the L1
    // (the barrier also fences, which drops this SM's L1 lines)
    const int mode = __ldcg(a.mode_ptr);
    if (mode == PASS_SKIP) break;
    unsigned long long t_begin = 0;
    if (m.dbg) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_begin));
    const bool last = p2p_mono_pass<ST, CT, THREADS, UNROLL, FLUSH_ROUNDS>(a, mode, m.loss, m.qrot != 0);
    const unsigned long long target = m.gen_base + 2ull + (unsigned long long)(slot);
    if (last) {
      __syncthreads();  // `out` complete (assembled by the whole CTA)
      unsigned long long t_pass = 0, t_step = 0, t_open = 0;
      if (m.dbg && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_pass));
      if (threadIdx.x < 32)
        lm_step_warp(m.st, a.out, m.slots, &s_sh, threadIdx.x, m.init.P, m.init.scalar_f32 != 0,
                     (m.dbg && slot < 16) ? reinterpret_cast<long long*>(m.dbg) + 256 + slot * 16 : nullptr);
      __syncthreads();
      if (threadIdx.x == 0) {
        if (m.dbg) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_step));
        open_barrier(target);
        if (m.dbg && slot < 64) {
          asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_open));
          m.dbg[slot * 4 + 0] = t_begin; m.dbg[slot * 4 + 1] = t_pass; m.dbg[slot * 4 + 2] = t_step; m.dbg[slot * 4 + 3] = t_open;
        }
      }
    } else if (threadIdx.x == 0) {
      wait_barrier(target);
    }
    __syncthreads();
  }
}

#endif  // __CUDACC__
}  // namespace mopt
