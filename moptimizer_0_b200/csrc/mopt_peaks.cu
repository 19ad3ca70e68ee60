// Measured ceilings of the box this library runs on, for the roofline denominators bench.py cannot take from
// MEASURED_PEAKS.json (which only holds an HBM copy and a bf16 GEMM figure): fp32 FMA rate (scalar FFMA and packed
// FFMA2), the warp-instruction issue rate that follows from it, a read-only HBM stream, and the pinned host -> device
// copy rate of this process.  SURVEY.md §8d: "an fp32 FMA peak must be measured on the box".  Micro-kernels only; nothing here is on
// the product path.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

#include "mopt_internal.h"

namespace {

using namespace mopt;

// ILP independent FMA chains per thread, `iters` rounds; b and c come from memory so nothing folds at compile time.
template <int ILP>
__global__ void __launch_bounds__(512) fma_peak_kernel(const float* bc, float* out, int iters) {
  const float b = bc[0], c = bc[1];
  float a[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) a[k] = float(threadIdx.x + k) * 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) a[k] = fmaf(a[k], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += a[k];
  if (s == 123.456f) out[0] = s;  // never true in practice: keeps the chains alive
}

template <int ILP>
__global__ void __launch_bounds__(512) fma2_peak_kernel(const float* bc, float* out, int iters) {
  const float2 b = make_float2(bc[0], bc[0]), c = make_float2(bc[1], bc[1]);
  float2 a[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) a[k] = make_float2(float(threadIdx.x + k) * 1e-3f, float(k) * 1e-3f);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) a[k] = __ffma2_rn(a[k], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += a[k].x + a[k].y;
  if (s == 123.456f) out[0] = s;
}

// Read-only stream: 16-byte streaming loads, 4 in flight per thread, one FADD per value.
__global__ void __launch_bounds__(256) read_peak_kernel(const float4* p, int64_t n4, float* out) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  float acc = 0.f;
  int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; g + 3 * stride < n4; g += 4 * stride) {
    const float4 v0 = ld_stream(p + g), v1 = ld_stream(p + g + stride), v2 = ld_stream(p + g + 2 * stride),
                 v3 = ld_stream(p + g + 3 * stride);
    acc += (v0.x + v0.y) + (v0.z + v0.w) + (v1.x + v1.y) + (v1.z + v1.w) + (v2.x + v2.y) + (v2.z + v2.w) +
           (v3.x + v3.y) + (v3.z + v3.w);
  }
  for (; g < n4; g += stride) {
    const float4 v = ld_stream(p + g);
    acc += (v.x + v.y) + (v.z + v.w);
  }
  if (acc == 123.456f) out[0] = acc;
}

template <class F>
int time_best(cudaStream_t stream, F launch, int reps, float* best_ms) {
  cudaEvent_t e0, e1;
  MOPT_CUDA_TRY(cudaEventCreate(&e0));
  MOPT_CUDA_TRY(cudaEventCreate(&e1));
  launch();  // warm-up
  MOPT_CUDA_TRY(cudaStreamSynchronize(stream));
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    MOPT_CUDA_TRY(cudaEventRecord(e0, stream));
    launch();
    MOPT_CUDA_TRY(cudaEventRecord(e1, stream));
    MOPT_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    MOPT_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  MOPT_CUDA_TRY(cudaGetLastError());
  *best_ms = best;
  return MOPT_OK;
}

}  // namespace

extern "C" int mopt_measure_peaks(mopt_ctx* ctx, mopt_peaks* out) try {
  MOPT_REQUIRE(ctx && out, "null argument");
  MOPT_CUDA_TRY(cudaSetDevice(ctx->device));
  std::memset(out, 0, sizeof(*out));
  cudaStream_t s = ctx->stream;
  float *d_bc = nullptr, *d_out = nullptr;
  MOPT_CUDA_TRY(cudaMalloc(&d_bc, 2 * sizeof(float)));
  MOPT_CUDA_TRY(cudaMalloc(&d_out, sizeof(float)));
  const float bc[2] = {0.999f, 1e-3f};
  MOPT_CUDA_TRY(cudaMemcpy(d_bc, bc, sizeof(bc), cudaMemcpyHostToDevice));
  constexpr int ILP = 8, THREADS = 512;
  const int grid = ctx->num_sms * 4;  // 2048 resident threads per SM
  const int iters = 20000;
  const double thread_ops = double(grid) * THREADS * double(iters) * ILP;
  float ms = 0.f;
  MOPT_TRY(time_best(s, [&] { fma_peak_kernel<ILP><<<grid, THREADS, 0, s>>>(d_bc, d_out, iters); }, 5, &ms));
  out->fp32_fma_tflops = 2.0 * thread_ops / (ms * 1e-3) / 1e12;
  MOPT_TRY(time_best(s, [&] { fma2_peak_kernel<ILP><<<grid, THREADS, 0, s>>>(d_bc, d_out, iters); }, 5, &ms));
  out->fp32_fma2_tflops = 4.0 * thread_ops / (ms * 1e-3) / 1e12;
  // one FFMA issues per clock per scheduler: the FFMA-only rate is the warp-instruction issue ceiling of the SM
  // (an FFMA + LOP3 mix measured lower, 0.49 T/s against 1.1 T/s: the integer pipe is narrower)
  out->issue_gwarp_inst_per_s = out->fp32_fma_tflops * 1e12 / 64.0 / 1e9;
  // read-only HBM stream over 2 GiB (>> the 126 MB L2)
  {
    const int64_t bytes = int64_t(2) << 30;
    float4* buf = nullptr;
    MOPT_CUDA_TRY(cudaMalloc(&buf, size_t(bytes)));
    MOPT_CUDA_TRY(cudaMemsetAsync(buf, 0, size_t(bytes), s));
    const int64_t n4 = bytes / 16;
    MOPT_TRY(time_best(s, [&] { read_peak_kernel<<<ctx->num_sms * 8, 256, 0, s>>>(buf, n4, d_out); }, 10, &ms));
    out->hbm_read_gbs = double(bytes) / (ms * 1e-3) / 1e9;
    cudaFree(buf);
  }
  cudaFree(d_bc);
  cudaFree(d_out);
  return MOPT_OK;
}
MOPT_ABI_CATCH

// Pinned host -> device copy rate of THIS process: `bytes` per sweep, one plain cudaMemcpyAsync per 64 MB chunk (what
// mopt_store_upload issues), repeated for at least `seconds` of wall clock so that concurrent callers on other GPUs
// of the box overlap; the caller synchronises the ranks before calling.  *gbs = bytes moved / elapsed.
extern "C" int mopt_measure_h2d(mopt_ctx* ctx, uint64_t bytes, double seconds, double* gbs) try {
  MOPT_REQUIRE(ctx && gbs && bytes >= (1u << 20), "bad argument");
  MOPT_CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t chunk = size_t(64) << 20;
  void *h = nullptr, *d = nullptr;
  MOPT_CUDA_TRY(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
  std::memset(h, 1, bytes);  // first touch by this thread: local NUMA node under the caller's CPU binding
  if (cudaMalloc(&d, bytes) != cudaSuccess) {
    cudaFreeHost(h);
    (void)cudaGetLastError();
    mopt::set_last_error("mopt_measure_h2d: out of device memory");
    return MOPT_ERR_OUT_OF_MEMORY;
  }
  cudaStream_t s = ctx->stream;
  auto sweep = [&] {
    for (size_t o = 0; o < bytes; o += chunk)
      cudaMemcpyAsync(static_cast<char*>(d) + o, static_cast<char*>(h) + o, std::min(chunk, size_t(bytes) - o),
                      cudaMemcpyHostToDevice, s);
  };
  sweep();
  MOPT_CUDA_TRY(cudaStreamSynchronize(s));
  const auto t0 = std::chrono::steady_clock::now();
  double elapsed = 0.0;
  uint64_t moved = 0;
  do {
    sweep();
    MOPT_CUDA_TRY(cudaStreamSynchronize(s));
    moved += bytes;
    elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  } while (elapsed < seconds);
  *gbs = double(moved) / elapsed / 1e9;
  cudaFree(d);
  cudaFreeHost(h);
  return MOPT_OK;
}
MOPT_ABI_CATCH
