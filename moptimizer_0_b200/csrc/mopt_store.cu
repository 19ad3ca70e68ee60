// Device-buffer residual store: planar (structure-of-arrays) fp32 or fp64 streams in HBM, one per
// data component, so every pass kernel reads them with fully coalesced 16-byte loads.
//   point2point / point_dist : src x,y,z | tgt x,y,z     (24 B per correspondence in fp32)
//   exp_curve / michaelis    : t | y                     ( 8 B per sample)
//   pinhole                  : X,Y,Z | u,v               (20 B per observation)
// Upload de-interleaves the caller's AoS arrays (the layout the reference models point into,
// tst/point2point.cpp:16-17,82-83; tst/curve_fitting.cpp:88-89) on the device.
#include <cstring>

#include "mopt_internal.h"

using namespace mopt;

namespace {

// AoS (host layout, staged in device memory) -> planar streams, with dtype conversion.
template <typename HT, typename ST>
__global__ void deinterleave_kernel(const HT* __restrict__ aos, int64_t stride, int ncomp, ST* __restrict__ s0,
                                    ST* __restrict__ s1, ST* __restrict__ s2, int64_t first, int64_t count) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    const HT* e = aos + i * stride;
    s0[first + i] = ST(e[0]);
    if (ncomp > 1) s1[first + i] = ST(e[1]);
    if (ncomp > 2) s2[first + i] = ST(e[2]);
  }
}

template <typename HT, typename ST>
__global__ void interleave_kernel(HT* __restrict__ aos, int ncomp, const ST* __restrict__ s0, const ST* __restrict__ s1,
                                  const ST* __restrict__ s2, int64_t first, int64_t count) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    HT* e = aos + i * ncomp;
    e[0] = HT(s0[first + i]);
    if (ncomp > 1) e[1] = HT(s1[first + i]);
    if (ncomp > 2) e[2] = HT(s2[first + i]);
  }
}

// ---- counter-based synthetic data -----------------------------------------------------------
// u32 = hi32(splitmix64(seed ^ splitmix64(index * 16 + lane)));  uniform = (u32 >> 8) * 2^-24 (exact).
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t seed, int64_t index, int lane) {
  const uint64_t h = splitmix64(seed ^ splitmix64(uint64_t(index) * 16ull + uint64_t(lane)));
  return float(uint32_t(h >> 40)) * (1.0f / 16777216.0f);
}
// Irwin-Hall(4) approximation of N(0,1): (u1+u2+u3+u4 - 2) * sqrt(3).  Every operation is an explicitly rounded
// intrinsic (no compiler-chosen FMA contraction) so that a host restatement with fmaf() reproduces the generated
// streams bit for bit (oracle/oracle_capi.cpp orc_generate_p2p: the CPU arm of bench.py times the same workload).
__device__ __forceinline__ float approx_normal(uint64_t seed, int64_t index, int lane0) {
  const float s = __fadd_rn(__fadd_rn(u01(seed, index, lane0), u01(seed, index, lane0 + 1)),
                            __fadd_rn(u01(seed, index, lane0 + 2), u01(seed, index, lane0 + 3)));
  return __fmul_rn(__fadd_rn(s, -2.0f), 1.7320508f);
}

struct SynthDev {
  uint64_t seed;
  int64_t first_index, n_total;
  float gt[24];  // p2p: R (9) t (3); curve: m, c; pinhole: M (12); pinhole_distort: TC (12), fx..k3 (9)
  int distort;
  float lo[3], hi[3];
  float sigma, outlier_fraction, outlier_range;
};

template <typename ST>
__global__ void generate_p2p_kernel(SynthDev d, int64_t n, ST* sx, ST* sy, ST* sz, ST* tx, ST* ty, ST* tz) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t gi = d.first_index + i;
    float p[3], q[3];
    for (int k = 0; k < 3; ++k) p[k] = __fmaf_rn(__fadd_rn(d.hi[k], -d.lo[k]), u01(d.seed, gi, k), d.lo[k]);
    const bool outlier = u01(d.seed, gi, 15) < d.outlier_fraction;
    for (int k = 0; k < 3; ++k) {
      // tgt = R src + t (+ noise) (+ outlier offset), one fused multiply-add per term, in this order
      float v = __fmaf_rn(d.gt[k * 3 + 2], p[2], __fmaf_rn(d.gt[k * 3 + 1], p[1], __fmaf_rn(d.gt[k * 3 + 0], p[0], d.gt[9 + k])));
      if (d.sigma > 0.0f) v = __fmaf_rn(d.sigma, approx_normal(d.seed, gi, 3 + 4 * k), v);  // lanes 3..14
      if (outlier) v = __fmaf_rn(d.outlier_range, __fmaf_rn(2.0f, u01(d.seed, gi ^ 0x5bd1e995, k), -1.0f), v);
      q[k] = v;
    }
    sx[i] = ST(p[0]); sy[i] = ST(p[1]); sz[i] = ST(p[2]);
    tx[i] = ST(q[0]); ty[i] = ST(q[1]); tz[i] = ST(q[2]);
  }
}

template <typename ST>
__global__ void generate_curve_kernel(SynthDev d, int64_t n, ST* t, ST* y) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t gi = d.first_index + i;
    const float tt = d.lo[0] + (d.hi[0] - d.lo[0]) * (float(double(gi) / double(d.n_total)));
    float v = expf(d.gt[0] * tt + d.gt[1]);
    if (d.sigma > 0.0f) v += d.sigma * approx_normal(d.seed, gi, 0);
    t[i] = ST(tt);
    y[i] = ST(v);
  }
}

template <typename ST>
__global__ void generate_pinhole_kernel(SynthDev d, int64_t n, ST* X, ST* Y, ST* Z, ST* u, ST* v) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t gi = d.first_index + i;
    float p[3];
    for (int k = 0; k < 3; ++k) p[k] = d.lo[k] + (d.hi[k] - d.lo[k]) * u01(d.seed, gi, k);
    float w[3];
    for (int k = 0; k < 3; ++k) w[k] = d.gt[k * 4 + 0] * p[0] + d.gt[k * 4 + 1] * p[1] + d.gt[k * 4 + 2] * p[2] + d.gt[k * 4 + 3];
    float uu = w[0] / w[2], vv = w[1] / w[2];
    if (d.distort) {  // same model as PinholeDistortModel::residual
      const float xn = uu, yn = vv, r2 = xn * xn + yn * yn;
      const float radial = 1.0f + r2 * (d.gt[16] + r2 * (d.gt[17] + r2 * d.gt[20]));
      const float xy2 = 2.0f * xn * yn;
      const float xd = xn * radial + d.gt[18] * xy2 + d.gt[19] * (r2 + 2.0f * xn * xn);
      const float yd = yn * radial + d.gt[18] * (r2 + 2.0f * yn * yn) + d.gt[19] * xy2;
      uu = d.gt[12] * xd + d.gt[14];
      vv = d.gt[13] * yd + d.gt[15];
    }
    if (d.sigma > 0.0f) {
      uu += d.sigma * approx_normal(d.seed, gi, 3);
      vv += d.sigma * approx_normal(d.seed, gi, 7);
    }
    X[i] = ST(p[0]); Y[i] = ST(p[1]); Z[i] = ST(p[2]);
    u[i] = ST(uu); v[i] = ST(vv);
  }
}

size_t dtype_size(int dt) { return dt == MOPT_F32 ? 4 : 8; }

void rodrigues_host(const double* w, double* R) {
  const double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if (n > 0.0) {
    const double a[3] = {w[0] / n, w[1] / n, w[2] / n};
    const double K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    const double s = std::sin(n), c = std::cos(n);
    for (int r = 0; r < 3; ++r)
      for (int col = 0; col < 3; ++col) {
        double kk = 0;
        for (int k = 0; k < 3; ++k) kk += K[r * 3 + k] * K[k * 3 + col];
        R[r * 3 + col] += s * K[r * 3 + col] + (1.0 - c) * kk;
      }
  }
}

int ensure_stage(mopt_ctx* ctx, size_t bytes) {
  if (ctx->stage_bytes >= bytes) return MOPT_OK;
  for (int i = 0; i < 2; ++i) {
    if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
    ctx->d_stage[i] = nullptr;
  }
  ctx->stage_bytes = 0;
  for (int i = 0; i < 2; ++i) MOPT_CUDA_TRY(cudaMalloc(&ctx->d_stage[i], bytes));
  ctx->stage_bytes = bytes;
  return MOPT_OK;
}

template <typename HT, typename ST>
void launch_deinterleave(cudaStream_t s, const void* aos, int64_t stride, int ncomp, void* const* streams, int64_t first,
                         int64_t count) {
  const int threads = 256;
  int64_t blocks = (count + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  deinterleave_kernel<HT, ST><<<int(blocks), threads, 0, s>>>(static_cast<const HT*>(aos), stride, ncomp,
                                                             static_cast<ST*>(streams[0]), static_cast<ST*>(streams[1]),
                                                             static_cast<ST*>(streams[2]), first, count);
}

template <typename HT, typename ST>
void launch_interleave(cudaStream_t s, void* aos, int ncomp, void* const* streams, int64_t first, int64_t count) {
  const int threads = 256;
  int64_t blocks = (count + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  interleave_kernel<HT, ST><<<int(blocks), threads, 0, s>>>(static_cast<HT*>(aos), ncomp, static_cast<const ST*>(streams[0]),
                                                           static_cast<const ST*>(streams[1]),
                                                           static_cast<const ST*>(streams[2]), first, count);
}

}  // namespace

namespace mopt {

// Enqueue the chunked, double-buffered upload of one data group; the caller synchronises.
// copy_stream: H2D of chunk c+1 overlaps the de-interleave of chunk c on ctx->stream.
int store_upload_async(mopt_store* st, int group, const void* host, int host_dtype, int64_t host_stride, int64_t first,
                       int64_t count) {
  mopt_ctx* ctx = st->ctx;
  const ModelShape sh = model_shape(st->model);
  MOPT_REQUIRE(group == 0 || group == 1, "group must be 0 (A) or 1 (B)");
  const int ncomp = group == 0 ? sh.ncomp_a : sh.ncomp_b;
  MOPT_REQUIRE(ncomp > 0, "this model has no such data group");
  MOPT_REQUIRE(first >= 0 && count >= 0 && first + count <= st->n, "upload range outside the store");
  MOPT_REQUIRE(host_dtype == MOPT_F32 || host_dtype == MOPT_F64, "bad host dtype");
  if (count == 0) return MOPT_OK;
  MOPT_REQUIRE(host != nullptr, "null host pointer");
  if (host_stride == 0) host_stride = ncomp;
  MOPT_REQUIRE(host_stride >= ncomp, "host stride smaller than the component count");
  void* const* streams = st->streams + (group == 0 ? 0 : sh.NA);
  const size_t hsz = dtype_size(host_dtype);
  const size_t chunk_bytes_target = size_t(64) << 20;
  int64_t chunk_elems = int64_t(chunk_bytes_target / (hsz * size_t(host_stride)));
  if (chunk_elems < 1) chunk_elems = 1;
  if (chunk_elems > count) chunk_elems = count;
  MOPT_TRY(ensure_stage(ctx, size_t(chunk_elems) * size_t(host_stride) * hsz));
  int buf = 0;
  for (int64_t off = 0; off < count; off += chunk_elems, buf ^= 1) {
    const int64_t m = (count - off < chunk_elems) ? (count - off) : chunk_elems;
    // the last element of a strided chunk may not have a full stride behind it
    const size_t bytes = (size_t(m - 1) * size_t(host_stride) + size_t(ncomp)) * hsz;
    MOPT_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[buf], 0));
    MOPT_CUDA_TRY(cudaMemcpyAsync(ctx->d_stage[buf], static_cast<const char*>(host) + size_t(off) * size_t(host_stride) * hsz,
                                  bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    MOPT_CUDA_TRY(cudaEventRecord(ctx->ev_copy[buf], ctx->copy_stream));
    MOPT_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[buf], 0));
    if (host_dtype == MOPT_F32 && st->dtype == MOPT_F32)
      launch_deinterleave<float, float>(ctx->stream, ctx->d_stage[buf], host_stride, ncomp, streams, first + off, m);
    else if (host_dtype == MOPT_F64 && st->dtype == MOPT_F32)
      launch_deinterleave<double, float>(ctx->stream, ctx->d_stage[buf], host_stride, ncomp, streams, first + off, m);
    else if (host_dtype == MOPT_F32 && st->dtype == MOPT_F64)
      launch_deinterleave<float, double>(ctx->stream, ctx->d_stage[buf], host_stride, ncomp, streams, first + off, m);
    else
      launch_deinterleave<double, double>(ctx->stream, ctx->d_stage[buf], host_stride, ncomp, streams, first + off, m);
    MOPT_CUDA_TRY(cudaGetLastError());
    MOPT_CUDA_TRY(cudaEventRecord(ctx->ev_free[buf], ctx->stream));
  }
  return MOPT_OK;
}

}  // namespace mopt

extern "C" {

int mopt_store_create(mopt_ctx* ctx, int model, int dtype, int64_t n, mopt_store** out) try {
  MOPT_REQUIRE(ctx && out, "null ctx/out");
  const ModelShape sh = model_shape(model);
  MOPT_REQUIRE(sh.P >= 0, "unknown model kind");
  MOPT_REQUIRE(dtype == MOPT_F32 || dtype == MOPT_F64, "store dtype must be MOPT_F32 or MOPT_F64");
  MOPT_REQUIRE(n >= 0, "negative store size");
  MOPT_CUDA_TRY(cudaSetDevice(ctx->device));
  mopt_store* st = new mopt_store();
  st->ctx = ctx;
  st->model = model;
  st->dtype = dtype;
  st->n = n;
  st->nstreams = sh.NS;
  // pad every stream to a whole number of 16-byte vectors so vector loads never straddle the end
  const size_t bytes = ((size_t(n) * dtype_size(dtype) + 255) / 256) * 256 + 256;
  for (int s = 0; s < sh.NS; ++s) {
    cudaError_t e = cudaMalloc(&st->streams[s], bytes);
    if (e != cudaSuccess) {
      set_last_error(std::string("cudaMalloc of a store stream failed: ") + cudaGetErrorString(e));
      for (int k = 0; k < s; ++k) cudaFree(st->streams[k]);
      delete st;
      return e == cudaErrorMemoryAllocation ? MOPT_ERR_OUT_OF_MEMORY : MOPT_ERR_CUDA;
    }
  }
  *out = st;
  return MOPT_OK;
}
MOPT_ABI_CATCH

int mopt_store_destroy(mopt_store* st) try {
  if (!st) return MOPT_OK;
  cudaSetDevice(st->ctx->device);
  cudaStreamSynchronize(st->ctx->stream);
  for (int s = 0; s < st->nstreams; ++s) cudaFree(st->streams[s]);
  delete st;
  return MOPT_OK;
}
MOPT_ABI_CATCH

int mopt_store_size(const mopt_store* st, int64_t* n) try {
  MOPT_REQUIRE(st && n, "null store/n");
  *n = st->n;
  return MOPT_OK;
}
MOPT_ABI_CATCH

int mopt_store_upload(mopt_store* st, int group, const void* host, int host_dtype, int64_t host_stride, int64_t first,
                      int64_t count) try {
  MOPT_REQUIRE(st, "null store");
  MOPT_CUDA_TRY(cudaSetDevice(st->ctx->device));
  MOPT_TRY(store_upload_async(st, group, host, host_dtype, host_stride, first, count));
  MOPT_CUDA_TRY(cudaStreamSynchronize(st->ctx->stream));
  return MOPT_OK;
}
MOPT_ABI_CATCH

int mopt_store_download(mopt_store* st, int group, void* host, int host_dtype, int64_t first, int64_t count) try {
  MOPT_REQUIRE(st && host, "null store/host");
  mopt_ctx* ctx = st->ctx;
  MOPT_CUDA_TRY(cudaSetDevice(ctx->device));
  const ModelShape sh = model_shape(st->model);
  MOPT_REQUIRE(group == 0 || group == 1, "group must be 0 (A) or 1 (B)");
  const int ncomp = group == 0 ? sh.ncomp_a : sh.ncomp_b;
  MOPT_REQUIRE(ncomp > 0, "this model has no such data group");
  MOPT_REQUIRE(first >= 0 && count >= 0 && first + count <= st->n, "download range outside the store");
  MOPT_REQUIRE(host_dtype == MOPT_F32 || host_dtype == MOPT_F64, "bad host dtype");
  void* const* streams = st->streams + (group == 0 ? 0 : sh.NA);
  const size_t hsz = dtype_size(host_dtype);
  const int64_t chunk_elems_max = int64_t((size_t(64) << 20) / (hsz * ncomp));
  const int64_t chunk = count < chunk_elems_max ? count : chunk_elems_max;
  if (count == 0) return MOPT_OK;
  MOPT_TRY(ensure_stage(ctx, size_t(chunk) * ncomp * hsz));
  for (int64_t off = 0; off < count; off += chunk) {
    const int64_t m = (count - off < chunk) ? (count - off) : chunk;
    if (host_dtype == MOPT_F32 && st->dtype == MOPT_F32)
      launch_interleave<float, float>(ctx->stream, ctx->d_stage[0], ncomp, streams, first + off, m);
    else if (host_dtype == MOPT_F64 && st->dtype == MOPT_F32)
      launch_interleave<double, float>(ctx->stream, ctx->d_stage[0], ncomp, streams, first + off, m);
    else if (host_dtype == MOPT_F32 && st->dtype == MOPT_F64)
      launch_interleave<float, double>(ctx->stream, ctx->d_stage[0], ncomp, streams, first + off, m);
    else
      launch_interleave<double, double>(ctx->stream, ctx->d_stage[0], ncomp, streams, first + off, m);
    MOPT_CUDA_TRY(cudaGetLastError());
    MOPT_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(host) + size_t(off) * ncomp * hsz, ctx->d_stage[0],
                                  size_t(m) * ncomp * hsz, cudaMemcpyDeviceToHost, ctx->stream));
    MOPT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  return MOPT_OK;
}
MOPT_ABI_CATCH

int mopt_store_generate(mopt_store* st, const mopt_synth* desc) try {
  MOPT_REQUIRE(st && desc, "null store/desc");
  mopt_ctx* ctx = st->ctx;
  MOPT_CUDA_TRY(cudaSetDevice(ctx->device));
  SynthDev d;
  std::memset(&d, 0, sizeof(d));
  d.seed = desc->seed;
  d.first_index = desc->first_index;
  d.n_total = desc->n_total > 0 ? desc->n_total : st->n;
  for (int k = 0; k < 3; ++k) {
    d.lo[k] = float(desc->lo[k]);
    d.hi[k] = float(desc->hi[k]);
  }
  d.sigma = float(desc->noise_sigma);
  d.outlier_fraction = float(desc->outlier_fraction);
  d.outlier_range = float(desc->outlier_range);
  const int threads = 256;
  int64_t blocks = (st->n + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  const bool f32 = st->dtype == MOPT_F32;
  switch (st->model) {
    case MOPT_MODEL_POINT2POINT:
    case MOPT_MODEL_POINT_DIST: {
      double R[9];
      rodrigues_host(desc->gt + 3, R);
      for (int i = 0; i < 9; ++i) d.gt[i] = float(R[i]);
      for (int i = 0; i < 3; ++i) d.gt[9 + i] = float(desc->gt[i]);
      if (f32)
        generate_p2p_kernel<float><<<int(blocks), threads, 0, ctx->stream>>>(
            d, st->n, (float*)st->streams[0], (float*)st->streams[1], (float*)st->streams[2], (float*)st->streams[3],
            (float*)st->streams[4], (float*)st->streams[5]);
      else
        generate_p2p_kernel<double><<<int(blocks), threads, 0, ctx->stream>>>(
            d, st->n, (double*)st->streams[0], (double*)st->streams[1], (double*)st->streams[2], (double*)st->streams[3],
            (double*)st->streams[4], (double*)st->streams[5]);
      break;
    }
    case MOPT_MODEL_EXP_CURVE: {
      d.gt[0] = float(desc->gt[0]);
      d.gt[1] = float(desc->gt[1]);
      if (f32)
        generate_curve_kernel<float><<<int(blocks), threads, 0, ctx->stream>>>(d, st->n, (float*)st->streams[0],
                                                                               (float*)st->streams[1]);
      else
        generate_curve_kernel<double><<<int(blocks), threads, 0, ctx->stream>>>(d, st->n, (double*)st->streams[0],
                                                                                (double*)st->streams[1]);
      break;
    }
    case MOPT_MODEL_PINHOLE:
    case MOPT_MODEL_PINHOLE_DISTORT: {
      if (st->model == MOPT_MODEL_PINHOLE) {
        // gt[0..12) = projection matrix M = K T(x_gt) C, supplied by the caller (3x4 row-major)
        for (int i = 0; i < 12; ++i) d.gt[i] = float(desc->gt[i]);
      } else {
        // gt = the 15 generating parameters [t, omega, fx, fy, cx, cy, k1, k2, p1, p2, k3]; consts = C (4x4)
        double R[9];
        rodrigues_host(desc->gt + 3, R);
        for (int r = 0; r < 3; ++r)
          for (int col = 0; col < 4; ++col) {
            double s = desc->gt[r] * desc->consts[12 + col];
            for (int k = 0; k < 3; ++k) s += R[r * 3 + k] * desc->consts[k * 4 + col];
            d.gt[r * 4 + col] = float(s);
          }
        for (int i = 0; i < 9; ++i) d.gt[12 + i] = float(desc->gt[6 + i]);
        d.distort = 1;
      }
      if (f32)
        generate_pinhole_kernel<float><<<int(blocks), threads, 0, ctx->stream>>>(
            d, st->n, (float*)st->streams[0], (float*)st->streams[1], (float*)st->streams[2], (float*)st->streams[3],
            (float*)st->streams[4]);
      else
        generate_pinhole_kernel<double><<<int(blocks), threads, 0, ctx->stream>>>(
            d, st->n, (double*)st->streams[0], (double*)st->streams[1], (double*)st->streams[2], (double*)st->streams[3],
            (double*)st->streams[4]);
      break;
    }
    default:
      set_last_error("no synthetic generator for this model");
      return MOPT_ERR_UNSUPPORTED;
  }
  MOPT_CUDA_TRY(cudaGetLastError());
  MOPT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MOPT_OK;
}
MOPT_ABI_CATCH

}  // extern "C"
