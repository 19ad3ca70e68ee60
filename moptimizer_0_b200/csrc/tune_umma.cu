// Stand-alone check of the tcgen05 building block worked out in DESIGN.md §8.2: the Gram matrix X^T X of a 16-column
// tile (K = 512 rows) by tcgen05.mma.kind::tf32 M64 N16 K8 from the canonical K-major no-swizzle shared-memory layout,
// accumulators in TMEM, read back with tcgen05.ld.  Prints the worst error against a double-precision product of the
// TF32-rounded operands, the TMEM row -> lane map it found, and cycles per MMA.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o moptimizer_0_b200/tune_umma moptimizer_0_b200/csrc/tune_umma.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int P = 16;        // columns of X (parameters)
constexpr int K = 512;       // rows of X
constexpr int LBO = 272;     // bytes between the two 16-byte K chunks of an instruction (and between consecutive chunks)
constexpr int SBO = 128;     // bytes between 8-row groups
constexpr int NCHUNK = K / 4;
constexpr int XBYTES = NCHUNK * LBO + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3fff);
  d |= uint64_t((LBO >> 4) & 0x3fff) << 16;
  d |= uint64_t((SBO >> 4) & 0x3fff) << 32;
  d |= uint64_t(1) << 46;  // descriptor version (sm_100)
  return d;                // base offset 0, lbo mode 0, layout type 0 = no swizzle
}

template <int nacc>
__global__ void __launch_bounds__(128, 1) umma_gram(const float* __restrict__ X /* [K][P] row-major */, float* out /* [128][32] */,
                                                    long long* cycles, int reps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* xs = smem;  // canonical layout of X^T: element (p, k) at (k/4)*LBO + (p/8)*SBO + (p%8)*16 + (k%4)*4
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < XBYTES / 4; i += 128) reinterpret_cast<float*>(xs)[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < K * P; i += 128) {
    const int k = i / P, p = i % P;
    float v = X[i];
    // round to TF32 (10-bit mantissa, nearest): what the tensor core would otherwise truncate
    uint32_t b = __float_as_uint(v);
    b = (b + 0x1000u) & 0xffffe000u;
    *reinterpret_cast<float*>(xs + (k / 4) * LBO + (p / 8) * SBO + (p % 8) * 16 + (k % 4) * 4) = __uint_as_float(b);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 16, M = 64
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((64u >> 4) << 24);
  long long t0 = 0, t1 = 0;
  uint32_t phase = 0;
  for (int r = 0; r < reps; ++r) {
    if (tid == 0) {
      t0 = clock64();
      const uint64_t d0 = make_desc(smem_u32(xs));
#pragma unroll
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t da = d0 + uint64_t(ks * ((2 * LBO) >> 4));  // the start-address field advances by one K-step
        const uint32_t acc = ks >= nacc ? 1u : 0u;  // the first MMA into an accumulator overwrites it
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem + uint32_t(ks % nacc) * 16u), "l"(da), "l"(da), "r"(idesc), "r"(acc)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
    }
    // everyone waits for the MMAs of this repetition
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(&s_bar)), "r"(phase)
            : "memory");
      }
      phase ^= 1u;
    }
    if (tid == 0) t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  // every warp reads its 32 TMEM lanes x 16 columns
  float sum[16];
  for (int c = 0; c < 16; ++c) sum[c] = 0.f;
  for (int q = 0; q < nacc; ++q) {
    uint32_t v[16];
    const uint32_t taddr = tmem + (uint32_t(warp * 32) << 16) + uint32_t(q) * 16u;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 16; ++c) sum[c] += __uint_as_float(v[c]);
  }
  for (int c = 0; c < 16; ++c) out[(warp * 32 + lane) * 16 + c] = sum[c];
  if (tid == 0) cycles[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

int main() {
  std::vector<float> X(size_t(K) * P);
  uint32_t s = 12345u;
  for (auto& x : X) {
    s = s * 1664525u + 1013904223u;
    x = float(int(s >> 8) % 2001 - 1000) / 500.f;
  }
  float *dX, *dout;
  long long* dcyc;
  cudaMalloc(&dX, X.size() * 4);
  cudaMalloc(&dout, 128 * 16 * 4);
  cudaMalloc(&dcyc, 8);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, 128 * 16 * 4);
  cudaFuncSetAttribute(umma_gram<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, XBYTES + 1024);
  cudaFuncSetAttribute(umma_gram<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, XBYTES + 1024);
  cudaFuncSetAttribute(umma_gram<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, XBYTES + 1024);
  cudaFuncSetAttribute(umma_gram<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, XBYTES + 1024);
  for (int nacc : {1, 2, 4, 8}) {
    const int reps = 20;
    if (nacc == 1) umma_gram<1><<<1, 128, XBYTES + 1024>>>(dX, dout, dcyc, reps);
    if (nacc == 2) umma_gram<2><<<1, 128, XBYTES + 1024>>>(dX, dout, dcyc, reps);
    if (nacc == 4) umma_gram<4><<<1, 128, XBYTES + 1024>>>(dX, dout, dcyc, reps);
    if (nacc == 8) umma_gram<8><<<1, 128, XBYTES + 1024>>>(dX, dout, dcyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      std::printf("CUDA error: %s\n", cudaGetErrorString(e));
      return 1;
    }
    long long cyc = 0;
    cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
    std::printf("%d accumulators round-robin: %lld cycles for %d MMAs = %.1f cycles per M64 N16 K8 tf32 MMA (issue to completion)\n", nacc,
                cyc, K / 8, double(cyc) / (K / 8));
  }
  std::vector<float> out(128 * 16);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  // reference: X_r^T X_r with X_r the TF32-rounded operands, in double
  std::vector<double> G(P * P, 0.0);
  for (int k = 0; k < K; ++k)
    for (int i = 0; i < P; ++i) {
      uint32_t bi;
      std::memcpy(&bi, &X[size_t(k) * P + i], 4);
      bi = (bi + 0x1000u) & 0xffffe000u;
      float xi;
      std::memcpy(&xi, &bi, 4);
      for (int j = 0; j < P; ++j) {
        uint32_t bj;
        std::memcpy(&bj, &X[size_t(k) * P + j], 4);
        bj = (bj + 0x1000u) & 0xffffe000u;
        float xj;
        std::memcpy(&xj, &bj, 4);
        G[i * P + j] += double(xi) * double(xj);
      }
    }
  double gmax = 0.0;
  for (double g : G) gmax = std::fmax(gmax, std::fabs(g));
  // which TMEM lane holds row i?  try the lanes and report the best match per row
  double worst = 0.0;
  for (int i = 0; i < P; ++i) {
    int best_lane = -1;
    double best = 1e300;
    for (int l = 0; l < 128; ++l) {
      double err = 0.0;
      for (int j = 0; j < P; ++j) err = std::fmax(err, std::fabs(double(out[l * 16 + j]) - G[i * P + j]));
      if (err < best) { best = err; best_lane = l; }
    }
    std::printf("row %2d -> TMEM lane %3d, max abs err %.3e (relative to max |G| %.3e)\n", i, best_lane, best, best / gmax);
    worst = std::fmax(worst, best / gmax);
  }
  std::printf("worst relative error %.3e\n", worst);
  return 0;
}
