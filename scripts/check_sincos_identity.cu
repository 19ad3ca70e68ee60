// Is CUDA's double sincos(x) bit-identical to sin(x) and cos(x) taken separately?  (Design input for sharing one
// sincos between so3::Exp and the left Jacobian in setup: the two must stay bit-identical to the current code.)
// Build + run on a GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/check_sincos_identity.cu -o /tmp/sc && /tmp/sc
#include <cstdio>
#include <cstdint>
__global__ void k(unsigned long long* bad, unsigned long long n) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long cnt = 0;
  for (; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long h = i * 0x9E3779B97F4A7C15ull; h ^= h >> 31; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29;
    // magnitudes from 1e-9 to ~10 (rotation vectors), both signs
    const double u = double(h >> 11) * (1.0 / 9007199254740992.0);
    const double mag = exp(-20.0 * u) * 10.0;
    const double x = (h & 1) ? -mag : mag;
    double s, c;
    sincos(x, &s, &c);
    if (__double_as_longlong(s) != __double_as_longlong(sin(x)) || __double_as_longlong(c) != __double_as_longlong(cos(x))) ++cnt;
  }
  if (cnt) atomicAdd(bad, cnt);
}
int main() {
  unsigned long long* d; unsigned long long h = 0;
  cudaMalloc(&d, 8); cudaMemset(d, 0, 8);
  const unsigned long long n = 1ull << 30;
  k<<<148 * 8, 256>>>(d, n);
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  std::printf("sincos vs sin/cos: %llu mismatches in %llu samples\n", h, n);
  return h ? 1 : 0;
}
