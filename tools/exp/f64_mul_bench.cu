// Round-2 experiment: throughput of the Fp Montgomery product with limb
// products on the FP64 pipe (tools/exp/mont_f64.cuh) against the shipped IMAD.WIDE product
// (csrc/mont.cuh), register resident, with a bit-exact cross-check of the two.  Standalone.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr \
//        -I zukelang_b200/csrc -o tools/exp/f64_mul_bench tools/exp/f64_mul_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include "ec.cuh"
#include "mont_f64.cuh"

using f64mont::Limbs;

__device__ __forceinline__ Limbs to48(const Fp& a) {       // 12 x 32 -> 8 x 48 (same integer)
  Limbs r;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int bit = 48 * k, w = bit >> 5, s = bit & 31;    // s is 0 or 16
    uint64_t lo = a.v[w], hi = a.v[w + 1], top = (w + 2 < 12) ? a.v[w + 2] : 0;
    uint64_t x = (lo | (hi << 32)) >> s;
    if (s) x |= top << (64 - s);
    r.v[k] = x & f64mont::MASK48;
  }
  return r;
}
__device__ __forceinline__ Fp from48(const Limbs& a) {
  Fp r;
#pragma unroll
  for (int w = 0; w < 12; w++) {
    const int bit = 32 * w, k = bit / 48, s = bit % 48;    // s in {0, 32, 16}
    uint64_t x = a.v[k] >> s;
    if (s > 16 && k + 1 < 8) x |= a.v[k + 1] << (48 - s);
    r.v[w] = (uint32_t)x;
  }
  return r;
}

__device__ Limbs modulus48() {
  Fp p;
#pragma unroll
  for (int i = 0; i < 12; i++) p.v[i] = FpParams::mod(i);
  return to48(p);
}

// mode 0: IMAD product chains, mode 1: FP64-pipe product chains, mode 2: one chain on each pipe
// (the compiler interleaves the two straight-line instruction streams); two independent chains per thread
template <int MODE>
__global__ void __launch_bounds__(128) k_mul(int iters, uint64_t n0inv48, uint32_t* out) {
  Fp x, y;
#pragma unroll
  for (int i = 0; i < 12; i++) { x.v[i] = FpParams::g1x(i) ^ (threadIdx.x * 0x9e37u & 0xffff); y.v[i] = FpParams::g1y(i); }
  x.v[11] &= 0x0fffffff;                                   // keep < p
  Fp a = x, b = y;
  if (MODE == 0) {
    for (int it = 0; it < iters; it++) { a = a * y; b = b * x; }
  } else if (MODE == 1) {
    const Limbs p = modulus48();
    Limbs A = to48(a), B = to48(b), X = to48(x), Y = to48(y);
    for (int it = 0; it < iters; it++) { A = f64mont::mul(A, Y, p, n0inv48); B = f64mont::mul(B, X, p, n0inv48); }
    a = from48(A);
    b = from48(B);
  } else {
    const Limbs p = modulus48();
    Limbs B = to48(b), X = to48(x);
    for (int it = 0; it < iters; it++) { a = a * y; B = f64mont::mul(B, X, p, n0inv48); }
    b = from48(B);
  }
  Fp s = a + b;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 12; i++) out[t * 12 + i] = s.v[i];
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
  // n0' = -p^-1 mod 2^48 (Newton on the low limb)
  const uint64_t p0 = 0xb9feffffffffaaabULL;
  uint64_t inv = 1;
  for (int i = 0; i < 6; i++) inv *= 2 - p0 * inv;
  const uint64_t n0inv48 = (0 - inv) & 0xffffffffffffULL;
  const int sms = prop.multiProcessorCount, blocks = sms * 4, threads = 128, iters = 2000;
  const size_t n = (size_t)blocks * threads * 12;
  uint32_t *d0, *d1, *d2;
  cudaMalloc(&d0, n * 4);
  cudaMalloc(&d1, n * 4);
  cudaMalloc(&d2, n * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms[3];
  for (int mode = 0; mode < 3; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k_mul<0><<<blocks, threads>>>(iters, n0inv48, d0);
      else if (mode == 1) k_mul<1><<<blocks, threads>>>(iters, n0inv48, d1);
      else k_mul<2><<<blocks, threads>>>(iters, n0inv48, d2);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms[mode], e0, e1);
    }
  }
  uint32_t* h0 = (uint32_t*)malloc(n * 4);
  uint32_t* h1 = (uint32_t*)malloc(n * 4);
  uint32_t* h2 = (uint32_t*)malloc(n * 4);
  cudaMemcpy(h0, d0, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h1, d1, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h2, d2, n * 4, cudaMemcpyDeviceToHost);
  size_t bad = 0;
  for (size_t i = 0; i < n; i++) bad += (h0[i] != h1[i]) + (h0[i] != h2[i]);
  const double muls = 2.0 * iters * blocks * threads;
  printf("{\"device\": \"%s\", \"imad_gmul_s\": %.2f, \"f64_gmul_s\": %.2f, \"one_on_each_pipe_gmul_s\": %.2f, "
         "\"imad_ms\": %.3f, \"f64_ms\": %.3f, \"hybrid_ms\": %.3f, \"mismatching_words\": %zu}\n", prop.name,
         muls / ms[0] / 1e6, muls / ms[1] / 1e6, muls / ms[2] / 1e6, ms[0], ms[1], ms[2], bad);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { fprintf(stderr, "cuda error: %s\n", cudaGetErrorString(e)); return 2; }
  return bad ? 3 : 0;
}
