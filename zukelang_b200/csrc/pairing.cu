// Verifier-side entry points of the C ABI: pairing products, the GT product
// and point decompression.  Formulas in pairing.cuh; this file is the kernels and the glue.
//
//   zk_pairing_product  <- Pairing.pairing / GT.( + ) / GT.( - ) as used by
//                          /root/reference/src/groth16/groth16.ml:103,163-173 and
//                          src/pinocchio/pinocchio.ml:254-420
//   zk_gt_mul           <- GT.( + ) (the GT product; curve.ml:212-220 wraps GT with ExtendG)
//   zk_g{1,2}_decompress<- G1/G2.of_compressed_bytes_exn (curve.ml:201,210), the reader of the
//                          yojson wire format
//
// One pairing is a few hundred thousand dependent Fp products: latency code.  The independent
// pieces of a call — per pair the Miller loop and the two subgroup checks — run as separate
// single-thread blocks so that they overlap across SMs; a last single-thread kernel multiplies the
// Miller values and runs the one final exponentiation.
#include "runtime.cuh"
#include "pairing.cuh"

namespace zk {

static __device__ void gt_serialize(const Fp12& f, uint8_t* out) {
  const Fp* c = reinterpret_cast<const Fp*>(&f);
  for (int k = 0; k < 12; k++) fp_to_be(c[k].from_mont(), out + 48 * k);
}
static __device__ bool gt_parse(const uint8_t* in, Fp12& f) {
  Fp* c = reinterpret_cast<Fp*>(&f);
  bool ok = true;
  for (int k = 0; k < 12; k++) {
    Fp raw;
    ok = fp_from_be(in + 48 * k, raw) && ok;
    c[k] = raw.to_mont();
  }
  return ok;
}

// blockIdx.x = pair, blockIdx.y = role (0 Miller loop, 1 G1 subgroup check, 2 G2 subgroup check)
static __global__ void __launch_bounds__(32)
k_pairing_miller(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2, const uint8_t* __restrict__ negate,
                 uint32_t n, Fp12* __restrict__ out, int* __restrict__ err) {
  const uint32_t i = blockIdx.x;
  if (i >= n || threadIdx.x != 0) return;
  const int role = blockIdx.y;
  Affine<Fp> p = Affine<Fp>::inf();
  Affine<Fp2> q = Affine<Fp2>::inf();
  if (role != 2 && G1Traits::parse(g1 + (size_t)i * G1Traits::RAW, p)) { if (role == 0) out[i] = Fp12::one(); atomicMax(err, 1); return; }
  if (role != 1 && G2Traits::parse(g2 + (size_t)i * G2Traits::RAW, q)) { if (role == 0) out[i] = Fp12::one(); atomicMax(err, 1); return; }
  if (role == 1) { if (!in_prime_subgroup(p)) atomicMax(err, 2); return; }
  if (role == 2) { if (!in_prime_subgroup(q)) atomicMax(err, 2); return; }
  if (p.is_inf() || q.is_inf()) { out[i] = Fp12::one(); return; }   // e(O, Q) = e(P, O) = 1
  if (negate && negate[i]) p.y = p.y.neg();
  out[i] = miller_loop(p, q);
}

// blockIdx.x = product g: multiplies the Miller values of pairs [first[g], first[g+1]) and runs the
// final exponentiation of that product (an empty range gives 1)
static __global__ void __launch_bounds__(32)
k_pairing_final(const Fp12* __restrict__ f, const uint32_t* __restrict__ first, uint8_t* __restrict__ out) {
  if (threadIdx.x != 0) return;
  const uint32_t g = blockIdx.x, lo = first[g], hi = first[g + 1];
  Fp12 acc = Fp12::one();
  if (lo < hi) {
    acc = f[lo];
    for (uint32_t i = lo + 1; i < hi; i++) acc = f12_mul(acc, f[i]);
    acc = final_exponentiation(acc);
  }
  gt_serialize(acc, out + (size_t)g * ZK_GT_BYTES);
}

static __global__ void __launch_bounds__(32)
k_gt_mul(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out, int* __restrict__ err) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Fp12 A, B;
  if (!gt_parse(a, A) || !gt_parse(b, B)) { *err = 1; return; }
  gt_serialize(f12_mul(A, B), out);
}

// ---- decompression ---------------------------------------------------------------------------------
// zcash / blst flags in the top three bits of byte 0: 0x80 compressed, 0x40 identity, 0x20 "y is the
// lexicographically larger root".  Returns 0 ok, 1 bad encoding, 2 no such point, 3 outside the subgroup.
static __device__ int decompress_one(const uint8_t* in, Affine<Fp>& out) {
  const uint8_t fl = in[0];
  if (!(fl & 0x80)) return 1;
  if (fl & 0x40) {
    bool ok = fl == 0xc0;
    for (int i = 1; i < 48; i++) ok = ok && in[i] == 0;
    out = Affine<Fp>::inf();
    return ok ? 0 : 1;
  }
  Fp x;
  if (!fp_from_be(in, x, 0x1f)) return 1;
  out.x = x.to_mont();
  Fp y2 = Fp::mul_call(Fp::mul_call(out.x, out.x), out.x) + G1Traits::curve_b();
  if (!fp_sqrt(y2, out.y)) return 2;
  if (fp_gt_half(out.y.from_mont()) != ((fl & 0x20) != 0)) out.y = out.y.neg();
  return in_prime_subgroup(out) ? 0 : 3;
}
static __device__ int decompress_one(const uint8_t* in, Affine<Fp2>& out) {
  const uint8_t fl = in[0];
  if (!(fl & 0x80)) return 1;
  if (fl & 0x40) {
    bool ok = fl == 0xc0;
    for (int i = 1; i < 96; i++) ok = ok && in[i] == 0;
    out = Affine<Fp2>::inf();
    return ok ? 0 : 1;
  }
  Fp x1, x0;  // wire order: x.c1 (with the flags) | x.c0
  if (!fp_from_be(in, x1, 0x1f) || !fp_from_be(in + 48, x0)) return 1;
  out.x = Fp2{x0.to_mont(), x1.to_mont()};
  Fp2 y2 = f2_mul(f2_sqr(out.x), out.x) + G2Traits::curve_b();
  if (!fp2_sqrt(y2, out.y)) return 2;
  const Fp y1 = out.y.c1.from_mont();
  const bool big = y1.is_zero() ? fp_gt_half(out.y.c0.from_mont()) : fp_gt_half(y1);   // c1 first, then c0
  if (big != ((fl & 0x20) != 0)) out.y = out.y.neg();
  return in_prime_subgroup(out) ? 0 : 3;
}

template <class T>
__global__ void __launch_bounds__(32)
k_decompress(const uint8_t* __restrict__ in, uint32_t n, uint8_t* __restrict__ out, int* __restrict__ err) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<typename T::F> p;
  const int rc = decompress_one(in + (size_t)i * T::COMP, p);
  if (rc) { atomicMax(err, rc); return; }
  uint8_t buf[T::RAW + T::COMP];
  T::serialize(p, buf);
  for (int j = 0; j < T::RAW; j++) out[(size_t)i * T::RAW + j] = buf[j];
}

template <class T>
static int api_decompress(const uint8_t* in, size_t n, uint8_t* out) {
  ZK_API_BEGIN
  ZK_REQUIRE(in && out && n > 0 && n <= (1u << 24), ZK_EARG, "decompress: bad arguments");
  cudaStream_t st = default_stream();
  DevBuf<uint8_t> d_in(n * T::COMP), d_out(n * T::RAW);
  DevBuf<int> d_err(1);
  ZK_CUDA(cudaMemcpyAsync(d_in.p, in, n * T::COMP, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
  k_decompress<T><<<cdiv(n, 32), 32, 0, st>>>(d_in.p, (uint32_t)n, d_out.p, d_err.p);
  ZK_CUDA(cudaGetLastError());
  int err = 0;
  ZK_CUDA(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaMemcpyAsync(out, d_out.p, n * T::RAW, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(err != 1, ZK_EPOINT, "decompress: bad encoding (flags or non-canonical coordinate)");
  ZK_REQUIRE(err != 2, ZK_EPOINT, "decompress: x is not the abscissa of a curve point");
  ZK_REQUIRE(err != 3, ZK_EPOINT, "decompress: point is not in the prime-order subgroup");
  ZK_API_END
}

}  // namespace zk

using namespace zk;

extern "C" {

// k products over consecutive runs of pairs: counts[g] pairs feed product g
static int pairing_products(const uint8_t* g1, const uint8_t* g2, const uint8_t* negate, const uint32_t* counts,
                            size_t k, uint8_t* out) {
  ZK_API_BEGIN
  ZK_REQUIRE(g1 && g2 && out && counts && k > 0 && k <= 1024, ZK_EARG, "pairing_product: bad arguments");
  std::vector<uint32_t> first(k + 1, 0);
  for (size_t g = 0; g < k; g++) first[g + 1] = first[g] + counts[g];
  const size_t n = first[k];
  ZK_REQUIRE(n > 0 && n <= 4096, ZK_EARG, "pairing_product: bad pair count");
  cudaStream_t st = default_stream();
  DevBuf<uint8_t> d_g1(n * G1Traits::RAW), d_g2(n * G2Traits::RAW), d_neg(n), d_out(k * ZK_GT_BYTES);
  DevBuf<Fp12> d_f(n);
  DevBuf<uint32_t> d_first(k + 1);
  DevBuf<int> d_err(1);
  ZK_CUDA(cudaMemcpyAsync(d_g1.p, g1, n * G1Traits::RAW, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(d_g2.p, g2, n * G2Traits::RAW, cudaMemcpyHostToDevice, st));
  if (negate) ZK_CUDA(cudaMemcpyAsync(d_neg.p, negate, n, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(d_first.p, first.data(), (k + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
  k_pairing_miller<<<dim3((unsigned)n, 3), 32, 0, st>>>(d_g1.p, d_g2.p, negate ? d_neg.p : nullptr, (uint32_t)n, d_f.p, d_err.p);
  ZK_CUDA(cudaGetLastError());
  k_pairing_final<<<(unsigned)k, 32, 0, st>>>(d_f.p, d_first.p, d_out.p);
  ZK_CUDA(cudaGetLastError());
  int err = 0;
  ZK_CUDA(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaMemcpyAsync(out, d_out.p, k * ZK_GT_BYTES, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));   // first[] and the outputs are done with
  ZK_REQUIRE(err != 1, ZK_EPOINT, "pairing_product: point not canonical or not on the curve");
  ZK_REQUIRE(err != 2, ZK_EPOINT, "pairing_product: point is not in the prime-order subgroup");
  ZK_API_END
}

int zk_pairing_product(const uint8_t* g1, const uint8_t* g2, const uint8_t* negate, size_t n, uint8_t* out) {
  const uint32_t count = (uint32_t)n;
  if (n == 0 || n > 4096) { zk::set_error("pairing_product: bad arguments"); return ZK_EARG; }
  return pairing_products(g1, g2, negate, &count, 1, out);
}

int zk_pairing_product_batch(const uint8_t* g1, const uint8_t* g2, const uint8_t* negate, const uint32_t* counts,
                             size_t k, uint8_t* out) {
  return pairing_products(g1, g2, negate, counts, k, out);
}

int zk_gt_mul(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  ZK_API_BEGIN
  ZK_REQUIRE(a && b && out, ZK_EARG, "gt_mul: bad arguments");
  cudaStream_t st = default_stream();
  DevBuf<uint8_t> d_a(ZK_GT_BYTES), d_b(ZK_GT_BYTES), d_out(ZK_GT_BYTES);
  DevBuf<int> d_err(1);
  ZK_CUDA(cudaMemcpyAsync(d_a.p, a, ZK_GT_BYTES, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(d_b.p, b, ZK_GT_BYTES, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
  k_gt_mul<<<1, 32, 0, st>>>(d_a.p, d_b.p, d_out.p, d_err.p);
  ZK_CUDA(cudaGetLastError());
  int err = 0;
  ZK_CUDA(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaMemcpyAsync(out, d_out.p, ZK_GT_BYTES, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(err == 0, ZK_EPOINT, "gt_mul: coefficient not canonical");
  ZK_API_END
}

int zk_g1_decompress(const uint8_t* comp, size_t n, uint8_t* out) { return api_decompress<G1Traits>(comp, n, out); }
int zk_g2_decompress(const uint8_t* comp, size_t n, uint8_t* out) { return api_decompress<G2Traits>(comp, n, out); }

}  // extern "C"
