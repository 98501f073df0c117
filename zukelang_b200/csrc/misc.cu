// Measurement and unit-test hooks: integer-pipe microbenchmarks,
// device field-op test entry, table introspection.
#include "msm_api.cuh"

namespace zk {

// ---- integer pipe microbenchmarks ----------------------------------------------
// Every kernel runs `iters` rounds of UNR independent dependent-chains per thread so the
// IMAD pipe, not latency, is the limit at 8+ warps per scheduler.
constexpr int MB_CHAINS = 8;

__global__ void __launch_bounds__(256) k_mb_mad_lo(uint32_t* out, int iters, uint32_t x, uint32_t y) {
  uint32_t a[MB_CHAINS];
#pragma unroll
  for (int j = 0; j < MB_CHAINS; j++) a[j] = threadIdx.x + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int j = 0; j < MB_CHAINS; j++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(x), "r"(y));
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < MB_CHAINS; j++) s ^= a[j];
  if (s == 0x12345678u) out[0] = s;
}

// carry chains as the Montgomery rows issue them: mad.lo.cc / madc.hi.cc pairs, 2 independent chains
__global__ void __launch_bounds__(256) k_mb_mad_cc(uint32_t* out, int iters, uint32_t x, uint32_t y) {
  uint32_t a[MB_CHAINS], b[MB_CHAINS];
#pragma unroll
  for (int j = 0; j < MB_CHAINS; j++) { a[j] = threadIdx.x + j; b[j] = threadIdx.x * 3 + j; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      asm volatile(
          "mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
          "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
          "madc.lo.cc.u32 %4, %9, %10, %4;\n\tmadc.hi.cc.u32 %5, %9, %10, %5;\n\t"
          "madc.lo.cc.u32 %6, %8, %8, %6;\n\tmadc.hi.u32 %7, %8, %8, %7;"
          : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
          : "r"(x), "r"(y), "r"(x ^ y));
      asm volatile(
          "mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
          "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
          "madc.lo.cc.u32 %4, %9, %10, %4;\n\tmadc.hi.cc.u32 %5, %9, %10, %5;\n\t"
          "madc.lo.cc.u32 %6, %8, %8, %6;\n\tmadc.hi.u32 %7, %8, %8, %7;"
          : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7])
          : "r"(y), "r"(x), "r"(x + y));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < MB_CHAINS; j++) s ^= a[j] ^ b[j];
  if (s == 0x12345678u) out[0] = s;
}

__global__ void __launch_bounds__(256) k_mb_mad_wide(uint32_t* out, int iters, uint32_t x, uint32_t y) {
  uint64_t a[MB_CHAINS];
#pragma unroll
  for (int j = 0; j < MB_CHAINS; j++) a[j] = threadIdx.x + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int j = 0; j < MB_CHAINS; j++) {
        uint32_t lo = (uint32_t)a[j];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[j]) : "r"(lo ^ x), "r"(y));
      }
  }
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < MB_CHAINS; j++) s ^= a[j];
  if (s == 0x12345678u) out[0] = (uint32_t)s;
}

template <class F>
__global__ void __launch_bounds__(128) k_mb_field_mul(uint32_t* out, int iters) {
  F a = F::one(), b = F::r2();
  a.v[0] += threadIdx.x;
  b.v[1] += blockIdx.x;
  for (int it = 0; it < iters; it++) {
    a = a * b;
    b = b * a;
  }
  if (a.v[0] == 0x12345678u && b.v[1] == 0x9abcdef0u) out[0] = a.v[2];
}

__global__ void __launch_bounds__(128) k_mb_g1_madd(uint32_t* out, int iters) {
  G1Affine g = G1Traits::generator();
  G1XYZZ acc = G1XYZZ::dbl_affine(g);
  for (int i = 0; i < (int)(threadIdx.x & 3); i++) acc = acc.dbl();
  for (int it = 0; it < iters; it++) acc.madd(g);
  if (acc.X.v[0] == 0x12345678u && acc.Y.v[1] == 0x9abcdef0u) out[0] = acc.ZZ.v[2];
}

template <class F, int MINB>
__global__ void __launch_bounds__(128, MINB) k_mb_madd_var(uint32_t* out, int iters) {
  Affine<F> g;
  {
    G1Affine gg = G1Traits::generator();
    g = *reinterpret_cast<Affine<F>*>(&gg);
  }
  XYZZ<F> acc = XYZZ<F>::dbl_affine(g);
  for (int i = 0; i < (int)(threadIdx.x & 3); i++) acc = acc.dbl();
  for (int it = 0; it < iters; it++) acc.madd(g);
  uint32_t* w = reinterpret_cast<uint32_t*>(&acc);
  if (w[0] == 0x12345678u && w[13] == 0x9abcdef0u) out[0] = w[26];
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_mb_madd_paired(uint32_t* out, int iters) {
  G1Affine g = G1Traits::generator();
  G1XYZZ acc = G1XYZZ::dbl_affine(g);
  for (int i = 0; i < (int)(threadIdx.x & 3); i++) acc = acc.dbl();
  for (int it = 0; it < iters; it++) acc.madd_paired(g);
  if (acc.X.v[0] == 0x12345678u && acc.Y.v[1] == 0x9abcdef0u) out[0] = acc.ZZ.v[2];
}

// ---- device field-op test hook ----------------------------------------------------
template <class F>
__global__ void k_test_field_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x, y, r;
  for (int j = 0; j < F::N; j++) { x.v[j] = a[(size_t)i * F::N + j]; y.v[j] = b[(size_t)i * F::N + j]; }
  switch (op) {
    case 0: r = x * y; break;
    case 1: r = x + y; break;
    case 2: r = x - y; break;
    case 3: r = x.neg(); break;
    case 4: r = x.to_mont(); break;
    case 5: r = x.from_mont(); break;
    case 6: r = x.inverse(); break;
    case 7: r = x.dbl(); break;
    case 8: r = x.sqr(); break;                                   // dedicated square (Fp) / plain product (Fr)
    case 9: case 10: {                                            // interleaved pair: (x y, (x + y)(x - y))
      F r1, r2;
      F::mul2(x, y, x + y, x - y, r1, r2);
      r = op == 9 ? r1 : r2;
      break;
    }
    default: r = F::zero();
  }
  for (int j = 0; j < F::N; j++) out[(size_t)i * F::N + j] = r.v[j];
}

// ---- device mixed-add test hook ------------------------------------------------------
// out[i] = 2 p[i] + q[i] through the bucket-accumulation formulas: the accumulator 2 p has
// ZZ != 1, q is the affine addend.  variant 0 = XYZZ::madd, 1 = XYZZ::madd_paired (the one
// k_accumulate<Fp> runs), 2 = general XYZZ::add of from_affine(q).
__global__ void k_test_g1_madd(const uint8_t* __restrict__ p, const uint8_t* __restrict__ q, int variant, uint32_t n,
                               uint8_t* __restrict__ out, int* __restrict__ err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine a, b;
  if (G1Traits::parse(p + (size_t)i * 96, a) || G1Traits::parse(q + (size_t)i * 96, b)) { atomicExch(err, 1); return; }
  G1XYZZ acc = G1XYZZ::dbl_affine(a);
  if (variant == 0) acc.madd(b);
  else if (variant == 1) acc.madd_paired(b);
  else acc.add(G1XYZZ::from_affine(b));
  G1Affine r = acc.to_affine();
  G1Traits::serialize(r, out + (size_t)i * (96 + 48));
}

}  // namespace zk

extern "C" {

int zk_test_g1_madd(const uint8_t* p, const uint8_t* q, int variant, size_t n, uint8_t* out) {
  ZK_API_BEGIN
  using namespace zk;
  ZK_REQUIRE(p && q && out && n > 0 && n < (1u << 24) && variant >= 0 && variant <= 2, ZK_EARG, "test_g1_madd: bad arguments");
  cudaStream_t st = default_stream();
  DevBuf<uint8_t> dp(n * 96), dq(n * 96), dout(n * 144);
  DevBuf<int> derr(1);
  ZK_CUDA(cudaMemcpyAsync(dp.p, p, n * 96, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(dq.p, q, n * 96, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemsetAsync(derr.p, 0, sizeof(int), st));
  k_test_g1_madd<<<cdiv(n, 64), 64, 0, st>>>(dp.p, dq.p, variant, (uint32_t)n, dout.p, derr.p);
  ZK_CUDA(cudaGetLastError());
  int err = 0;
  ZK_CUDA(cudaMemcpyAsync(out, dout.p, n * 144, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaMemcpyAsync(&err, derr.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(err == 0, ZK_EPOINT, "test_g1_madd: point not canonical or not on the curve");
  ZK_API_END
}

int zk_bench_intpipe(int kind, int iters, double* ops_per_s, double* elapsed_ms) {
  ZK_API_BEGIN
  using namespace zk;
  ZK_REQUIRE(ops_per_s && iters > 0, ZK_EARG, "bench_intpipe: bad arguments");
  cudaStream_t st = default_stream();
  DevBuf<uint32_t> d_out(4);
  cudaEvent_t e0, e1;
  ZK_CUDA(cudaEventCreate(&e0));
  ZK_CUDA(cudaEventCreate(&e1));
  int blocks = sm_count() * 8;
  double ops = 0;
  auto launch = [&](int it) {
    switch (kind) {
      case 0: k_mb_mad_lo<<<blocks, 256, 0, st>>>(d_out.p, it, 0x9e3779b9u, 0x7f4a7c15u); ops = (double)blocks * 256 * it * 8 * MB_CHAINS; break;
      case 1: k_mb_mad_cc<<<blocks, 256, 0, st>>>(d_out.p, it, 0x9e3779b9u, 0x7f4a7c15u); ops = (double)blocks * 256 * it * 4 * 16; break;
      case 2: k_mb_mad_wide<<<blocks, 256, 0, st>>>(d_out.p, it, 0x9e3779b9u, 0x7f4a7c15u); ops = (double)blocks * 256 * it * 8 * MB_CHAINS; break;
      case 3: k_mb_field_mul<Fp><<<blocks, 128, 0, st>>>(d_out.p, it); ops = (double)blocks * 128 * it * 2; break;
      case 4: k_mb_field_mul<Fr><<<blocks, 128, 0, st>>>(d_out.p, it); ops = (double)blocks * 128 * it * 2; break;
      case 5: k_mb_g1_madd<<<blocks, 128, 0, st>>>(d_out.p, it); ops = (double)blocks * 128 * it; break;
      case 6: k_mb_madd_var<FpCall, 4><<<blocks, 128, 0, st>>>(d_out.p, it); ops = (double)blocks * 128 * it; break;
      case 7: k_mb_madd_var<Fp, 4><<<blocks, 128, 0, st>>>(d_out.p, it); ops = (double)blocks * 128 * it; break;
      case 8: k_mb_madd_var<FpCall, 5><<<blocks, 128, 0, st>>>(d_out.p, it); ops = (double)blocks * 128 * it; break;
      case 9: k_mb_madd_var<FpCall, 3><<<blocks, 128, 0, st>>>(d_out.p, it); ops = (double)blocks * 128 * it; break;
      case 10: k_mb_madd_paired<4><<<blocks, 128, 0, st>>>(d_out.p, it); ops = (double)blocks * 128 * it; break;
      case 11: k_mb_madd_paired<3><<<blocks, 128, 0, st>>>(d_out.p, it); ops = (double)blocks * 128 * it; break;
      case 12: k_mb_madd_paired<2><<<blocks, 128, 0, st>>>(d_out.p, it); ops = (double)blocks * 128 * it; break;
      default: throw Error{ZK_EARG, "bench_intpipe: unknown kind"};
    }
  };
  launch(iters > 8 ? iters / 8 : 1);  // warm-up
  ZK_CUDA(cudaEventRecord(e0, st));
  launch(iters);
  ZK_CUDA(cudaEventRecord(e1, st));
  ZK_CUDA(cudaGetLastError());
  ZK_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  ZK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ops_per_s = ops / (ms * 1e-3);
  if (elapsed_ms) *elapsed_ms = ms;
  ZK_API_END
}

int zk_test_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  ZK_API_BEGIN
  using namespace zk;
  ZK_REQUIRE(a && b && out && n > 0 && (field == 0 || field == 1), ZK_EARG, "test_field_op: bad arguments");
  cudaStream_t st = default_stream();
  size_t limbs = (field == 0 ? 12 : 8) * n;
  DevBuf<uint32_t> da(limbs), db(limbs), dout(limbs);
  ZK_CUDA(cudaMemcpyAsync(da.p, a, limbs * 4, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(db.p, b, limbs * 4, cudaMemcpyHostToDevice, st));
  if (field == 0) k_test_field_op<Fp><<<cdiv(n, 64), 64, 0, st>>>(op, da.p, db.p, dout.p, (uint32_t)n);
  else k_test_field_op<Fr><<<cdiv(n, 64), 64, 0, st>>>(op, da.p, db.p, dout.p, (uint32_t)n);
  ZK_CUDA(cudaGetLastError());
  ZK_CUDA(cudaMemcpyAsync(out, dout.p, limbs * 4, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_API_END
}

int zk_table_profile(uint64_t handle, int enable, float stage_ms[4]) {
  ZK_API_BEGIN
  using namespace zk;
  HandleBase* hb = lookup_handle(handle, 0);
  ZK_REQUIRE(hb->kind == 1 || hb->kind == 2, ZK_EARG, "table_profile: not a table handle");
  auto go = [&](auto* h) {          // stage times of the primary device's part
    auto& t = h->parts[0]->table;
    if (stage_ms) t.stage_ms(stage_ms);
    for (auto& p : h->parts) p->table.set_profile(enable != 0);
  };
  if (hb->kind == 1) go(static_cast<TableHandle<G1Traits>*>(hb));
  else go(static_cast<TableHandle<G2Traits>*>(hb));
  ZK_API_END
}

// Sums of the stage times over the joins recorded since profiling was switched on (at most the last
// 64), primary device's part: totals[0..3] as in zk_table_profile, counts[0] = MSMs in those joins,
// counts[1] = joins.  The launching stream must have been synchronised.
int zk_table_profile_totals(uint64_t handle, float totals[4], uint64_t counts[2]) {
  ZK_API_BEGIN
  using namespace zk;
  HandleBase* hb = lookup_handle(handle, 0);
  ZK_REQUIRE((hb->kind == 1 || hb->kind == 2) && totals && counts, ZK_EARG, "table_profile_totals: bad arguments");
  if (hb->kind == 1) static_cast<TableHandle<G1Traits>*>(hb)->parts[0]->table.stage_totals(totals, &counts[0], &counts[1]);
  else static_cast<TableHandle<G2Traits>*>(hb)->parts[0]->table.stage_totals(totals, &counts[0], &counts[1]);
  ZK_API_END
}

int zk_table_batch_timing(uint64_t handle, int enable, float* out, size_t cap, size_t* steps) {
  ZK_API_BEGIN
  using namespace zk;
  HandleBase* hb = lookup_handle(handle, 0);
  ZK_REQUIRE(hb->kind == 1 || hb->kind == 2, ZK_EARG, "table_batch_timing: not a table handle");
  ZK_REQUIRE((out == nullptr) == (steps == nullptr), ZK_EARG, "table_batch_timing: out and steps go together");
  if (hb->kind == 1) api_table_batch_timing<G1Traits>(static_cast<TableHandle<G1Traits>*>(hb), enable, out, cap, steps);
  else api_table_batch_timing<G2Traits>(static_cast<TableHandle<G2Traits>*>(hb), enable, out, cap, steps);
  ZK_API_END
}

int zk_table_pipeline(uint64_t handle, int enable) {
  ZK_API_BEGIN
  using namespace zk;
  HandleBase* hb = lookup_handle(handle, 0);
  ZK_REQUIRE(hb->kind == 1 || hb->kind == 2, ZK_EARG, "table_pipeline: not a table handle");
  auto go = [&](auto* h) {
    auto& t = h->single();
    if (!enable && t.queued) { ZK_CUDA(cudaDeviceSynchronize()); t.join(default_stream()); ZK_CUDA(cudaStreamSynchronize(default_stream())); }
    t.set_pipelined(enable != 0, enable > 1 ? enable : 0);   // enable > 1 = queue depth
  };
  if (hb->kind == 1) go(static_cast<TableHandle<G1Traits>*>(hb));
  else go(static_cast<TableHandle<G2Traits>*>(hb));
  ZK_API_END
}

int zk_table_join(uint64_t handle, void* stream) {
  ZK_API_BEGIN
  using namespace zk;
  HandleBase* hb = lookup_handle(handle, 0);
  ZK_REQUIRE(hb->kind == 1 || hb->kind == 2, ZK_EARG, "table_join: not a table handle");
  cudaStream_t st = stream ? (cudaStream_t)stream : default_stream();
  if (hb->kind == 1) static_cast<TableHandle<G1Traits>*>(hb)->single().join(st);
  else static_cast<TableHandle<G2Traits>*>(hb)->single().join(st);
  ZK_API_END
}

int zk_table_info(uint64_t handle, uint64_t info[8]) {
  ZK_API_BEGIN
  using namespace zk;
  HandleBase* hb = lookup_handle(handle, 0);
  ZK_REQUIRE(info && (hb->kind == 1 || hb->kind == 2), ZK_EARG, "table_info: not a table handle");
  MsmConfig cfg;
  size_t bytes = 0, n;
  bool pre;
  auto go = [&](auto* h) {          // window shape of the primary part; bytes and points of the whole table
    cfg = h->parts[0]->table.cfg; n = h->n; pre = h->parts[0]->table.precomputed;
    for (auto& p : h->parts) bytes += p->table.device_bytes();
  };
  if (hb->kind == 1) go(static_cast<TableHandle<G1Traits>*>(hb));
  else go(static_cast<TableHandle<G2Traits>*>(hb));
  info[0] = cfg.c; info[1] = cfg.W; info[2] = cfg.nwb; info[3] = cfg.B; info[4] = cfg.S;
  info[5] = bytes; info[6] = n; info[7] = pre;
  ZK_API_END
}

}  // extern "C"
