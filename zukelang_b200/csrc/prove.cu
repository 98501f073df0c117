// Groth16 and Pinocchio provers on top of the MSM and QAP kernels.
//
// Replaces Groth16.Make(C).prove (/root/reference/src/groth16/groth16.ml:123-161, 235-237),
// Pinocchio.Make(C).Compute.f (src/pinocchio/pinocchio.ml:210-248) and ZKCompute.f (:427-514).
//
// Every proof element is rewritten, by linearity, as ONE multi-scalar multiplication over a
// resident base table made of the relevant proving-key fields; the single scalar
// multiplications of the reference (d1 * r, a * s, b1 * r, vt * dv, ...) become extra
// (base, scalar) pairs of those MSMs instead of 255-step double-and-add tails:
//
//   Groth16   A = [a, d1 | ti1] . [1, r | V]
//             B = [b2, d2 | ti2] . [1, s | W]
//             C = [a, b1, d1 | ti1 | tiztd | ltd_mid] . [s, r, r s | s V + r W | h | w_mid]
//   (C = L + H + s A + r B1 - r s d1 with A, B1 expanded; groth16.ml:154-159.)
//
// The group elements — and so the serialised proofs — are identical to the reference's.
//
// Sharding: a key handle loaded with (shard_index, shard_count) keeps the
// slice [len * i / cnt, len * (i+1) / cnt) of every list-valued field; the single points
// live on shard 0.  prove then returns the shard's partial sums; the caller adds the
// shards' partials (one tiny gather) to obtain the proof.
#include <string.h>
#include <algorithm>
#include <functional>
#include "fr_poly.cuh"
#include "msm.cuh"
#include "runtime.cuh"

namespace zk {

template <class T>
struct Query {
  BaseTable<T> table;
  DevBuf<uint32_t> scalars;  // table.n canonical scalars
  void load(const std::vector<uint8_t>& raw, uint32_t n, cudaStream_t st) {
    bool pre = env_int("ZKB200_KEY_PRECOMPUTE", 1) != 0;
    table.load(raw.data(), nullptr, n, pre, 0, st);
    scalars.alloc((size_t)n * 8);
  }
};

static void slice(size_t len, int idx, int cnt, uint32_t* lo, uint32_t* n) {
  size_t a = len * (size_t)idx / cnt, b = len * (size_t)(idx + 1) / cnt;
  *lo = (uint32_t)a;
  *n = (uint32_t)(b - a);
}
static void append(std::vector<uint8_t>& dst, const uint8_t* src, size_t bytes) { dst.insert(dst.end(), src, src + bytes); }

// ---- scalar helpers ------------------------------------------------------------------
__device__ __forceinline__ void put_raw(uint32_t* dst, size_t i, const Fr& raw) { store_vec(reinterpret_cast<Fr*>(dst) + i, raw); }
__device__ __forceinline__ Fr get_raw(const uint32_t* src, size_t i) { return load_vec_rw(reinterpret_cast<const Fr*>(src) + i); }
__device__ __forceinline__ Fr raw_one() { Fr o = Fr::zero(); o.v[0] = 1; return o; }

// ======================================================================================
// Groth16
// ======================================================================================
struct G16Layout {
  uint32_t n, m, ti_lo, ti_cnt, h_lo, h_cnt, mid_lo, mid_cnt;
  int singles;  // 1 on the shard that owns a, b1, d1, b2, d2
};

// One device's share of the key: slice [len * i / cnt, len * (i + 1) / cnt) of every list-valued field.
struct Groth16Part {
  int ctx = 0;
  G16Layout lay;
  Query<G1Traits> qC;       // [a, b1, d1 | ti1 | tiztd | ltd_mid]; A uses the prefix 3 + ti_cnt
  Query<G2Traits> qB;       // [b2, d2 | ti2]
  DevBuf<uint32_t> sA;      // scalars of A (prefix of the qC table)
  DevBuf<uint32_t> mid_index;
  DevBuf<XYZZ<Fp>> r1;      // scratch results when this is the only part
  DevBuf<XYZZ<Fp2>> r2;
  cudaEvent_t done = nullptr;
  ~Groth16Part() { if (done) cudaEventDestroy(done); }
};

// Proving-key handle.  One part: the whole key on one device, or shard (i, N) of a key sharded
// across PROCESSES by the caller (prove then returns that shard's partial sums).  Several parts:
// after zk_init_devices the key's base ranges are spread over the devices of this process and
// prove returns the finished proof — the parts read their scalars from the primary device and
// store their partial sums into g1 / g2 there (peer-to-peer over NVLink).
struct Groth16Key : HandleBase {
  uint32_t n = 0, m = 0, n_mid = 0;
  std::vector<std::unique_ptr<Groth16Part>> parts;
  DevBuf<uint32_t> d_sol, d_rs;
  DevBuf<XYZZ<Fp>> g1;      // [A, C][part]
  DevBuf<XYZZ<Fp2>> g2;     // [part]
  DevBuf<uint8_t> d_out;
  cudaEvent_t ready = nullptr, ready_h = nullptr;   // primary stream: V | W ready / h ready (the other devices wait on them)
  cudaEvent_t t_begin = nullptr, t_end = nullptr;   // device time of the last prove (zk_groth16_last_device_ms)
  bool early_b_last = true;                         // order of the last prove: B before the quotient (large circuits) or after A and C
  // stage marks of the last prove on the primary device's stream (zk_groth16_last_stage_ms):
  // 0 witness uploaded, 1 V | W | Y ready, 2 B sorted and accumulated, 3 h ready, 4 A and C sorted,
  // 5 A and C accumulated, 6 tails joined (primary device's part throughout)
  static constexpr int NMARK = 7;
  cudaEvent_t mark[NMARK] = {};
  Groth16Key() { kind = 4; }
  ~Groth16Key() {
    if (ready) cudaEventDestroy(ready);
    if (ready_h) cudaEventDestroy(ready_h);
    if (t_begin) cudaEventDestroy(t_begin);
    if (t_end) cudaEventDestroy(t_end);
    for (cudaEvent_t e : mark) if (e) cudaEventDestroy(e);
  }
};

// which & 1: the scalars of B (needs V | W only); which & 2: those of A and C (needs h as well)
static __global__ void __launch_bounds__(128)
k_groth16_scalars(G16Layout L, int which, const Fr* __restrict__ vwy, const Fr* __restrict__ H,
                  const uint32_t* __restrict__ sol_raw, const uint32_t* __restrict__ mid_index,
                  const uint32_t* __restrict__ rs_raw, uint32_t* __restrict__ sA, uint32_t* __restrict__ sB,
                  uint32_t* __restrict__ sC) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const Fr r_raw = get_raw(rs_raw, 0), s_raw = get_raw(rs_raw, 1);
  const bool own = L.singles != 0;
  const Fr zero = Fr::zero();
  if (which & 1) {
    if (t == 0) { put_raw(sB, 0, own ? raw_one() : zero); put_raw(sB, 1, own ? s_raw : zero); }
    if (t < L.ti_cnt) put_raw(sB, 2 + t, load_vec_rw(&vwy[(size_t)L.n + L.ti_lo + t]).from_mont());
  }
  if (!(which & 2)) return;
  if (t == 0) {
    Fr rs = (r_raw.to_mont() * s_raw.to_mont()).from_mont();
    put_raw(sA, 0, own ? raw_one() : zero); put_raw(sA, 1, zero); put_raw(sA, 2, own ? r_raw : zero);
    put_raw(sC, 0, own ? s_raw : zero); put_raw(sC, 1, own ? r_raw : zero); put_raw(sC, 2, own ? rs : zero);
  }
  if (t < L.ti_cnt) {
    uint32_t i = L.ti_lo + t;
    Fr v = load_vec_rw(&vwy[i]), w = load_vec_rw(&vwy[(size_t)L.n + i]);
    put_raw(sA, 3 + t, v.from_mont());
    Fr c = s_raw.to_mont() * v + r_raw.to_mont() * w;
    put_raw(sC, 3 + t, c.from_mont());
  }
  if (t < L.h_cnt) put_raw(sC, 3 + (size_t)L.ti_cnt + t, load_vec_rw(&H[L.h_lo + t]).from_mont());
  if (t < L.mid_cnt) {
    ZK_DCHECK(mid_index[L.mid_lo + t] < L.m);
    put_raw(sC, 3 + (size_t)L.ti_cnt + L.h_cnt + t, get_raw(sol_raw, mid_index[L.mid_lo + t]));
  }
}

// Sum of the shards' partial proof elements (one process per GPU: every rank returns the partial
// a | b | c of its base ranges): block 0 adds the k partial a, block 1 the b, block 2 the c, each
// from the uncompressed bytes at the head of its proof-out slot, and writes the finished
// uncompressed | compressed bytes.  A partial that does not parse raises err.
template <class T>
__device__ void sum_wire_points(const uint8_t* first, size_t stride, uint32_t k, uint8_t* out, int* err) {
  XYZZ<typename T::F> acc = XYZZ<typename T::F>::inf();
  for (uint32_t r = 0; r < k; r++) {
    Affine<typename T::F> p;
    if (T::parse(first + (size_t)r * stride, p)) { atomicExch(err, 1); continue; }
    acc.add(XYZZ<typename T::F>::from_affine(p));      // the compact out-of-line addition: one thread, a short chain
  }
  Affine<typename T::F> a = acc.to_affine();
  T::serialize(a, out);
}
static __global__ void k_groth16_combine(const uint8_t* __restrict__ parts, uint32_t k, uint8_t* __restrict__ out, int* err) {
  if (threadIdx.x) return;
  if (blockIdx.x == 0) sum_wire_points<G1Traits>(parts, ZK_GROTH16_PROOF_OUT, k, out, err);
  else if (blockIdx.x == 1) sum_wire_points<G2Traits>(parts + ZK_G1_OUT, ZK_GROTH16_PROOF_OUT, k, out + ZK_G1_OUT, err);
  else sum_wire_points<G1Traits>(parts + ZK_G1_OUT + ZK_G2_OUT, ZK_GROTH16_PROOF_OUT, k, out + ZK_G1_OUT + ZK_G2_OUT, err);
}

}  // namespace zk

extern "C" {

// parts: k proof-out buffers (ZK_GROTH16_PROOF_OUT bytes each) returned by zk_groth16_prove* on keys
// loaded with shard_count = k; proof_out: the finished proof.  One upload, one launch, one download.
int zk_groth16_combine(const uint8_t* parts, size_t k, uint8_t* proof_out) {
  ZK_API_BEGIN
  using namespace zk;
  ZK_REQUIRE(parts && proof_out && k > 0 && k <= 4096, ZK_EARG, "groth16_combine: bad arguments");
  cudaStream_t st = default_stream();
  static thread_local DevBuf<uint8_t> d_in, d_out;     // calls are serialised (ApiGuard) and stream-ordered
  static thread_local DevBuf<int> d_err;
  d_in.ensure(k * ZK_GROTH16_PROOF_OUT);
  d_out.ensure(ZK_GROTH16_PROOF_OUT);
  d_err.ensure(1);
  ZK_CUDA(cudaMemcpyAsync(d_in.p, parts, k * ZK_GROTH16_PROOF_OUT, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
  k_groth16_combine<<<3, 32, 0, st>>>(d_in.p, (uint32_t)k, d_out.p, d_err.p);
  ZK_CUDA(cudaGetLastError());
  int err = 0;
  ZK_CUDA(cudaMemcpyAsync(proof_out, d_out.p, ZK_GROTH16_PROOF_OUT, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(err == 0, ZK_EPOINT, "groth16_combine: partial point not canonical or not on the curve");
  ZK_API_END
}

int zk_groth16_pk_load(const zk_groth16_pkey* pk, int shard_index, int shard_count, uint64_t* handle) {
  ZK_API_BEGIN
  using namespace zk;
  ZK_REQUIRE(pk && handle && pk->n >= 2 && pk->m >= 1 && shard_count >= 1 && shard_index >= 0 && shard_index < shard_count,
             ZK_EARG, "groth16_pk_load: bad arguments");
  ZK_REQUIRE(pk->a && pk->b1 && pk->d1 && pk->b2 && pk->d2 && pk->ti1 && pk->ti2 && pk->tiztd &&
                 (pk->n_mid == 0 || (pk->ltd_mid && pk->mid_index)),
             ZK_EARG, "groth16_pk_load: null key field");
  for (size_t j = 0; j < pk->n_mid; j++) ZK_REQUIRE(pk->mid_index[j] < pk->m, ZK_EARG, "groth16_pk_load: mid_index out of range");
  auto k = std::make_unique<Groth16Key>();
  k->n = (uint32_t)pk->n;
  k->m = (uint32_t)pk->m;
  k->n_mid = (uint32_t)pk->n_mid;
  // a key given whole to a process that drives several devices is spread over them; a key the
  // caller shards itself (shard_count > 1: one process per GPU) stays on this process' device
  int nparts = shard_count == 1 ? device_count() : 1;
  while (nparts > 1 && pk->n / nparts < 4096) nparts--;
  for (int p = 0; p < nparts; p++) {
    const int idx = shard_count == 1 ? p : shard_index, cnt = shard_count == 1 ? nparts : shard_count;
    auto part = std::make_unique<Groth16Part>();
    part->ctx = p;
    CtxScope scope(p);
    cudaStream_t st = stream_of(p);
    G16Layout& L = part->lay;
    L.n = (uint32_t)pk->n;
    L.m = (uint32_t)pk->m;
    slice(pk->n, idx, cnt, &L.ti_lo, &L.ti_cnt);
    slice(pk->n_h ? pk->n_h : pk->n - 1, idx, cnt, &L.h_lo, &L.h_cnt);
    slice(pk->n_mid, idx, cnt, &L.mid_lo, &L.mid_cnt);
    L.singles = idx == 0;
    std::vector<uint8_t> t1, t2;
    append(t1, pk->a, 96); append(t1, pk->b1, 96); append(t1, pk->d1, 96);
    append(t1, pk->ti1 + (size_t)L.ti_lo * 96, (size_t)L.ti_cnt * 96);
    append(t1, pk->tiztd + (size_t)L.h_lo * 96, (size_t)L.h_cnt * 96);
    if (L.mid_cnt) append(t1, pk->ltd_mid + (size_t)L.mid_lo * 96, (size_t)L.mid_cnt * 96);
    append(t2, pk->b2, 192); append(t2, pk->d2, 192);
    append(t2, pk->ti2 + (size_t)L.ti_lo * 192, (size_t)L.ti_cnt * 192);
    part->qC.load(t1, 3 + L.ti_cnt + L.h_cnt + L.mid_cnt, st);
    part->qB.load(t2, 2 + L.ti_cnt, st);
    part->sA.alloc((size_t)(3 + L.ti_cnt) * 8);
    part->mid_index.alloc(pk->n_mid ? pk->n_mid : 1);
    if (pk->n_mid) ZK_CUDA(cudaMemcpyAsync(part->mid_index.p, pk->mid_index, pk->n_mid * 4, cudaMemcpyHostToDevice, st));
    part->r1.alloc(2);
    part->r2.alloc(1);
    ZK_CUDA(cudaEventCreateWithFlags(&part->done, cudaEventDisableTiming));
    ZK_CUDA(cudaStreamSynchronize(st));
    k->parts.push_back(std::move(part));
  }
  CtxScope primary(0);
  k->d_sol.alloc((size_t)pk->m * 8);
  k->d_rs.alloc(16);
  k->g1.alloc(2 * (size_t)nparts);
  k->g2.alloc(nparts);
  k->d_out.alloc(ZK_GROTH16_PROOF_OUT);
  ZK_CUDA(cudaEventCreateWithFlags(&k->ready, cudaEventDisableTiming));
  ZK_CUDA(cudaEventCreateWithFlags(&k->ready_h, cudaEventDisableTiming));
  ZK_CUDA(cudaEventCreate(&k->t_begin));
  ZK_CUDA(cudaEventCreate(&k->t_end));
  for (cudaEvent_t& e : k->mark) ZK_CUDA(cudaEventCreate(&e));
  *handle = register_handle(std::move(k));
  ZK_API_END
}

// vwy (V | W | Y: coefficients or domain values) is ready on the primary device in stream order on
// st0; `quotient(st0)` enqueues the computation of Hq.  Order of work on every device:
//   B's scalars -> B sorted and accumulated (G2) -> B's tail on the auxiliary stream, while
//   the quotient runs (primary device) -> scalars of A and C -> A and C sorted and accumulated as one
//   list (G1) -> their tail; so the G2 tail hides behind the quotient's transforms instead of
//   standing at the end of the proof next to the G1 tail.
static void groth16_finish(zk::Groth16Key* k, const Fr* vwy, const Fr* Hq, int* flag, cudaStream_t st0, uint8_t* proof_out,
                           const std::function<void(cudaStream_t)>& quotient) {
  using namespace zk;
  const int np = (int)k->parts.size();
  ZK_CUDA(cudaEventRecord(k->mark[1], st0));
  if (np > 1) ZK_CUDA(cudaEventRecord(k->ready, st0));
  // the MSMs are queued on their tables; an error before their tails have been enqueued drains the
  // devices and drops the queues (PipelineScope)
  std::vector<std::unique_ptr<PipelineScope<G1Traits>>> scopes1;
  std::vector<std::unique_ptr<PipelineScope<G2Traits>>> scopes2;
  for (int p = 0; p < np; p++) {
    Groth16Part& P = *k->parts[p];
    CtxScope scope(P.ctx);                            // the first use sizes the queue: on the part's own device
    scopes1.push_back(std::make_unique<PipelineScope<G1Traits>>(P.qC.table, P.ctx, nullptr, 2));
    scopes2.push_back(std::make_unique<PipelineScope<G2Traits>>(P.qB.table, P.ctx, nullptr, 1));
  }
  auto span_of = [](const G16Layout& L) { return std::max(std::max(L.ti_cnt, L.h_cnt), std::max(L.mid_cnt, 1u)); };
  // Where B goes.  From 2^15 constraints on: B first, so that its G2 tail hides behind the quotient's
  // transforms (3.3 ms at 2^20 constraints) and the G1 work.  The price is that tail blocks still
  // resident keep the persistent G1 accumulation from taking its two blocks per SM for a while; at
  // 2^16 constraints that costs a circuit with well-spread B scalars 0.3 ms and gains the multiply
  // chain (whose G2 tail carries a heavy-bucket fix-up) 0.9 ms.  Tiny circuits keep B after A and C,
  // with the two tails side by side at the end.
  const bool early_b = k->n >= (1u << 15);
  k->early_b_last = early_b;
  auto run_b = [&](Groth16Part& P, int p, cudaStream_t st, bool tail_now) {
    const G16Layout& L = P.lay;
    k_groth16_scalars<<<cdiv(std::max(L.ti_cnt, 1u), 128), 128, 0, st>>>(L, 1, vwy, Hq, k->d_sol.p, P.mid_index.p, k->d_rs.p,
                                                                          P.sA.p, P.qB.scalars.p, P.qC.scalars.p);
    XYZZ<Fp2>* rB = np > 1 ? k->g2.p + p : P.r2.p;
    uint8_t* o = np > 1 ? nullptr : k->d_out.p;       // one part: wire bytes straight from the tail
    P.qB.table.run(P.qB.scalars.p, P.qB.table.n, rB, o ? o + ZK_G1_OUT : nullptr, st);                     // B
    P.qB.table.sort_accumulate(st);
    if (P.ctx == 0) ZK_CUDA(cudaEventRecord(k->mark[2], st));
    if (tail_now) {
      cudaStream_t aux = fork_aux(st);
      P.qB.table.tail(aux);                           // joined after the G1 tail has been enqueued
    }
  };
  // ---- B on every device (large circuits) ---------------------------------------------------------
  if (early_b)
    for (int p = 0; p < np; p++) {
      Groth16Part& P = *k->parts[p];
      CtxScope scope(P.ctx);
      cudaStream_t st = stream_of(P.ctx);
      if (P.ctx != 0) ZK_CUDA(cudaStreamWaitEvent(st, k->ready, 0));
      run_b(P, p, st, true);
    }
  // ---- h on the primary device -------------------------------------------------------------------
  quotient(st0);
  ZK_CUDA(cudaEventRecord(k->mark[3], st0));
  if (np > 1) ZK_CUDA(cudaEventRecord(k->ready_h, st0));
  // ---- A and C on every device (then B, for small circuits) -----------------------------------------
  for (int p = 0; p < np; p++) {
    Groth16Part& P = *k->parts[p];
    const G16Layout& L = P.lay;
    CtxScope scope(P.ctx);
    cudaStream_t st = stream_of(P.ctx);
    const bool primary = P.ctx == 0;
    if (!primary) ZK_CUDA(cudaStreamWaitEvent(st, k->ready_h, 0));
    k_groth16_scalars<<<cdiv(span_of(L), 128), 128, 0, st>>>(L, 2, vwy, Hq, k->d_sol.p, P.mid_index.p, k->d_rs.p, P.sA.p,
                                                              P.qB.scalars.p, P.qC.scalars.p);
    XYZZ<Fp>* rA = np > 1 ? k->g1.p + p : P.r1.p;
    XYZZ<Fp>* rC = np > 1 ? k->g1.p + np + p : P.r1.p + 1;
    uint8_t* o = np > 1 ? nullptr : k->d_out.p;
    P.qC.table.run(P.sA.p, 3 + L.ti_cnt, rA, o, st);                                                       // A  } queued: one sort and
    P.qC.table.run(P.qC.scalars.p, P.qC.table.n, rC, o ? o + ZK_G1_OUT + ZK_G2_OUT : nullptr, st);         // C  } one accumulation
    P.qC.table.sort_accumulate(st, primary ? k->mark[4] : nullptr);
    if (primary) ZK_CUDA(cudaEventRecord(k->mark[5], st));
    if (early_b) {
      P.qC.table.tail(st);
    } else {
      run_b(P, p, st, false);
      cudaStream_t aux = fork_aux(st);
      P.qB.table.tail(aux, 2);
      P.qC.table.tail(st, 2);
    }
    join_aux(st);
    if (primary) ZK_CUDA(cudaEventRecord(k->mark[6], st));
    if (np > 1) ZK_CUDA(cudaEventRecord(P.done, st));
  }
  if (np > 1) {
    for (int p = 0; p < np; p++) ZK_CUDA(cudaStreamWaitEvent(st0, k->parts[p]->done, 0));
    k_sum_parts<G1Traits><<<2, 32, 0, st0>>>(k->g1.p, (uint32_t)np, k->d_out.p, ZK_G1_OUT + ZK_G2_OUT);   // A, C
    k_sum_parts<G2Traits><<<1, 32, 0, st0>>>(k->g2.p, (uint32_t)np, k->d_out.p + ZK_G1_OUT, ZK_G2_OUT);   // B
    ZK_CUDA(cudaGetLastError());
  }
  int fl[2];
  ZK_CUDA(cudaMemcpyAsync(proof_out, k->d_out.p, ZK_GROTH16_PROOF_OUT, cudaMemcpyDeviceToHost, st0));
  ZK_CUDA(cudaMemcpyAsync(fl, flag, sizeof(fl), cudaMemcpyDeviceToHost, st0));
  ZK_CUDA(cudaEventRecord(k->t_end, st0));
  ZK_CUDA(cudaStreamSynchronize(st0));
  ZK_REQUIRE(fl[0] == 0, ZK_EPOINT, "groth16_prove: scalar is not canonical (>= r)");
  ZK_REQUIRE(fl[1] == 0, ZK_EREMAINDER, "groth16_prove: V*W - Y is not divisible by the target (QAP.ml:134)");
}

static void check_rs(const uint8_t* r, const uint8_t* s) {
  // canonical check of the two blinding scalars on the host side of the ABI: byte compare with r
  static const uint8_t R_LE[32] = {0x01, 0x00, 0x00, 0x00, 0xff, 0xff, 0xff, 0xff, 0xfe, 0x5b, 0xfe, 0xff, 0x02, 0xa4, 0xbd, 0x53,
                                   0x05, 0xd8, 0xa1, 0x09, 0x08, 0xd8, 0x39, 0x33, 0x48, 0x7d, 0x9d, 0x29, 0x53, 0xa7, 0xed, 0x73};
  for (const uint8_t* x : {r, s}) {
    bool less = false;
    for (int i = 31; i >= 0; i--) {
      if (x[i] != R_LE[i]) { less = x[i] < R_LE[i]; break; }
    }
    ZK_REQUIRE(less, ZK_EPOINT, "blinding scalar is not canonical (>= r)");
  }
}

int zk_groth16_prove(uint64_t pk_handle, uint64_t qap_handle, const uint8_t* sol, const uint8_t* r, const uint8_t* s,
                     uint8_t* proof_out) {
  ZK_API_BEGIN
  using namespace zk;
  auto* k = static_cast<Groth16Key*>(lookup_handle(pk_handle, 4));
  auto* qh = static_cast<QapHandle*>(lookup_handle(qap_handle, 3));
  QapDevice& q = qh->q;
  ZK_REQUIRE(sol && r && s && proof_out, ZK_EARG, "groth16_prove: null argument");
  ZK_REQUIRE(q.n == k->n && q.m == k->m, ZK_EARG, "groth16_prove: key and QAP dimensions differ");
  check_rs(r, s);
  cudaStream_t st = default_stream();
  ZK_CUDA(cudaEventRecord(k->t_begin, st));
  ZK_CUDA(cudaMemcpyAsync(k->d_sol.p, sol, (size_t)k->m * 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(k->d_rs.p, r, 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(k->d_rs.p + 8, s, 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaEventRecord(k->mark[0], st));
  q.combine(k->d_sol.p, st);
  groth16_finish(k, q.Vc.p, q.H.p, q.flag.p, st, proof_out, [&](cudaStream_t s0) { q.quotient_from_work(s0); });
  ZK_API_END
}

// Same, with V | W | Y given as coefficient vectors (3 * n scalars) instead of a dense QAP:
// `qap_handle` then only carries the target (zk_qap_load with m = 0 ... see zk_quotient_domain_load).
int zk_groth16_prove_coeffs(uint64_t pk_handle, uint64_t qap_handle, const uint8_t* vwy, const uint8_t* sol,
                            const uint8_t* r, const uint8_t* s, uint8_t* proof_out) {
  ZK_API_BEGIN
  using namespace zk;
  auto* k = static_cast<Groth16Key*>(lookup_handle(pk_handle, 4));
  auto* qh = static_cast<QapHandle*>(lookup_handle(qap_handle, 3));
  QapDevice& q = qh->q;
  ZK_REQUIRE(vwy && sol && r && s && proof_out, ZK_EARG, "groth16_prove_coeffs: null argument");
  ZK_REQUIRE(q.n == k->n, ZK_EARG, "groth16_prove_coeffs: key and domain dimensions differ");
  check_rs(r, s);
  cudaStream_t st = default_stream();
  ZK_CUDA(cudaEventRecord(k->t_begin, st));
  qh->d_raw.ensure(3 * (size_t)q.n * 8);
  ZK_CUDA(cudaMemcpyAsync(qh->d_raw.p, vwy, 3 * (size_t)q.n * 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(k->d_sol.p, sol, (size_t)k->m * 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(k->d_rs.p, r, 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(k->d_rs.p + 8, s, 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaEventRecord(k->mark[0], st));
  q.set_coeffs(qh->d_raw.p, st);
  groth16_finish(k, q.Vc.p, q.H.p, q.flag.p, st, proof_out, [&](cudaStream_t s0) { q.quotient_from_work(s0); });
  ZK_API_END
}

// Evaluation-form prover for large circuits: `pk` holds the Lagrange-basis derived key (ti1 = [L_j(tau)]1,
// ti2 = [L_j(tau)]2, tiztd = [L'_k(tau) Z(tau)/delta]1 with n_h = n), `domain` the evaluation domain
// with the circuit's sparse matrices.  Same proof elements as zk_groth16_prove on the dense QAP.
int zk_groth16_prove_r1cs(uint64_t pk_handle, uint64_t domain_handle, const uint8_t* sol, const uint8_t* r,
                          const uint8_t* s, uint8_t* proof_out) {
  ZK_API_BEGIN
  using namespace zk;
  auto* k = static_cast<Groth16Key*>(lookup_handle(pk_handle, 4));
  auto* dh = static_cast<EvalDomainHandle*>(lookup_handle(domain_handle, 6));
  EvalDomain& d = dh->d;
  ZK_REQUIRE(sol && r && s && proof_out, ZK_EARG, "groth16_prove_r1cs: null argument");
  ZK_REQUIRE(d.n == k->n && d.m == k->m, ZK_EARG, "groth16_prove_r1cs: key and domain dimensions differ");
  check_rs(r, s);
  cudaStream_t st = default_stream();
  ZK_CUDA(cudaEventRecord(k->t_begin, st));
  // `sol` may also be a device pointer (a witness all-gathered over NVLink by a sharded caller): the
  // copy infers its direction; a device buffer must be complete before the call
  ZK_CUDA(cudaMemcpyAsync(k->d_sol.p, sol, (size_t)k->m * 32, cudaMemcpyDefault, st));
  ZK_CUDA(cudaMemcpyAsync(k->d_rs.p, r, 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(k->d_rs.p + 8, s, 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaEventRecord(k->mark[0], st));
  d.values(k->d_sol.p, st);
  groth16_finish(k, d.evals.p, d.H.p, d.flag.p, st, proof_out, [&](cudaStream_t s0) { d.quotient(s0); });
  ZK_API_END
}

// Device time of the last zk_groth16_prove* on this key: from the first upload to the last
// download, CUDA events on the primary device's stream (the other devices' work is inside: the
// primary stream waits for it before the combine).
int zk_groth16_last_device_ms(uint64_t pk_handle, float* ms) {
  ZK_API_BEGIN
  using namespace zk;
  auto* k = static_cast<Groth16Key*>(lookup_handle(pk_handle, 4));
  ZK_REQUIRE(ms, ZK_EARG, "groth16_last_device_ms: null argument");
  ZK_CUDA(cudaEventElapsedTime(ms, k->t_begin, k->t_end));
  ZK_API_END
}

// Stage split of the last zk_groth16_prove* on this key, in ms (CUDA events on the primary device's
// stream): [0] witness upload, [1] QAP evaluation (V | W | Y), [2] B: scalars, sort, accumulation,
// [3] quotient h (B's tail runs beside it on the auxiliary stream), [4] A and C: scalars and counting
// sort (one list), [5] their accumulation (one launch), [6] G1 tail + what is left of the G2 tail,
// [7] wait for the other devices + combine + download.
int zk_groth16_last_stage_ms(uint64_t pk_handle, float out[8]) {
  ZK_API_BEGIN
  using namespace zk;
  auto* k = static_cast<Groth16Key*>(lookup_handle(pk_handle, 4));
  ZK_REQUIRE(out, ZK_EARG, "groth16_last_stage_ms: null argument");
  // consecutive marks when B ran first; for small circuits the events fall in the order
  // begin, 0, 1, 3, 4, 5, 2, 6, end and the same eight quantities are read off pairwise
  cudaEvent_t* m = k->mark;
  cudaEvent_t early[8][2] = {{k->t_begin, m[0]}, {m[0], m[1]}, {m[1], m[2]}, {m[2], m[3]}, {m[3], m[4]}, {m[4], m[5]}, {m[5], m[6]}, {m[6], k->t_end}};
  cudaEvent_t late[8][2] = {{k->t_begin, m[0]}, {m[0], m[1]}, {m[5], m[2]}, {m[1], m[3]}, {m[3], m[4]}, {m[4], m[5]}, {m[2], m[6]}, {m[6], k->t_end}};
  for (int i = 0; i < 8; i++) {
    cudaEvent_t* pr = k->early_b_last ? early[i] : late[i];
    ZK_CUDA(cudaEventElapsedTime(&out[i], pr[0], pr[1]));
  }
  ZK_API_END
}

int zk_key_free(uint64_t handle) {
  ZK_API_BEGIN
  zk::HandleBase* h = zk::lookup_handle(handle, 0);
  ZK_REQUIRE(h->kind == 4 || h->kind == 5, ZK_EARG, "key_free: not a key handle");
  zk::sync_all_devices();
  zk::drop_handle(handle);
  ZK_API_END
}

}  // extern "C"

// ======================================================================================
// Pinocchio (Protocol 2), NonZK and ZK
// ======================================================================================
namespace zk {

struct PinLayout {
  uint32_t n, m, mid_lo, mid_cnt, si_lo, si_cnt, all_lo, all_cnt;
  int singles;
};

struct PinocchioKey : HandleBase {
  PinLayout lay;
  uint32_t m = 0, n_mid = 0;
  // All G1 queries live back to back in ONE table (likewise the two G2 queries), each proof element
  // is an MSM over its range: the six (two) latency-bound tails then run as one batched launch.
  // G1 ranges: 0 vv, 1 yy, 2 vav, 3 yay, 4 bvwy, 5 h ; G2 ranges: 0 ww, 1 waw
  Query<G1Traits> q1;
  Query<G2Traits> q2;
  uint32_t first1[6], cnt1[6], first2[2], cnt2[2];
  DevBuf<uint32_t> mid_index, d_sol, d_d;
  DevBuf<uint8_t> d_out;
  PinocchioKey() { kind = 5; }
};

// Scalar vectors of the eight queries (pinocchio.ml:438-505):
//   [dv | c_mid] for vv, vav ; [dw | c_mid] for ww, waw ; [dy | c_mid] for yy, yay ;
//   [dv, dw, dy | c_mid] for bvwy ;
//   h' : [-dy | h_i + dv dw target_i (si) | dw c (v_all) | dv c (w_all)]
static __global__ void __launch_bounds__(128)
k_pinocchio_scalars(PinLayout L, const Fr* __restrict__ H, const Fr* __restrict__ target,
                    const uint32_t* __restrict__ sol_raw, const uint32_t* __restrict__ mid_index,
                    const uint32_t* __restrict__ d_raw, uint32_t* s_vv, uint32_t* s_ww, uint32_t* s_yy, uint32_t* s_vav,
                    uint32_t* s_waw, uint32_t* s_yay, uint32_t* s_bvwy, uint32_t* s_h) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const Fr dv = get_raw(d_raw, 0), dw = get_raw(d_raw, 1), dy = get_raw(d_raw, 2);
  const Fr zero = Fr::zero();
  const bool own = L.singles != 0;
  if (t == 0) {
    put_raw(s_vv, 0, own ? dv : zero); put_raw(s_vav, 0, own ? dv : zero);
    put_raw(s_ww, 0, own ? dw : zero); put_raw(s_waw, 0, own ? dw : zero);
    put_raw(s_yy, 0, own ? dy : zero); put_raw(s_yay, 0, own ? dy : zero);
    put_raw(s_bvwy, 0, own ? dv : zero); put_raw(s_bvwy, 1, own ? dw : zero); put_raw(s_bvwy, 2, own ? dy : zero);
    put_raw(s_h, 0, own ? dy.to_mont().neg().from_mont() : zero);
  }
  if (t < L.mid_cnt) {
    Fr c = get_raw(sol_raw, mid_index[L.mid_lo + t]);
    put_raw(s_vv, 1 + t, c); put_raw(s_vav, 1 + t, c); put_raw(s_ww, 1 + t, c); put_raw(s_waw, 1 + t, c);
    put_raw(s_yy, 1 + t, c); put_raw(s_yay, 1 + t, c); put_raw(s_bvwy, 3 + t, c);
  }
  if (t < L.si_cnt) {
    uint32_t i = L.si_lo + t;
    Fr hv = (i + 1 < L.n) ? load_vec_rw(&H[i]) : zero;                 // h has n - 1 coefficients
    Fr dvdw = dv.to_mont() * dw.to_mont();
    Fr tv = load_vec_rw(&target[i]);                                    // n + 1 coefficients
    put_raw(s_h, 1 + t, (hv + dvdw * tv).from_mont());
  }
  if (t < L.all_cnt) {
    Fr c = get_raw(sol_raw, L.all_lo + t).to_mont();
    put_raw(s_h, 1 + (size_t)L.si_cnt + t, (dw.to_mont() * c).from_mont());
    put_raw(s_h, 1 + (size_t)L.si_cnt + L.all_cnt + t, (dv.to_mont() * c).from_mont());
  }
}

}  // namespace zk

extern "C" {

int zk_pinocchio_pk_load(const zk_pinocchio_pkey* pk, int shard_index, int shard_count, uint64_t* handle) {
  ZK_API_BEGIN
  using namespace zk;
  ZK_REQUIRE(pk && handle && pk->n >= 2 && pk->m >= 1 && shard_count >= 1 && shard_index >= 0 && shard_index < shard_count,
             ZK_EARG, "pinocchio_pk_load: bad arguments");
  ZK_REQUIRE(pk->si && pk->v_all && pk->w_all && pk->one && pk->vt && pk->wt && pk->yt && pk->vavt && pk->wawt &&
                 pk->yayt && pk->vbt && pk->wbt && pk->ybt &&
                 (pk->n_mid == 0 || (pk->vv && pk->ww && pk->yy && pk->vav && pk->waw && pk->yay && pk->bvwy && pk->mid_index)),
             ZK_EARG, "pinocchio_pk_load: null key field");
  for (size_t j = 0; j < pk->n_mid; j++) ZK_REQUIRE(pk->mid_index[j] < pk->m, ZK_EARG, "pinocchio_pk_load: mid_index out of range");
  cudaStream_t st = default_stream();
  auto k = std::make_unique<PinocchioKey>();
  PinLayout& L = k->lay;
  L.n = (uint32_t)pk->n;
  L.m = (uint32_t)pk->m;
  k->m = L.m;
  k->n_mid = (uint32_t)pk->n_mid;
  slice(pk->n_mid, shard_index, shard_count, &L.mid_lo, &L.mid_cnt);
  slice(pk->n + 1, shard_index, shard_count, &L.si_lo, &L.si_cnt);
  slice(pk->m, shard_index, shard_count, &L.all_lo, &L.all_cnt);
  L.singles = shard_index == 0;
  std::vector<uint8_t> raw1, raw2;
  uint32_t at1 = 0, at2 = 0;
  auto g1q = [&](int slot, std::initializer_list<const uint8_t*> singles, const uint8_t* list) {
    for (const uint8_t* sp : singles) append(raw1, sp, 96);
    if (L.mid_cnt) append(raw1, list + (size_t)L.mid_lo * 96, (size_t)L.mid_cnt * 96);
    k->first1[slot] = at1;
    k->cnt1[slot] = (uint32_t)singles.size() + L.mid_cnt;
    at1 += k->cnt1[slot];
  };
  auto g2q = [&](int slot, const uint8_t* single, const uint8_t* list) {
    append(raw2, single, 192);
    if (L.mid_cnt) append(raw2, list + (size_t)L.mid_lo * 192, (size_t)L.mid_cnt * 192);
    k->first2[slot] = at2;
    k->cnt2[slot] = 1 + L.mid_cnt;
    at2 += k->cnt2[slot];
  };
  g1q(0, {pk->vt}, pk->vv);
  g1q(1, {pk->yt}, pk->yy);
  g1q(2, {pk->vavt}, pk->vav);
  g1q(3, {pk->yayt}, pk->yay);
  g1q(4, {pk->vbt, pk->wbt, pk->ybt}, pk->bvwy);
  append(raw1, pk->one, 96);
  append(raw1, pk->si + (size_t)L.si_lo * 96, (size_t)L.si_cnt * 96);
  append(raw1, pk->v_all + (size_t)L.all_lo * 96, (size_t)L.all_cnt * 96);
  append(raw1, pk->w_all + (size_t)L.all_lo * 96, (size_t)L.all_cnt * 96);
  k->first1[5] = at1;
  k->cnt1[5] = 1 + L.si_cnt + 2 * L.all_cnt;
  at1 += k->cnt1[5];
  g2q(0, pk->wt, pk->ww);
  g2q(1, pk->wawt, pk->waw);
  k->q1.load(raw1, at1, st);
  k->q2.load(raw2, at2, st);
  k->mid_index.alloc(pk->n_mid ? pk->n_mid : 1);
  if (pk->n_mid) ZK_CUDA(cudaMemcpyAsync(k->mid_index.p, pk->mid_index, pk->n_mid * 4, cudaMemcpyHostToDevice, st));
  k->d_sol.alloc((size_t)pk->m * 8);
  k->d_d.alloc(24);
  k->d_out.alloc(ZK_PINOCCHIO_PROOF_OUT);
  ZK_CUDA(cudaStreamSynchronize(st));
  *handle = register_handle(std::move(k));
  ZK_API_END
}

// d: three blinding scalars dv | dw | dy (ZK.prove), or NULL for NonZK.prove.
// proof_out: vv | ww | yy | h | vavv | waww | yayy | bvwy as point results (pinocchio.ml:195-208 order).
int zk_pinocchio_prove(uint64_t pk_handle, uint64_t qap_handle, const uint8_t* sol, const uint8_t* d, uint8_t* proof_out) {
  ZK_API_BEGIN
  using namespace zk;
  auto* k = static_cast<PinocchioKey*>(lookup_handle(pk_handle, 5));
  auto* qh = static_cast<QapHandle*>(lookup_handle(qap_handle, 3));
  QapDevice& q = qh->q;
  ZK_REQUIRE(sol && proof_out, ZK_EARG, "pinocchio_prove: null argument");
  ZK_REQUIRE(q.n == k->lay.n && q.m == k->m, ZK_EARG, "pinocchio_prove: key and QAP dimensions differ");
  uint8_t zeros[96] = {0};
  const uint8_t* dd = d ? d : zeros;
  if (d) { check_rs(d, d + 32); check_rs(d + 64, d + 64); }
  cudaStream_t st = default_stream();
  const PinLayout& L = k->lay;
  ZK_CUDA(cudaMemcpyAsync(k->d_sol.p, sol, (size_t)k->m * 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(k->d_d.p, dd, 96, cudaMemcpyHostToDevice, st));
  q.eval(k->d_sol.p, st);
  uint32_t span = std::max(std::max(L.mid_cnt, L.si_cnt), std::max(L.all_cnt, 1u));
  uint32_t* s1 = k->q1.scalars.p;
  uint32_t* s2 = k->q2.scalars.p;
  auto at1 = [&](int slot) { return s1 + 8 * (size_t)k->first1[slot]; };
  auto at2 = [&](int slot) { return s2 + 8 * (size_t)k->first2[slot]; };
  k_pinocchio_scalars<<<cdiv(span, 128), 128, 0, st>>>(L, q.H.p, q.target.p, k->d_sol.p, k->mid_index.p, k->d_d.p,
                                                        at1(0), at2(0), at1(1), at1(2), at2(1), at1(3), at1(4), at1(5));
  // output order: vv ww yy h vavv waww yayy bvwy (pinocchio.ml:195-208)
  uint8_t* o = k->d_out.p;
  uint8_t* o_vv = o;                       o += ZK_G1_OUT;
  uint8_t* o_ww = o;                       o += ZK_G2_OUT;
  uint8_t* o_yy = o;                       o += ZK_G1_OUT;
  uint8_t* o_h = o;                        o += ZK_G1_OUT;
  uint8_t* o_vav = o;                      o += ZK_G1_OUT;
  uint8_t* o_waw = o;                      o += ZK_G2_OUT;
  uint8_t* o_yay = o;                      o += ZK_G1_OUT;
  uint8_t* o_bvwy = o;
  PipelineScope<G1Traits> scope1(k->q1.table, k->ctx, nullptr, 6);
  PipelineScope<G2Traits> scope2(k->q2.table, k->ctx, nullptr, 2);
  auto run1 = [&](int slot, uint8_t* out) { k->q1.table.run(at1(slot), k->cnt1[slot], nullptr, out, st, k->first1[slot]); };
  auto run2 = [&](int slot, uint8_t* out) { k->q2.table.run(at2(slot), k->cnt2[slot], nullptr, out, st, k->first2[slot]); };
  run1(0, o_vv); run1(1, o_yy); run1(5, o_h); run1(2, o_vav); run1(3, o_yay); run1(4, o_bvwy);
  run2(0, o_ww); run2(1, o_waw);
  // the throughput-bound halves back to back on this stream (one sort and one accumulation per
  // group), then the two latency-bound tails side by side
  k->q1.table.sort_accumulate(st);
  k->q2.table.sort_accumulate(st);
  cudaStream_t aux = fork_aux(st);
  k->q2.table.tail(aux, 2);  // one batched tail for the two G2 elements, next to ...
  k->q1.table.tail(st, 2);   // ... the one for the six G1 elements
  join_aux(st);
  int fl[2];
  ZK_CUDA(cudaMemcpyAsync(proof_out, k->d_out.p, ZK_PINOCCHIO_PROOF_OUT, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaMemcpyAsync(fl, q.flag.p, sizeof(fl), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(fl[0] == 0, ZK_EPOINT, "pinocchio_prove: scalar is not canonical (>= r)");
  ZK_REQUIRE(fl[1] == 0, ZK_EREMAINDER, "pinocchio_prove: V*W - Y is not divisible by the target (QAP.ml:134)");
  ZK_API_END
}

}  // extern "C"
