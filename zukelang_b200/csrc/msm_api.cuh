// extern "C" bodies for one group (included by msm_g1.cu and msm_g2.cu with T fixed).
#pragma once
#include "msm_impl.cuh"
#include "runtime.cuh"

namespace zk {

// Staging of zk_g*_table_msm_batch, owned by the handle: nothing is allocated, created or freed
// per call once the ring has its size (round 1 paid three cudaMalloc, four cudaEventCreate and three
// device-synchronising cudaFree inside the timed call).  One scalar buffer per queued MSM: a group
// of MSMs is uploaded on the copy stream while the previous group is being accumulated.
constexpr int BATCH_TIMED_STEPS = 64;
struct BatchStage {
  std::vector<DevBuf<uint32_t>> d_sc;    // ring of scalar vectors, one per MSM of a group
  cudaStream_t copy = nullptr;           // upload stream
  cudaEvent_t copied = nullptr;          // copy stream: the current group is on the device
  cudaEvent_t scattered = nullptr;       // compute stream: the current group's scalars have been read for the last time
  // per-step timing (zk_table_batch_timing): upload start / end on the copy stream; first / last kernel
  // of the step's GROUP on the compute stream
  bool timed = false;
  int steps_timed = 0;
  int group_of[BATCH_TIMED_STEPS] = {};
  cudaEvent_t t_c0[BATCH_TIMED_STEPS] = {}, t_c1[BATCH_TIMED_STEPS] = {}, t_k0[BATCH_TIMED_STEPS] = {},
              t_k1[BATCH_TIMED_STEPS] = {};
  void ensure(size_t n, int depth) {
    if ((int)d_sc.size() < depth) d_sc.resize(depth);
    for (int i = 0; i < depth; i++) d_sc[i].ensure(n * 8);
    if (!copy) {
      ZK_CUDA(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
      ZK_CUDA(cudaEventCreateWithFlags(&copied, cudaEventDisableTiming));
      ZK_CUDA(cudaEventCreateWithFlags(&scattered, cudaEventDisableTiming));
    }
    if (timed && !t_c0[0])
      for (int i = 0; i < BATCH_TIMED_STEPS; i++) {
        ZK_CUDA(cudaEventCreate(&t_c0[i])); ZK_CUDA(cudaEventCreate(&t_c1[i]));
        ZK_CUDA(cudaEventCreate(&t_k0[i])); ZK_CUDA(cudaEventCreate(&t_k1[i]));
      }
  }
  ~BatchStage() {
    if (copy) cudaStreamDestroy(copy);
    if (copied) cudaEventDestroy(copied);
    if (scattered) cudaEventDestroy(scattered);
    if (t_c0[0])
      for (int i = 0; i < BATCH_TIMED_STEPS; i++) {
        cudaEventDestroy(t_c0[i]); cudaEventDestroy(t_c1[i]); cudaEventDestroy(t_k0[i]); cudaEventDestroy(t_k1[i]);
      }
  }
};

// One device's share of a table: the base range [lo, lo + table.n).
template <class T>
struct TablePart {
  int ctx = 0;
  uint32_t lo = 0;
  BaseTable<T> table;
  DevBuf<uint32_t> d_scalars;            // staging for host-scalar calls
  BatchStage batch;
  cudaEvent_t done = nullptr;            // this part's share of the current call is enqueued up to here
  ~TablePart() { if (done) cudaEventDestroy(done); }
};

// A table handle: one part when the library drives one device or the caller shards across
// processes itself, one part per device after zk_init_devices.  Partial sums of the parts are
// stored into d_gather on the primary device (peer-to-peer stores) and added there.
template <class T>
struct TableHandle : HandleBase {
  typedef typename T::F F;
  uint32_t n = 0;
  std::vector<std::unique_ptr<TablePart<T>>> parts;
  DevBuf<XYZZ<F>> d_result;              // one XYZZ result (single part)
  DevBuf<XYZZ<F>> d_gather;              // [slot][part]
  DevBuf<uint8_t> d_out;                 // RAW + COMP bytes per slot
  DevBuf<int> d_err;
  cudaEvent_t ready = nullptr;           // primary stream: error flag cleared, parts may start
  TableHandle() { kind = T::ID; }
  ~TableHandle() { if (ready) cudaEventDestroy(ready); }
  BaseTable<T>& single() {
    ZK_REQUIRE(parts.size() == 1, ZK_EARG, "this entry point takes device pointers: it needs a table on ONE device");
    return parts[0]->table;
  }
  int active_parts(size_t count) const {   // parts that hold some of the first `count` points
    int k = 0;
    for (auto& p : parts) if (p->lo < count) k++;
    return k;
  }
};

// scalars must be canonical (< r): checked on the device for host-facing calls
__global__ void k_check_scalars(const uint32_t* __restrict__ scalars, uint32_t n, int* __restrict__ err);

template <class T>
void table_alloc_common(TableHandle<T>* h) {
  h->d_result.alloc(1);
  h->d_out.alloc(T::RAW + T::COMP);
  h->d_err.alloc(1);
  ZK_CUDA(cudaEventCreateWithFlags(&h->ready, cudaEventDisableTiming));
}

// Splits [0, n) over `nparts` devices and loads every part on its own device.
template <class T>
void table_load_parts(TableHandle<T>* h, const uint8_t* bases, const uint8_t* inf_flags, uint32_t n, bool precompute,
                      int window_bits, int nparts) {
  h->n = n;
  h->ctx = 0;
  for (int p = 0; p < nparts; p++) {
    uint32_t lo = (uint32_t)((uint64_t)n * p / nparts), hi = (uint32_t)((uint64_t)n * (p + 1) / nparts);
    auto part = std::make_unique<TablePart<T>>();
    part->ctx = p;
    part->lo = lo;
    CtxScope scope(p);
    part->table.load(bases + (size_t)lo * T::RAW, inf_flags ? inf_flags + lo : nullptr, hi - lo, precompute, window_bits,
                     stream_of(p));
    ZK_CUDA(cudaEventCreateWithFlags(&part->done, cudaEventDisableTiming));
    h->parts.push_back(std::move(part));
  }
  CtxScope scope(0);
  table_alloc_common(h);
}

template <class T>
int api_table_load(const uint8_t* bases, const uint8_t* inf_flags, size_t n, int precompute, int window_bits,
                   uint64_t* handle) {
  ZK_API_BEGIN
  ZK_REQUIRE(bases && handle && n > 0 && n < (1ull << 28), ZK_EARG, "table_load: bad arguments");
  auto h = std::make_unique<TableHandle<T>>();
  // small tables stay on the primary device: below ~2^12 points per device the fixed costs dominate
  int nparts = device_count();
  while (nparts > 1 && n / nparts < 4096) nparts--;
  table_load_parts<T>(h.get(), bases, inf_flags, (uint32_t)n, precompute != 0, window_bits, nparts);
  *handle = register_handle(std::move(h));
  ZK_API_END
}

// One MSM with host scalars.  Several parts: every device uploads ITS slice of the scalars over its
// own PCIe link, reduces its base range and stores the XYZZ partial sum into the primary device's
// gather buffer; the primary device adds the partials and converts to wire bytes.
template <class T>
void table_msm_host(TableHandle<T>* h, const uint8_t* scalars, size_t n, uint8_t* out) {
  ZK_REQUIRE(scalars && out && n > 0 && n <= h->n, ZK_EARG, "msm: scalar count out of range");
  CtxScope primary(0);
  cudaStream_t st0 = stream_of(0);
  ZK_CUDA(cudaMemsetAsync(h->d_err.p, 0, sizeof(int), st0));
  const int active = h->active_parts(n);
  if (h->parts.size() == 1) {
    TablePart<T>& P = *h->parts[0];
    P.d_scalars.ensure(n * 8);
    ZK_CUDA(cudaMemcpyAsync(P.d_scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, st0));
    P.table.run(P.d_scalars.p, (uint32_t)n, h->d_result.p, h->d_out.p, st0, 0, h->d_err.p);
    P.table.join(st0);
  } else {
    h->d_gather.ensure(active);
    ZK_CUDA(cudaEventRecord(h->ready, st0));
    for (int p = 0; p < active; p++) {
      TablePart<T>& P = *h->parts[p];
      CtxScope scope(P.ctx);
      cudaStream_t st = stream_of(P.ctx);
      const uint32_t cnt = (uint32_t)std::min<size_t>(n, (size_t)P.lo + P.table.n) - P.lo;
      P.d_scalars.ensure((size_t)cnt * 8);
      ZK_CUDA(cudaStreamWaitEvent(st, h->ready, 0));
      ZK_CUDA(cudaMemcpyAsync(P.d_scalars.p, scalars + (size_t)P.lo * 32, (size_t)cnt * 32, cudaMemcpyHostToDevice, st));
      P.table.run(P.d_scalars.p, cnt, h->d_gather.p + p, nullptr, st, 0, h->d_err.p);
      P.table.join(st);
      ZK_CUDA(cudaEventRecord(P.done, st));
    }
    for (int p = 0; p < active; p++) ZK_CUDA(cudaStreamWaitEvent(st0, h->parts[p]->done, 0));
    k_sum_parts<T><<<1, 32, 0, st0>>>(h->d_gather.p, (uint32_t)active, h->d_out.p, T::RAW + T::COMP);
    ZK_CUDA(cudaGetLastError());
  }
  int err = 0;
  ZK_CUDA(cudaMemcpyAsync(out, h->d_out.p, T::RAW + T::COMP, cudaMemcpyDeviceToHost, st0));
  ZK_CUDA(cudaMemcpyAsync(&err, h->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st0));
  ZK_CUDA(cudaStreamSynchronize(st0));
  ZK_REQUIRE(err == 0, ZK_EPOINT, "msm: scalar is not canonical (>= r)");
}

template <class T>
int api_table_msm(uint64_t handle, const uint8_t* scalars, size_t n, uint8_t* out) {
  ZK_API_BEGIN
  auto* h = static_cast<TableHandle<T>*>(lookup_handle(handle, T::ID));
  table_msm_host<T>(h, scalars, n, out);
  ZK_API_END
}

// `count` MSMs over the same table with host scalars, in groups: on every device a group's scalar
// vectors are uploaded on a copy stream (into one staging buffer per MSM) while the previous group
// is being accumulated, then the group is queued and joined: ONE sort, accumulation and tail for
// the whole group.  The first group is short — just long enough to cover the upload of the rest —
// the following ones take the queue depth.
template <class T>
int api_table_msm_batch(uint64_t handle, const uint8_t* const* scalars, size_t n, size_t count, uint8_t* out) {
  ZK_API_BEGIN
  auto* h = static_cast<TableHandle<T>*>(lookup_handle(handle, T::ID));
  ZK_REQUIRE(scalars && out && count > 0 && count <= 65535 && n > 0 && n <= h->n, ZK_EARG, "msm_batch: bad arguments");
  for (size_t i = 0; i < count; i++) ZK_REQUIRE(scalars[i], ZK_EARG, "msm_batch: null scalar vector");
  constexpr size_t OUT = T::RAW + T::COMP;
  CtxScope primary(0);
  cudaStream_t st0 = stream_of(0);
  const int active = h->active_parts(n);
  const bool multi = h->parts.size() > 1;
  h->d_out.ensure(count * OUT);
  if (multi) h->d_gather.ensure(count * (size_t)active);
  ZK_CUDA(cudaMemsetAsync(h->d_err.p, 0, sizeof(int), st0));
  ZK_CUDA(cudaEventRecord(h->ready, st0));
  // pipelined for the call; on unwind every device is drained and the queue is dropped
  std::vector<std::unique_ptr<PipelineScope<T>>> scopes;
  int depth = MSM_QUEUE;
  for (int p = 0; p < active; p++) {
    TablePart<T>& P = *h->parts[p];
    CtxScope scope(P.ctx);
    if (!P.batch.copy) P.batch.ensure(0, 0);                     // the copy stream, for the scope below
    scopes.push_back(std::make_unique<PipelineScope<T>>(P.table, P.ctx, P.batch.copy));
    depth = std::min(depth, std::min(P.table.queue_cap, P.table.queue_limit));
  }
  depth = (int)std::min<size_t>((size_t)depth, count);
  size_t widest = 0;
  for (int p = 0; p < active; p++) {
    TablePart<T>& P = *h->parts[p];
    CtxScope scope(P.ctx);
    const uint32_t cnt = (uint32_t)std::min<size_t>(n, (size_t)P.lo + P.table.n) - P.lo;
    widest = std::max<size_t>(widest, cnt);
    P.batch.ensure(cnt, depth);
    P.batch.steps_timed = 0;
    if (multi) ZK_CUDA(cudaStreamWaitEvent(stream_of(P.ctx), h->ready, 0));
  }
  // Size of the first group: large enough that its computation (s per MSM, plus one tail) covers the
  // upload of everything after it (u per MSM), small enough that little upload time is exposed
  // before the first kernel:  (count - g) u <= g s + tail.  u from 25 GB/s of pinned-host bandwidth
  // (what one of 8 processes uploading at once can count on; a lone process sees ~50), s from
  // 0.38 ns per bucket addition, tail ~ 1.2 ms.
  size_t first_group = 1;
  {
    const double u = (double)widest * 32 / 25e9, sdur = (double)widest * h->parts[0]->table.cfg.W * 0.38e-9, tail = 1.2e-3;
    const double g = ((double)count * u - tail) / (sdur + u);
    if (g > 1) first_group = (size_t)g + 1;
    first_group = std::min<size_t>(first_group, (size_t)depth);
  }
  int group_id = 0;
  for (size_t g0 = 0; g0 < count; group_id++) {
    const size_t g1 = std::min(count, g0 + (g0 == 0 ? first_group : (size_t)depth));
    for (int p = 0; p < active; p++) {     // the host thread interleaves the devices group by group
      TablePart<T>& P = *h->parts[p];
      BatchStage& B = P.batch;
      CtxScope scope(P.ctx);
      cudaStream_t st = stream_of(P.ctx), cs = B.copy;
      const uint32_t cnt = (uint32_t)std::min<size_t>(n, (size_t)P.lo + P.table.n) - P.lo;
      if (g0 > 0) ZK_CUDA(cudaStreamWaitEvent(cs, B.scattered, 0));   // the staging ring is free again
      for (size_t i = g0; i < g1; i++) {
        const bool tm = B.timed && i < (size_t)BATCH_TIMED_STEPS;
        if (tm) ZK_CUDA(cudaEventRecord(B.t_c0[i], cs));
        ZK_CUDA(cudaMemcpyAsync(B.d_sc[i - g0].p, scalars[i] + (size_t)P.lo * 32, (size_t)cnt * 32, cudaMemcpyHostToDevice, cs));
        if (tm) ZK_CUDA(cudaEventRecord(B.t_c1[i], cs));
      }
      ZK_CUDA(cudaEventRecord(B.copied, cs));
      ZK_CUDA(cudaStreamWaitEvent(st, B.copied, 0));
      const bool tmg = B.timed && g0 < (size_t)BATCH_TIMED_STEPS;
      if (tmg) ZK_CUDA(cudaEventRecord(B.t_k0[g0], st));
      for (size_t i = g0; i < g1; i++) {
        // (the canonical-scalar check rides in the first digit pass)
        if (multi) P.table.run(B.d_sc[i - g0].p, cnt, h->d_gather.p + i * active + p, nullptr, st, 0, h->d_err.p);
        else P.table.run(B.d_sc[i - g0].p, cnt, nullptr, h->d_out.p + i * OUT, st, 0, h->d_err.p);
      }
      P.table.join(st, B.scattered);
      if (tmg) {
        ZK_CUDA(cudaEventRecord(B.t_k1[g0], st));
        for (size_t i = g0; i < g1 && i < (size_t)BATCH_TIMED_STEPS; i++) { B.group_of[i] = (int)g0; B.steps_timed = (int)i + 1; }
      }
    }
    g0 = g1;
  }
  for (int p = 0; p < active; p++) {
    TablePart<T>& P = *h->parts[p];
    CtxScope scope(P.ctx);
    if (multi) ZK_CUDA(cudaEventRecord(P.done, stream_of(P.ctx)));
  }
  if (multi) {
    for (int p = 0; p < active; p++) ZK_CUDA(cudaStreamWaitEvent(st0, h->parts[p]->done, 0));
    k_sum_parts<T><<<(unsigned)count, 32, 0, st0>>>(h->d_gather.p, (uint32_t)active, h->d_out.p, T::RAW + T::COMP);
    ZK_CUDA(cudaGetLastError());
  }
  int err = 0;
  ZK_CUDA(cudaMemcpyAsync(out, h->d_out.p, count * OUT, cudaMemcpyDeviceToHost, st0));
  ZK_CUDA(cudaMemcpyAsync(&err, h->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st0));
  ZK_CUDA(cudaStreamSynchronize(st0));
  for (int p = 0; p < active; p++) ZK_CUDA(cudaStreamSynchronize(h->parts[p]->batch.copy));
  ZK_REQUIRE(err == 0, ZK_EPOINT, "msm_batch: scalar is not canonical (>= r)");
  ZK_API_END
}

// Per-step timing of the last zk_g*_table_msm_batch on this handle (enable before the call):
// out[4 i + 0] = upload start, [1] = upload end of step i, [2] = first kernel, [3] = last kernel of
// the GROUP step i was joined in, in ms since the first upload started (first device's part).
// *steps = steps written.
template <class T>
int api_table_batch_timing(TableHandle<T>* h, int enable, float* out, size_t cap, size_t* steps) {
  BatchStage& B = h->parts[0]->batch;
  if (out && steps) {
    size_t k = std::min((size_t)B.steps_timed, cap / 4);
    for (size_t i = 0; i < k; i++) {
      ZK_CUDA(cudaEventElapsedTime(&out[4 * i + 0], B.t_c0[0], B.t_c0[i]));
      ZK_CUDA(cudaEventElapsedTime(&out[4 * i + 1], B.t_c0[0], B.t_c1[i]));
      ZK_CUDA(cudaEventElapsedTime(&out[4 * i + 2], B.t_c0[0], B.t_k0[B.group_of[i]]));   // the step's group
      ZK_CUDA(cudaEventElapsedTime(&out[4 * i + 3], B.t_c0[0], B.t_k1[B.group_of[i]]));
    }
    *steps = k;
  }
  for (auto& p : h->parts) p->batch.timed = enable != 0;
  return ZK_OK;
}

template <class T>
int api_table_msm_dev(uint64_t handle, const void* d_scalars, size_t n, void* d_out, void* stream) {
  ZK_API_BEGIN
  auto* h = static_cast<TableHandle<T>*>(lookup_handle(handle, T::ID));
  ZK_REQUIRE(d_scalars && d_out && n > 0 && n <= h->n, ZK_EARG, "msm_dev: bad arguments");
  cudaStream_t st = stream ? (cudaStream_t)stream : default_stream();
  h->single().run((const uint32_t*)d_scalars, (uint32_t)n, nullptr, (uint8_t*)d_out, st);
  ZK_API_END
}

template <class T>
int api_msm_oneshot(const uint8_t* bases, const uint8_t* inf_flags, const uint8_t* scalars, size_t n, uint8_t* out) {
  ZK_API_BEGIN
  ZK_REQUIRE(out, ZK_EARG, "msm: null output");
  if (n == 0) {  // empty sum = identity (curve.ml:91 folds from zero)
    memset(out, 0, T::RAW + T::COMP);
    out[0] = 0x40;
    out[T::RAW] = 0xc0;
    return ZK_OK;
  }
  ZK_REQUIRE(bases && scalars && n < (1ull << 28), ZK_EARG, "msm: bad arguments");
  TableHandle<T> h;     // one-shot: bases are parsed and uploaded for this call only, on the primary device
  table_load_parts<T>(&h, bases, inf_flags, (uint32_t)n, false, 0, 1);
  table_msm_host<T>(&h, scalars, n, out);
  ZK_API_END
}

// sum of k points given in the uncompressed wire format (the shards' partial sums), one thread
template <class T>
__global__ void k_sum_raw(const uint8_t* __restrict__ raw, uint32_t k, XYZZ<typename T::F>* __restrict__ out, int* err) {
  if (threadIdx.x || blockIdx.x) return;
  XYZZ<typename T::F> acc = XYZZ<typename T::F>::inf();
  for (uint32_t i = 0; i < k; i++) {
    Affine<typename T::F> p;
    if (T::parse(raw + (size_t)i * T::RAW, p)) { if (err) atomicExch(err, 1); continue; }
    acc.madd(p);
  }
  store_vec(out, acc);
}

// batched form for the shard gather: out[q] = sum_r points[r * batch + q], one block per q.
// Asynchronous (no place to return an error): a partial that does not parse POISONS its output
// slot (every byte 0xff, which no parser accepts) instead of being skipped.
template <class T>
__global__ void k_sum_strided(const uint8_t* __restrict__ raw, uint32_t k, uint32_t batch, uint8_t* __restrict__ out) {
  if (threadIdx.x) return;
  const uint32_t q = blockIdx.x;
  XYZZ<typename T::F> acc = XYZZ<typename T::F>::inf();
  bool bad = false;
  for (uint32_t r = 0; r < k; r++) {
    Affine<typename T::F> p;
    if (T::parse(raw + ((size_t)r * batch + q) * T::RAW, p)) { bad = true; continue; }
    acc.madd(p);
  }
  uint8_t* o = out + (size_t)q * (T::RAW + T::COMP);
  if (bad) {
    for (int i = 0; i < T::RAW + T::COMP; i++) o[i] = 0xff;
    return;
  }
  Affine<typename T::F> a = acc.to_affine();
  T::serialize(a, o);
}

// poison marker for the asynchronous single sum (see k_sum_strided)
template <class T>
__global__ void k_poison_if(const int* __restrict__ err, uint8_t* __restrict__ out) {
  if (*err == 0) return;
  for (int i = threadIdx.x; i < T::RAW + T::COMP; i += blockDim.x) out[i] = 0xff;
}

template <class T>
int api_sum_strided_dev(const void* d_points, size_t k, size_t batch, void* d_out, void* stream) {
  ZK_API_BEGIN
  ZK_REQUIRE(d_points && d_out && k > 0 && k <= 4096 && batch > 0 && batch <= 65535, ZK_EARG, "sum_strided_dev: bad arguments");
  cudaStream_t st = stream ? (cudaStream_t)stream : default_stream();
  k_sum_strided<T><<<(unsigned)batch, 32, 0, st>>>((const uint8_t*)d_points, (uint32_t)k, (uint32_t)batch, (uint8_t*)d_out);
  ZK_CUDA(cudaGetLastError());
  ZK_API_END
}

template <class T>
int api_sum_dev(const void* d_points, size_t k, void* d_out, void* stream) {
  ZK_API_BEGIN
  ZK_REQUIRE(d_points && d_out && k > 0 && k <= 4096, ZK_EARG, "sum_dev: bad arguments");
  cudaStream_t st = stream ? (cudaStream_t)stream : default_stream();
  static thread_local DevBuf<XYZZ<typename T::F>> scratch;
  static thread_local DevBuf<int> flag;
  scratch.ensure(1);
  flag.ensure(1);
  ZK_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
  k_sum_raw<T><<<1, 32, 0, st>>>((const uint8_t*)d_points, (uint32_t)k, scratch.p, flag.p);
  finalize_points<T>(scratch.p, 1, (uint8_t*)d_out, st);
  k_poison_if<T><<<1, 32, 0, st>>>(flag.p, (uint8_t*)d_out);
  ZK_API_END
}

template <class T>
int api_sum(const uint8_t* points, size_t k, uint8_t* out) {
  ZK_API_BEGIN
  ZK_REQUIRE(points && out && k > 0 && k <= 4096, ZK_EARG, "sum: bad arguments");
  cudaStream_t st = default_stream();
  DevBuf<uint8_t> d_in(k * T::RAW), d_out(T::RAW + T::COMP);
  DevBuf<XYZZ<typename T::F>> d_acc(1);
  DevBuf<int> d_err(1);
  ZK_CUDA(cudaMemcpyAsync(d_in.p, points, k * T::RAW, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
  k_sum_raw<T><<<1, 32, 0, st>>>(d_in.p, (uint32_t)k, d_acc.p, d_err.p);
  finalize_points<T>(d_acc.p, 1, d_out.p, st);
  int err = 0;
  ZK_CUDA(cudaMemcpyAsync(out, d_out.p, T::RAW + T::COMP, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(err == 0, ZK_EPOINT, "sum: point not canonical or not on the curve");
  ZK_API_END
}

// out[i] = scalars[i] * generator, uncompressed
template <class T>
__global__ void __launch_bounds__(128)
k_fixed_base(const uint32_t* __restrict__ scalars, uint32_t n, XYZZ<typename T::F>* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[8];
  for (int j = 0; j < 8; j++) k[j] = scalars[8 * (size_t)i + j];
  XYZZ<typename T::F> g = XYZZ<typename T::F>::from_affine(T::generator());
  XYZZ<typename T::F> r = scalar_mul(g, k);
  store_vec(&out[i], r);
}
template <class T>
__global__ void __launch_bounds__(128)
k_serialize_raw(const Affine<typename T::F>* __restrict__ pts, uint32_t n, uint8_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t buf[T::RAW + T::COMP];
  Affine<typename T::F> p = load_vec_rw(&pts[i]);
  T::serialize(p, buf);
  for (int j = 0; j < T::RAW; j++) out[(size_t)i * T::RAW + j] = buf[j];
}

template <class T>
int api_fixed_base_mul(const uint8_t* scalars, size_t n, uint8_t* out) {
  ZK_API_BEGIN
  ZK_REQUIRE(scalars && out && n > 0 && n < (1ull << 28), ZK_EARG, "fixed_base_mul: bad arguments");
  typedef typename T::F F;
  cudaStream_t st = default_stream();
  DevBuf<uint32_t> d_s(n * 8);
  DevBuf<XYZZ<F>> d_x(n);
  DevBuf<Affine<F>> d_a(n);
  DevBuf<uint8_t> d_o(n * T::RAW);
  DevBuf<int> d_err(1);
  ZK_CUDA(cudaMemcpyAsync(d_s.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
  k_check_scalars<<<cdiv(n, 256), 256, 0, st>>>(d_s.p, (uint32_t)n, d_err.p);
  k_fixed_base<T><<<cdiv(n, 128), 128, 0, st>>>(d_s.p, (uint32_t)n, d_x.p);
  k_batch_to_affine<F, 16><<<cdiv(cdiv(n, 16), 128), 128, 0, st>>>(d_x.p, (uint32_t)n, d_a.p);
  k_serialize_raw<T><<<cdiv(n, 128), 128, 0, st>>>(d_a.p, (uint32_t)n, d_o.p);
  ZK_CUDA(cudaGetLastError());
  int err = 0;
  ZK_CUDA(cudaMemcpyAsync(out, d_o.p, n * T::RAW, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(err == 0, ZK_EPOINT, "fixed_base_mul: scalar is not canonical (>= r)");
  ZK_API_END
}

}  // namespace zk
