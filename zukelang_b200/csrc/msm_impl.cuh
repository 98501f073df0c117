// Host orchestration of the MSM pipeline (template bodies; included by msm_g1.cu / msm_g2.cu).
#pragma once
#include <algorithm>
#include "msm.cuh"

#ifndef ZK_ACC_VARIANT_DEFAULT
#define ZK_ACC_VARIANT_DEFAULT 9
#endif

namespace zk {

static inline int log2_ceil(uint32_t n) {
  int l = 0;
  while ((1ull << l) < n) l++;
  return l;
}

template <class T>
MsmConfig BaseTable<T>::choose_config(uint32_t n, bool precompute, int force_c) {
  MsmConfig cfg;
  int lg = log2_ceil(n < 2 ? 2 : n);
  // precomputed tables: one shared bucket set of 2^(c-1) buckets against n*W entries; the bucket
  // reduction costs ~100 Fp products per bucket, so keep ~30+ entries per bucket
  int c = precompute ? lg - 1 : lg - 3;
  if (c > 17) c = 17;
  if (!precompute && c > 16) c = 16;
  if (c < 4) c = 4;
  int env_c = env_int(precompute ? "ZKB200_WINDOW_BITS_PRE" : "ZKB200_WINDOW_BITS", 0);
  if (env_c > 0) c = env_c;
  if (force_c > 0) c = force_c;
  if (c < 2) c = 2;
  if (c > 22) c = 22;
  cfg.c = c;
  // scalars are folded into [0, (r-1)/2] < 2^254 by k_digits; the top window must leave room for
  // the incoming carry: its data bits must not exceed c - 1
  auto windows = [](int cc) { int w = (254 + cc - 1) / cc; if (254 - (w - 1) * cc > cc - 1) w++; return w; };
  // with one shared bucket set, a top window holding only a few data bits would pile every point
  // into a handful of buckets: step down to a window size whose top window is reasonably full
  if (precompute && force_c <= 0 && env_c <= 0)
    while (c > 4 && 254 - (windows(c) - 1) * c < c - 6) c--;
  cfg.c = c;
  cfg.W = windows(c);
  cfg.nwb = precompute ? 1 : cfg.W;
  cfg.B = 1u << (c - 1);
  cfg.S = 1;
  int L = 4;
  int env_l = env_int("ZKB200_REDUCE_CHUNK", 0);
  if (env_l > 0) L = env_l;
  while ((uint32_t)L > cfg.B) L >>= 1;
  cfg.L = L;
  return cfg;
}

template <class T>
void BaseTable<T>::load(const uint8_t* host_raw, const uint8_t* host_inf, uint32_t n_, bool precompute,
                        int force_c, cudaStream_t st) {
  ZK_REQUIRE(n_ > 0, ZK_EARG, "empty base table");
  n = n_;
  precomputed = precompute;
  cfg = choose_config(n, precompute, force_c);
  size_t nwin = precompute ? cfg.W : 1;
  ZK_REQUIRE((uint64_t)nwin * n < (1ull << 31), ZK_EARG, "table too large for 31-bit point indices");
  ZK_REQUIRE((uint64_t)n * cfg.W < (1ull << 32), ZK_EARG, "table too large for 32-bit entry positions");
  pts.alloc(nwin * n);
  skip.alloc(n);
  DevBuf<uint8_t> d_raw((size_t)n * T::RAW);
  DevBuf<uint8_t> d_inf;
  DevBuf<int> d_err(1);
  ZK_CUDA(cudaMemcpyAsync(d_raw.p, host_raw, d_raw.bytes(), cudaMemcpyHostToDevice, st));
  if (host_inf) {
    d_inf.alloc(n);
    ZK_CUDA(cudaMemcpyAsync(d_inf.p, host_inf, n, cudaMemcpyHostToDevice, st));
  }
  ZK_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
  k_parse_bases<T><<<cdiv(n, 128), 128, 0, st>>>(d_raw.p, d_inf.p, n, pts.p, skip.p, d_err.p);
  ZK_CUDA(cudaGetLastError());
  int err = 0;
  ZK_CUDA(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(err == 0, ZK_EPOINT, "base point not canonical or not on the curve");
  build_tables(st);
}

template <class T>
void BaseTable<T>::build_tables(cudaStream_t st) {
  if (precomputed) {
    DevBuf<XYZZ<F>> scratch(n);
    for (int w = 1; w < cfg.W; w++) {
      k_precompute_shift<F><<<cdiv(n, 128), 128, 0, st>>>(pts.p + (size_t)(w - 1) * n, n, cfg.c, scratch.p);
      k_batch_to_affine<F, 16><<<cdiv(cdiv(n, 16), 128), 128, 0, st>>>(scratch.p, n, pts.p + (size_t)w * n);
    }
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaStreamSynchronize(st));
  }
  {
    // 9 (default): mixed add with paired products + cp.async staging, 2 blocks/SM; 8: paired, direct
    // loads; 5: plain mixed add + staging; 4: plain, direct loads, 2 blocks/SM; 1: plain, 4 blocks/SM
    constexpr bool is_g1 = sizeof(F) == sizeof(Fp);
    acc_variant = env_int(is_g1 ? "ZKB200_ACC_VARIANT" : "ZKB200_ACC_VARIANT_G2", ZK_ACC_VARIANT_DEFAULT);
    if (acc_variant != 1 && acc_variant != 4 && acc_variant != 5 && acc_variant != 8 && acc_variant != 9)
      acc_variant = ZK_ACC_VARIANT_DEFAULT;
    int per_sm = acc_occupancy();
    if (per_sm < 1) per_sm = 1;
    acc_blocks = (uint32_t)per_sm * (uint32_t)sm_count();
  }
  // partial pieces and heavy-bucket queue of ONE accumulation launch (whatever the number of MSMs in it)
  partial.alloc(2 * (size_t)acc_blocks * ACC_THREADS);
  heavy.alloc((size_t)acc_blocks * ACC_THREADS / 4 + 2);
  open_bucket.alloc((size_t)acc_blocks * ACC_THREADS);
  queued = 0;
  queue_cap = 0;
  queue_limit = 1;
  pipelined = false;
  ensure_queue(1);
  if (env_int("ZKB200_PIPELINE", 0) != 0) set_pipelined(true);
  ZK_CUDA(cudaStreamSynchronize(st));
}

// Queues one MSM; a table that is not pipelined runs it at once.
template <class T>
void BaseTable<T>::run(const uint32_t* d_scalars, uint32_t count, XYZZ<F>* d_result, uint8_t* d_out_bytes,
                       cudaStream_t st, uint32_t first, int* d_err) {
  ZK_REQUIRE(count > 0 && (uint64_t)first + count <= n, ZK_EARG, "scalar range exceeds the base table");
  ZK_REQUIRE(d_scalars, ZK_EARG, "null scalar vector");
  if (pending) tail(st);              // a sort_accumulate() whose tail was never asked for
  if (queued >= std::min(queue_cap, queue_limit)) join(st);
  const int slot = queued;
  slots.scalars[slot] = d_scalars;
  slots.err[slot] = d_err;
  slots.count[slot] = count;
  slots.first[slot] = first;
  outs.result[slot] = d_result ? d_result : window_sums.p + (size_t)slot * (cfg.nwb + 1) + cfg.nwb;
  outs.bytes[slot] = d_out_bytes;
  queued++;
  if (!pipelined) join(st);
}

// the accumulation kernel selected by acc_variant: occupancy query and launch
template <class T>
template <class Fn>
void BaseTable<T>::acc_dispatch(Fn&& fn) {
  constexpr size_t SM = acc_stage_bytes<F>();
  switch (acc_variant) {
    case 1: fn(k_accumulate<F, 4, false, false>, (size_t)0); break;
    case 4: fn(k_accumulate<F, 2, false, false>, (size_t)0); break;
    case 5: fn(k_accumulate<F, 2, true, false>, SM); break;
    case 8: fn(k_accumulate<F, 2, false, true>, (size_t)0); break;
    default: fn(k_accumulate<F, 2, true, true>, SM);
  }
}
template <class T>
int BaseTable<T>::acc_occupancy() {
  int per_sm = 0;
  acc_dispatch([&](auto kern, size_t smem) {
    if (smem) ZK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ZK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ACC_THREADS, smem));
  });
  return per_sm;
}
template <class T>
void BaseTable<T>::acc_launch(uint32_t grid, uint32_t nbuckets, cudaStream_t st) {
  acc_dispatch([&](auto kern, size_t smem) {
    kern<<<grid, ACC_THREADS, smem, st>>>(pts.p, entries.p, offsets.p, bucket_sums.p, partial.p, open_bucket.p, nbuckets, (uint32_t)pts.n);
  });
}

template <class T>
void BaseTable<T>::ensure_queue(int slots) {
  // the queued MSMs are sorted and accumulated as ONE entry list: keep it below 2^29 entries
  // (2 GB of point indices; entry positions are 32-bit)
  const uint64_t per_msm = (uint64_t)n * cfg.W;
  const int most = (int)std::max<uint64_t>(1, (1ull << 29) / per_msm);
  slots = std::max(1, std::min(std::min(slots, most), MSM_QUEUE));
  if (slots <= queue_cap) return;
  ZK_REQUIRE(queued == 0 && pending == 0, ZK_EARG, "cannot resize the queue while MSMs are queued");
  const uint32_t nb = cfg.nbuckets();
  const size_t nbq = (size_t)slots * nb;
  const uint32_t cpw = cfg.B / cfg.L;
  counts.alloc(nbq);
  ZK_CUDA(cudaMemset(counts.p, 0, counts.bytes()));   // kept zero between joins by k_scan_apply
  ZK_CUDA(cudaDeviceSynchronize());
  cursor.alloc(nbq);
  tile_sums.alloc(cdiv(nbq, SCAN_TILE) + 1);
  offsets.alloc(nbq + 1);
  entries.alloc((size_t)slots * per_msm);
  bucket_sums.alloc(nbq);
  chunk_out.alloc((size_t)slots * cfg.nwb * cpw);
  tree_tmp.alloc((size_t)slots * cfg.nwb * cdiv(cpw, TAIL_THREADS) + 1);
  window_sums.alloc((size_t)slots * (cfg.nwb + 1));
  queue_cap = slots;
}

template <class T>
void BaseTable<T>::set_pipelined(bool on, int depth) {
  if (on) {
    if (depth <= 0) depth = env_int("ZKB200_QUEUE", MSM_QUEUE);
    if (queued == 0 && pending == 0) ensure_queue(depth);   // cannot grow under queued MSMs: the current capacity stays
    queue_limit = std::max(std::max(1, queued), std::min(depth, queue_cap));
  }
  pipelined = on;
}

// Everything queued, as one launch sequence: sort and accumulate all Q MSMs over Q * nb buckets, then
// the batched tail (partial fix-up, bucket reduction, window combine, affine conversion).
template <class T>
void BaseTable<T>::join(cudaStream_t st, cudaEvent_t after_scatter) {
  sort_accumulate(st, after_scatter);
  tail(st);
}

template <class T>
void BaseTable<T>::sort_accumulate(cudaStream_t st, cudaEvent_t after_scatter) {
  if (pending) tail(st);
  if (queued == 0) {
    if (after_scatter) ZK_CUDA(cudaEventRecord(after_scatter, st));
    return;
  }
  const uint32_t nb = cfg.nbuckets();
  const int Q = queued;
  const uint32_t nbq = (uint32_t)Q * nb;
  ProfRec* rec = nullptr;
  if (profile) {
    rec = &prof[prof_joins % PROF_RING];
    rec->msms = Q;
    prof_joins++;
  }
  auto mark = [&](int i) { if (rec) ZK_CUDA(cudaEventRecord(rec->ev[i], st)); };
  uint32_t most = 0;
  uint64_t total = 0;
  for (int q = 0; q < Q; q++) { most = std::max(most, slots.count[q]); total += (uint64_t)slots.count[q] * cfg.W; }
  // ---- sort: five launches for all Q MSMs, no memsets (the histogram comes back zeroed) -----------
  mark(0);
  k_digits<false><<<dim3(cdiv(most, 256), Q), 256, 0, st>>>(slots, skip.p, n, cfg, counts.p, nullptr, 0);
  const uint32_t ntiles = cdiv(nbq, SCAN_TILE);
  k_scan_tile_sums<<<ntiles, SCAN_THREADS, 0, st>>>(counts.p, nbq, tile_sums.p);
  k_scan_apply<<<ntiles, SCAN_THREADS, 0, st>>>(counts.p, nbq, tile_sums.p, offsets.p, cursor.p);
  k_digits<true><<<dim3(cdiv(most, 256), Q), 256, 0, st>>>(slots, skip.p, n, cfg, cursor.p, entries.p, (uint32_t)entries.n);
  if (after_scatter) ZK_CUDA(cudaEventRecord(after_scatter, st));
  mark(1);
  // ---- accumulate: one wave at most; for small inputs fewer blocks, so that a slice still holds
  // >= 16 entries (otherwise the partial fix-up, not the mixed adds, would end up doing the additions)
  uint32_t grid = (uint32_t)std::min<uint64_t>(acc_blocks, cdiv(total, (uint64_t)16 * ACC_THREADS));
  if (grid < 1) grid = 1;
  acc_launch(grid, nbq, st);
  mark(2);
  ZK_CUDA(cudaGetLastError());
  pending = Q;
  pending_threads = grid * ACC_THREADS;
  queued = 0;
}

template <class T>
void BaseTable<T>::tail(cudaStream_t st, int share) {
  if (pending == 0) return;
  const uint32_t nb = cfg.nbuckets();
  const int Q = pending;
  const uint32_t nbq = (uint32_t)Q * nb;
  const uint32_t acc_threads = pending_threads;
  ProfRec* rec = profile && prof_joins ? &prof[(prof_joins - 1) % PROF_RING] : nullptr;
  auto mark = [&](int i) { if (rec) ZK_CUDA(cudaEventRecord(rec->ev[i], st)); };
  // ---- buckets that were split over several accumulation slices ...
  ZK_CUDA(cudaMemsetAsync(heavy.p, 0, sizeof(uint32_t), st));   // heavy[0] = 0
  k_fix_partials<F><<<cdiv(acc_threads, 128), 128, 0, st>>>(offsets.p, bucket_sums.p, partial.p, open_bucket.p, nbq, acc_threads, heavy.p);
  {
    const int ht = sizeof(XYZZ<F>) > 192 ? 128 : 256;   // 48 KB of shared memory either way
    k_fix_heavy<F><<<sm_count(), ht, ht * sizeof(XYZZ<F>), st>>>(offsets.p, bucket_sums.p, partial.p, nbq, acc_threads, heavy.p);
  }
  // ... bucket reduction.  Chunk width L of the running sum: a chunk costs 2 L - 1 additions plus a
  // ~22-operation scalar multiplication, all dependent, so narrow chunks mean a short chain but
  // 7.3 additions per bucket (L = 4) against 3.3 (L = 16) or 2.7 (L = 32).  A lone warp already
  // keeps its scheduler's multiplier slot busy, so the kernel is latency-bound up to one warp per
  // scheduler and throughput-bound beyond: take the narrowest L that keeps the chunk threads of all
  // queued MSMs within that (halved when another table's tail runs beside this one).
  MsmConfig rc = cfg;
  if (env_int("ZKB200_REDUCE_CHUNK", 0) <= 0) {
    const uint64_t target = (uint64_t)sm_count() * 4 * 32 / (uint64_t)std::max(1, share);
    int L = 4;
    while (L < 32 && (uint64_t)Q * nb / L > target) L <<= 1;
    rc.L = L;
    while ((uint32_t)rc.L > rc.B) rc.L >>= 1;
  }
  const uint32_t cpw = rc.B / rc.L;
  k_reduce_chunks<F><<<dim3(cdiv((size_t)cpw * cfg.nwb, TAIL_THREADS), 1, Q), TAIL_THREADS, 0, st>>>(bucket_sums.p, offsets.p, rc,
                                                                                                     chunk_out.p);
  {
    // sum tree per (window, queued MSM): cpw -> ceil(cpw / TAIL_THREADS) -> ... -> 1
    const XYZZ<F>* src = chunk_out.p;
    size_t src_stride = (size_t)cfg.nwb * cpw;
    uint32_t cnt = cpw;
    XYZZ<F>* bufs[2] = {tree_tmp.p, chunk_out.p};   // ping-pong (chunk_out is dead after level 1)
    size_t buf_stride[2] = {(size_t)cfg.nwb * cdiv(cpw, TAIL_THREADS), (size_t)cfg.nwb * cpw};
    int which = 0;
    while (true) {
      uint32_t blocks = cdiv(cnt, TAIL_THREADS);
      XYZZ<F>* dst = blocks == 1 ? window_sums.p : bufs[which];
      size_t dst_stride = blocks == 1 ? (size_t)(cfg.nwb + 1) : buf_stride[which];
      k_reduce_tree<F><<<dim3(blocks, cfg.nwb, Q), TAIL_THREADS, TAIL_THREADS * sizeof(XYZZ<F>), st>>>(src, cnt, src_stride,
                                                                                                      dst, dst_stride);
      if (blocks == 1) break;
      src = dst;
      src_stride = dst_stride;
      cnt = blocks;
      which ^= 1;
    }
  }
  mark(3);
  k_combine_finalize<T><<<Q, 32, 0, st>>>(window_sums.p, cfg, outs);
  mark(4);
  ZK_CUDA(cudaGetLastError());
  pending = 0;
}

template <class T>
void BaseTable<T>::set_profile(bool on) {
  if (on && prof.empty()) {
    prof.resize(PROF_RING);
    for (auto& r : prof)
      for (auto& e : r.ev) ZK_CUDA(cudaEventCreate(&e));
  }
  if (on && !profile) prof_joins = 0;
  profile = on;
}

template <class T>
void BaseTable<T>::stage_ms(float out[4]) {
  for (int i = 0; i < 4; i++) out[i] = 0.f;
  if (prof_joins == 0) return;
  const ProfRec& r = prof[(prof_joins - 1) % PROF_RING];
  for (int i = 0; i < 4; i++) ZK_CUDA(cudaEventElapsedTime(&out[i], r.ev[i], r.ev[i + 1]));
}

template <class T>
void BaseTable<T>::stage_totals(float out[4], uint64_t* msms, uint64_t* joins) {
  for (int i = 0; i < 4; i++) out[i] = 0.f;
  const uint64_t cnt = std::min<uint64_t>(prof_joins, PROF_RING);
  uint64_t m = 0;
  for (uint64_t j = 0; j < cnt; j++) {
    const ProfRec& r = prof[j];
    for (int i = 0; i < 4; i++) {
      float ms = 0.f;
      ZK_CUDA(cudaEventElapsedTime(&ms, r.ev[i], r.ev[i + 1]));
      out[i] += ms;
    }
    m += (uint64_t)r.msms;
  }
  *msms = m;
  *joins = cnt;
}

template <class T>
BaseTable<T>::~BaseTable() {
  for (auto& r : prof)
    for (auto& e : r.ev)
      if (e) cudaEventDestroy(e);
}

template <class T>
size_t BaseTable<T>::device_bytes() const {
  return pts.bytes() + skip.bytes() + counts.bytes() + offsets.bytes() + cursor.bytes() + tile_sums.bytes() +
         entries.bytes() + heavy.bytes() + bucket_sums.bytes() + partial.bytes() + chunk_out.bytes() + tree_tmp.bytes() + window_sums.bytes();
}

template <class T>
void finalize_points(const XYZZ<typename T::F>* d_results, int count, uint8_t* d_out, cudaStream_t st) {
  k_finalize<T><<<cdiv(count, 32), 32, 0, st>>>(d_results, count, d_out);
  ZK_CUDA(cudaGetLastError());
}

}  // namespace zk
