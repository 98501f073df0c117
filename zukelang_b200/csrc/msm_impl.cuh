// Host orchestration of the MSM pipeline (template bodies; included by msm_g1.cu / msm_g2.cu).
#pragma once
#include "msm.cuh"

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
  int c = precompute ? lg : lg - 3;
  if (c > 16) c = 16;
  if (c < 4) c = 4;
  int env_c = env_int(precompute ? "ZKB200_WINDOW_BITS_PRE" : "ZKB200_WINDOW_BITS", 0);
  if (env_c > 0) c = env_c;
  if (force_c > 0) c = force_c;
  if (c < 2) c = 2;
  if (c > 22) c = 22;
  cfg.c = c;
  cfg.W = (256 + c - 1) / c;
  cfg.nwb = precompute ? 1 : cfg.W;
  cfg.B = 1u << (c - 1);
  // segments: enough threads to fill the machine, but keep >= 16 entries per segment on average
  uint64_t entries = (uint64_t)n * cfg.W;
  uint32_t nb = cfg.nbuckets();
  uint64_t avg = entries / nb;
  int want = (int)((148ull * 768 + nb - 1) / nb);
  int cap = (int)(avg / 16);
  int S = want < cap ? want : cap;
  if (S < 1) S = 1;
  if (S > 64) S = 64;
  int env_s = env_int("ZKB200_SEGMENTS", 0);
  if (env_s > 0) S = env_s;
  cfg.S = S;
  int L = precompute ? 8 : 16;
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
void BaseTable<T>::load_device_affine(const Affine<F>* d_affine, uint32_t n_, bool precompute, int force_c,
                                      cudaStream_t st) {
  ZK_REQUIRE(n_ > 0, ZK_EARG, "empty base table");
  n = n_;
  precomputed = precompute;
  cfg = choose_config(n, precompute, force_c);
  size_t nwin = precompute ? cfg.W : 1;
  ZK_REQUIRE((uint64_t)nwin * n < (1ull << 31), ZK_EARG, "table too large for 31-bit point indices");
  pts.alloc(nwin * n);
  skip.alloc(n);
  ZK_CUDA(cudaMemcpyAsync(pts.p, d_affine, (size_t)n * sizeof(Affine<F>), cudaMemcpyDeviceToDevice, st));
  k_mark_skip<F><<<cdiv(n, 256), 256, 0, st>>>(pts.p, n, skip.p);
  ZK_CUDA(cudaGetLastError());
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
  uint32_t nb = cfg.nbuckets();
  counts.alloc(nb);
  offsets.alloc(nb + 1);
  cursor.alloc(nb);
  tile_sums.alloc(cdiv(nb, SCAN_TILE) + 1);
  entries.alloc((size_t)n * cfg.W);
  parts.alloc((size_t)nb * cfg.S);
  chunk_out.alloc((size_t)cfg.nwb * (cfg.B / cfg.L));
  window_sums.alloc(cfg.nwb);
}

template <class T>
void BaseTable<T>::run(const uint32_t* d_scalars, uint32_t count, XYZZ<F>* d_result, cudaStream_t st) {
  ZK_REQUIRE(count > 0 && count <= n, ZK_EARG, "scalar count exceeds the base table");
  const uint32_t nb = cfg.nbuckets();
  if (profile && !ev[0])
    for (auto& e : ev) ZK_CUDA(cudaEventCreate(&e));
  auto mark = [&](int i) { if (profile) ZK_CUDA(cudaEventRecord(ev[i], st)); };
  mark(0);
  ZK_CUDA(cudaMemsetAsync(counts.p, 0, nb * sizeof(uint32_t), st));
  k_digits<false><<<cdiv(count, 256), 256, 0, st>>>(d_scalars, skip.p, count, n, cfg, counts.p, nullptr);
  uint32_t ntiles = cdiv(nb, SCAN_TILE);
  k_scan_tile_sums<<<ntiles, SCAN_THREADS, 0, st>>>(counts.p, nb, tile_sums.p);
  k_scan_spine<<<1, 1024, 0, st>>>(tile_sums.p, ntiles);
  k_scan_apply<<<ntiles, SCAN_THREADS, 0, st>>>(counts.p, nb, tile_sums.p, offsets.p, cursor.p);
  k_digits<true><<<cdiv(count, 256), 256, 0, st>>>(d_scalars, skip.p, count, n, cfg, cursor.p, entries.p);
  mark(1);
  uint32_t nthreads = nb * cfg.S;
  k_accumulate<F><<<cdiv(nthreads, 128), 128, 0, st>>>(pts.p, entries.p, offsets.p, parts.p, nb, cfg.S);
  mark(2);
  uint32_t cpw = cfg.B / cfg.L;
  k_reduce_chunks<F><<<cdiv((size_t)cpw * cfg.nwb, 128), 128, 0, st>>>(parts.p, cfg, chunk_out.p);
  k_reduce_tree<F><<<cfg.nwb, 128, 128 * sizeof(XYZZ<F>), st>>>(chunk_out.p, cpw, window_sums.p);
  mark(3);
  k_horner<F><<<1, 32, 0, st>>>(window_sums.p, cfg, d_result);
  mark(4);
  ZK_CUDA(cudaGetLastError());
}

template <class T>
void BaseTable<T>::stage_ms(float out[4]) {
  for (int i = 0; i < 4; i++) {
    out[i] = 0.f;
    if (ev[0]) ZK_CUDA(cudaEventElapsedTime(&out[i], ev[i], ev[i + 1]));
  }
}

template <class T>
BaseTable<T>::~BaseTable() {
  for (auto& e : ev)
    if (e) cudaEventDestroy(e);
}

template <class T>
size_t BaseTable<T>::device_bytes() const {
  return pts.bytes() + skip.bytes() + counts.bytes() + offsets.bytes() + cursor.bytes() + tile_sums.bytes() +
         entries.bytes() + parts.bytes() + chunk_out.bytes() + window_sums.bytes();
}

template <class T>
void finalize_points(const XYZZ<typename T::F>* d_results, int count, uint8_t* d_out, cudaStream_t st) {
  k_finalize<T><<<cdiv(count, 32), 32, 0, st>>>(d_results, count, d_out);
  ZK_CUDA(cudaGetLastError());
}

}  // namespace zk
