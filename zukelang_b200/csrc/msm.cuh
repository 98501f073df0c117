// Pippenger multi-scalar multiplication for BLS12-381 G1 / G2 on sm_100a.
//
// Replaces the reference's "MSM" primitives — the left folds of scalar_mul + add
// at /root/reference/src/lib/zk/curve.ml:91 (sum_map), :94-103 (dot), :112-118
// (apply_powers) and the double loop of src/groth16/groth16.ml:116-121
// (sum_apply_powers) — by one bucket-method MSM per proof element.  The result
// is the same group element, hence the same serialised bytes.
//
// Pipeline (all on the caller's stream, no host round trips).  A pipelined table QUEUES its MSMs
// (scalar pointer, range, output slots) and join() runs everything queued — up to 32 MSMs over the
// same bases — as ONE launch sequence: queued MSM q owns the buckets [q * nb, (q + 1) * nb), so the
// Q MSMs are one counting sort and one balanced accumulation over Q * nb buckets:
//   1. k_digits<COUNT>    scalar -> (negate if > (r-1)/2) -> signed c-bit digits -> bucket histogram
//   2. k_scan_*           exclusive prefix sum of the histogram
//   3. k_digits<SCATTER>  (bucket, point) pairs placed by counting sort
//   4. k_accumulate       persistent, exactly balanced: thread t folds slice t of the sorted list
//                         with XYZZ mixed adds (paired interleaved products, dedicated squares)
//      k_fix_partials / k_fix_heavy   buckets split over several slices: add up their pieces
//   5. k_reduce_chunks    running-sum reduction of L buckets per thread        } the latency-bound
//      k_reduce_tree      per-window tree sum of the chunk results             } tail: one batched
//   6. k_combine_finalize window combine (a copy when precomputed), affine     } launch for all the
//                         conversion, wire bytes                               } queued MSMs
// Batching the sort and the accumulation (not only the tail) matters for small shards (2^17 points
// per GPU on 8 GPUs): the at most 2 T partial bucket pieces of the T accumulation threads, the
// five launches and their gaps are paid once per join instead of once per MSM.
//
// A BaseTable built with `precompute` holds 2^(c*w) * P_i for every window w, so all windows
// share ONE set of 2^(c-1) buckets: the bucket reduction shrinks W times and the window combine
// disappears, at the price of W times the table bytes — cheap against 180 GB of HBM3e.
#pragma once
#include <exception>
#include <vector>
#include "common.cuh"

namespace zk {

struct MsmConfig {
  int c;        // window bits
  int W;        // number of windows  = ceil(256 / c)
  int nwb;      // bucket windows: 1 when precomputed, else W
  uint32_t B;   // buckets per window = 2^(c-1)
  int S;        // always 1 (the balanced accumulation has no per-bucket segments); kept in zk_table_info
  int L;        // buckets per thread in the chunk reduction
  __host__ __device__ uint32_t nbuckets() const { return (uint32_t)nwb * B; }
};

int env_int(const char* name, int dflt);

// ------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------
#ifdef __CUDACC__

// c bits of the 256-bit little-endian scalar k starting at bit `pos`
__device__ __forceinline__ uint32_t scalar_bits(const uint32_t* k, int pos, int c) {
  int limb = pos >> 5, sh = pos & 31;
  if (limb >= 8) return 0;
  uint64_t lo = k[limb];
  uint64_t hi = (limb + 1 < 8) ? k[limb + 1] : 0;
  return (uint32_t)(((hi << 32) | lo) >> sh) & ((1u << c) - 1);
}

// The MSMs of one join (kernel parameter, by value): MSM q reads count[q] canonical scalars at
// scalars[q] and applies them to the table points [first[q], first[q] + count[q]) (a table may hold
// several key queries back to back); err[q] (nullable) is set when a scalar is >= r.
constexpr int MSM_QUEUE = 32;   // most MSMs one join can take (slot tables are kernel parameters)
struct QueueSlots {
  const uint32_t* scalars[MSM_QUEUE];
  int* err[MSM_QUEUE];
  uint32_t count[MSM_QUEUE];
  uint32_t first[MSM_QUEUE];
};

// Steps 1 and 3.  Signed-digit recoding.  A scalar k > (r-1)/2 is first replaced by r - k with
// every digit sign flipped (k P = (r - k)(-P)), which bounds the recoded value by 2^254 and saves
// a window for c = 17, 19, 20.  Digits d lie in [-(2^(c-1) - 1), 2^(c-1)] with a carry into the
// next window; zero digits (and identity bases) emit nothing.
// entry = point index (w * stride + i when precomputed, else i) | sign << 31;
// blockIdx.y = queued MSM (its buckets start at blockIdx.y * nbuckets), stride = points per window
// of the table.
template <bool SCATTER>
__global__ void __launch_bounds__(256)
k_digits(QueueSlots q, const uint8_t* __restrict__ skip, uint32_t stride, MsmConfig cfg,
         uint32_t* __restrict__ counts_or_cursor, uint32_t* __restrict__ entries, uint32_t entries_cap) {
  const int z = blockIdx.y;
  const uint32_t n = q.count[z], first = q.first[z];
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* __restrict__ scalars = q.scalars[z];
  int* err = q.err[z];
  uint32_t k[8];
  const uint4* sp = reinterpret_cast<const uint4*>(scalars + 8 * (size_t)i);
  uint4 a = __ldg(sp), b = __ldg(sp + 1);
  k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
  if (!SCATTER && err) {   // host-facing calls: a scalar >= r is an error (ZK_EPOINT), never reduced silently
    uint32_t m[8];
#pragma unroll
    for (int j = 0; j < 8; j++) m[j] = FrParams::mod(j);
    if (Fr::geq_raw(k, m)) { atomicExch(err, 1); return; }
  }
  if (SCATTER && err) {    // the count pass skipped this scalar: skip it here too
    uint32_t m[8];
#pragma unroll
    for (int j = 0; j < 8; j++) m[j] = FrParams::mod(j);
    if (Fr::geq_raw(k, m)) return;
  }
  if (skip && skip[first + i]) return;
  if ((k[0] | k[1] | k[2] | k[3] | k[4] | k[5] | k[6] | k[7]) == 0) return;
  uint32_t flip = 0;
  {
    uint32_t h[8];
#pragma unroll
    for (int j = 0; j < 8; j++) h[j] = FrParams::half(j);
    if (!Fr::geq_raw(h, k)) {  // k > (r-1)/2  ->  k = r - k
      flip = 1;
      k[0] = ptx::sub_cc(FrParams::mod(0), k[0]);
#pragma unroll
      for (int j = 1; j < 8; j++) k[j] = ptx::subc_cc(FrParams::mod(j), k[j]);
    }
  }
  uint32_t carry = 0;
  const uint32_t half = cfg.B;  // 2^(c-1)
  const uint32_t zbase = (uint32_t)z * cfg.nbuckets();
  for (int w = 0; w < cfg.W; w++) {
    uint32_t d = scalar_bits(k, w * cfg.c, cfg.c) + carry;
    uint32_t neg = flip;
    if (d > half) { d = (1u << cfg.c) - d; neg ^= 1; carry = 1; } else carry = 0;
    if (d == 0) continue;
    uint32_t bucket = zbase + (cfg.nwb == 1 ? 0u : (uint32_t)w * cfg.B) + d - 1;
    ZK_DCHECK(d - 1 < cfg.B && bucket < gridDim.y * cfg.nbuckets());
    if (SCATTER) {
      uint32_t pos = atomicAdd(&counts_or_cursor[bucket], 1u);
      uint32_t idx = (cfg.nwb == 1) ? (uint32_t)w * stride + first + i : first + i;
      ZK_DCHECK(pos < entries_cap && first + i < stride);
      __stcs(&entries[pos], idx | (neg << 31));
    } else {
      atomicAdd(&counts_or_cursor[bucket], 1u);
    }
  }
}

// Step 2: two-kernel exclusive scan (2048 elements per block): tile sums, then every block adds up
// the sums of the tiles before it (at most a few hundred values) and scans its own tile.
constexpr int SCAN_THREADS = 512, SCAN_ITEMS = 4, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* warp_sums) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t ws = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
    uint32_t winc = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_sums[lane] = winc - ws;  // exclusive
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  return warp_sums[wid] + inc - v;
}

static __global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tile_sums(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t ws[32];
  __shared__ uint32_t total;
  uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; j++) if (base + j < n) s += in[base + j];
  block_exclusive_scan(s, &total, ws);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// writes offsets[0..n] (exclusive scan, offsets[n] = grand total) and a copy in cursor[0..n);
// clears `in` (the bucket histogram) behind itself, so that the next MSM on the table finds it zeroed
static __global__ void __launch_bounds__(SCAN_THREADS)
k_scan_apply(uint32_t* __restrict__ in, uint32_t n, const uint32_t* __restrict__ tile_sums,
             uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor) {
  __shared__ uint32_t ws[32];
  __shared__ uint32_t total;
  uint32_t before = 0;
  for (uint32_t i = threadIdx.x; i < blockIdx.x; i += blockDim.x) before += tile_sums[i];
  block_exclusive_scan(before, &total, ws);
  const uint32_t tile_base = total;
  __syncthreads();
  uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; j++) {
    v[j] = (base + j < n) ? in[base + j] : 0;
    if (base + j < n) in[base + j] = 0;
    s += v[j];
  }
  uint32_t ex = block_exclusive_scan(s, &total, ws) + tile_base;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; j++) {
    if (base + j < n) { offsets[base + j] = ex; cursor[base + j] = ex; }
    ex += v[j];
    if (base + j == n - 1) offsets[n] = ex;
  }
}

// Step 4: bucket accumulation, exactly balanced.  The grid is persistent (one wave: resident
// blocks per SM x SM count) and thread t folds the contiguous slice [t * per, (t+1) * per) of the
// bucket-sorted entry list, per = ceil(E / T): every thread performs the same number of mixed adds
// whatever the scalar distribution.  A bucket that lies entirely inside one slice is written to
// bucket_sums directly; the (at most two) partial pieces at the ends of a slice go to
// partial[2t] (slice starts inside that bucket) / partial[2t+1] (bucket runs past the slice end) and
// k_fix_partials adds them up.  Bases are gathered with 128-bit loads, the next base is fetched
// while the current one is added.
// MINB = resident blocks per SM the register allocation is capped for; STAGED = the next base is
// fetched by cp.async (LDGSTS) into a per-thread shared-memory slot while the current one is being
// added: the gather latency (entry -> base, two dependent DRAM accesses) disappears behind the
// ~13 us of a mixed add without costing the 24-48 registers a register prefetch would; PAIRED =
// mixed add with its independent products issued as interleaved pairs (Fp only; wants the
// registers of MINB = 2).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int ACC_THREADS = 128;
template <class F>
constexpr size_t acc_stage_bytes() { return 2 * sizeof(Affine<F>) * ACC_THREADS; }   // double-buffered slot per thread

template <class F, int MINB, bool STAGED, bool PAIRED = false>
__global__ void __launch_bounds__(ACC_THREADS, MINB)
k_accumulate(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ entries,
             const uint32_t* __restrict__ offsets, XYZZ<F>* __restrict__ bucket_sums, XYZZ<F>* __restrict__ partial,
             uint32_t* __restrict__ open_bucket, uint32_t nbuckets, uint32_t npoints) {
  extern __shared__ uint4 acc_stage[];   // [2][VEC][ACC_THREADS] when STAGED: conflict-free 16-byte columns
  constexpr int VEC = sizeof(Affine<F>) / 16;
  const uint32_t T = gridDim.x * blockDim.x;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t E = offsets[nbuckets];
  const uint32_t per = (E + T - 1) / T;
  const uint64_t e0_64 = (uint64_t)t * per;
  if (per == 0 || e0_64 >= E) { open_bucket[t] = 0xffffffffu; return; }
  const uint32_t e0 = (uint32_t)e0_64;
  const uint32_t e1 = min(E, e0 + per);
  auto issue = [&](int buf, uint32_t ent) {
    ZK_DCHECK((ent & 0x7fffffffu) < npoints);
    const uint4* src = reinterpret_cast<const uint4*>(&bases[ent & 0x7fffffffu]);
#pragma unroll
    for (int j = 0; j < VEC; j++) cp_async16(&acc_stage[(buf * VEC + j) * ACC_THREADS + threadIdx.x], src + j);
    cp_async_commit();
  };
  uint32_t e = __ldcs(&entries[e0]);
  uint32_t e_next = 0;
  if (STAGED) {
    issue(0, e);
    if (e0 + 1 < e1) e_next = __ldcs(&entries[e0 + 1]);
  }
  // bucket of entry e0: largest b with offsets[b] <= e0
  uint32_t lo = 0, hi = nbuckets;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= e0) lo = mid; else hi = mid;
  }
  uint32_t b = lo;
  uint32_t b_end = offsets[b + 1];
  while (b_end <= e0) { b++; ZK_DCHECK(b < nbuckets); b_end = offsets[b + 1]; }   // skip empty buckets sharing the offset
  uint32_t seg_start = e0;
  XYZZ<F> acc = XYZZ<F>::inf();
  Affine<F> cur;
  for (uint32_t k = e0; k < e1; k++) {
    if (STAGED) {
      // base k + 1 starts travelling now; the index of base k + 2 is fetched one iteration early so
      // that issuing a copy never waits on its address
      const int buf = (int)((k - e0) & 1);
      const bool more = k + 1 < e1;
      if (more) issue(buf ^ 1, e_next);
      uint32_t e_next2 = 0;
      if (k + 2 < e1) e_next2 = __ldcs(&entries[k + 2]);
      if (k >= b_end) {  // bucket b is finished: it started at seg_start
        const bool complete = seg_start == offsets[b];
        store_vec_stream(complete ? &bucket_sums[b] : &partial[2 * (size_t)t], acc);   // incomplete => started before e0
        acc = XYZZ<F>::inf();
        do { b++; ZK_DCHECK(b < nbuckets); b_end = offsets[b + 1]; } while (b_end <= k);
        seg_start = k;
      }
      if (more) cp_async_wait<1>(); else cp_async_wait<0>();
      uint4* c4 = reinterpret_cast<uint4*>(&cur);
#pragma unroll
      for (int j = 0; j < VEC; j++) c4[j] = acc_stage[(buf * VEC + j) * ACC_THREADS + threadIdx.x];
      if (e >> 31) cur.y = cur.y.neg();
      if constexpr (PAIRED) acc.madd_paired(cur); else acc.madd(cur);
      e = e_next;
      e_next = e_next2;
    } else {
      if (k >= b_end) {
        const bool complete = seg_start == offsets[b];
        store_vec_stream(complete ? &bucket_sums[b] : &partial[2 * (size_t)t], acc);
        acc = XYZZ<F>::inf();
        do { b++; ZK_DCHECK(b < nbuckets); b_end = offsets[b + 1]; } while (b_end <= k);
        seg_start = k;
      }
      e = __ldcs(&entries[k]);
      ZK_DCHECK((e & 0x7fffffffu) < npoints);
      cur = load_vec(&bases[e & 0x7fffffffu]);
      if (e >> 31) cur.y = cur.y.neg();
      if constexpr (PAIRED) acc.madd_paired(cur); else acc.madd(cur);
    }
  }
  // last piece: bucket b from seg_start to e1
  ZK_DCHECK(b < nbuckets && 2 * (size_t)t + 1 < 2 * (size_t)T);
  const bool starts_here = seg_start == offsets[b];
  const bool ends_here = e1 == b_end;
  if (starts_here && ends_here) store_vec_stream(&bucket_sums[b], acc);
  else store_vec_stream(&partial[2 * (size_t)t + (e0 >= offsets[b] ? 0 : 1)], acc);
  // the bucket this slice stops inside of (it continues in slice t + 1), for k_fix_partials
  open_bucket[t] = ends_here ? 0xffffffffu : b;
}

// Step 4b: buckets split over several slices.  One thread per accumulation SLICE (not per bucket: the
// split buckets are at most T of the Q * nb buckets, and a thread per bucket would run the addition
// code with one or two live lanes per warp): slice t stopped inside bucket open_bucket[t]; the
// thread of the slice the bucket STARTS in adds up its pieces (slot rule as in k_accumulate) — or
// queues it for k_fix_heavy when it spans more than HEAVY_PIECES slices (skewed scalars, SURVEY.md
// H4).  Empty buckets are never written: k_reduce_chunks reads them as the identity.
constexpr uint32_t HEAVY_PIECES = 8;
template <class F>
__global__ void __launch_bounds__(128)
k_fix_partials(const uint32_t* __restrict__ offsets, XYZZ<F>* __restrict__ bucket_sums,
               const XYZZ<F>* __restrict__ partial, const uint32_t* __restrict__ open_bucket, uint32_t nbuckets, uint32_t T,
               uint32_t* __restrict__ heavy) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint32_t b = open_bucket[t];
  if (b == 0xffffffffu) return;
  ZK_DCHECK(b < nbuckets);
  const uint32_t E = offsets[nbuckets];
  const uint32_t per = (E + T - 1) / T;
  const uint32_t lo = offsets[b], hi = offsets[b + 1];
  const uint32_t t_first = lo / per, t_last = (hi - 1) / per;
  ZK_DCHECK(lo < hi && hi <= E && t_last < T && t_first <= t && t < t_last);
  if (t_first != t) return;       // the bucket started in an earlier slice: that slice's thread owns it
  if (t_last - t_first + 1 > HEAVY_PIECES) {
    const uint32_t slot = atomicAdd(heavy, 1u);
    ZK_DCHECK(slot + 1 < T / 4 + 2);   // heavy queue capacity: a heavy bucket spans > 8 of the T slices
    heavy[1 + slot] = b;
    return;
  }
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t s = t_first; s <= t_last; s++) {
    const uint32_t slot = ((uint64_t)s * per >= lo) ? 0 : 1;
    XYZZ<F> p = load_vec_rw(&partial[2 * (size_t)s + slot]);
    acc.add(p);
  }
  store_vec(&bucket_sums[b], acc);
}

// Step 4c: heavy buckets, one BLOCK each (grid-stride over the queue): threads stride over the
// pieces, a shared-memory tree adds the per-thread sums: pieces / blockDim + log2(blockDim)
// sequential additions (a 2^20-point MSM whose scalars are half ones has a 17 000-piece bucket).
template <class F>
__global__ void __launch_bounds__(256)
k_fix_heavy(const uint32_t* __restrict__ offsets, XYZZ<F>* __restrict__ bucket_sums,
            const XYZZ<F>* __restrict__ partial, uint32_t nbuckets, uint32_t T, const uint32_t* __restrict__ heavy) {
  extern __shared__ uint4 smem_raw[];
  XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(smem_raw);
  const uint32_t count = heavy[0];
  if (count == 0) return;   // block-uniform
  const uint32_t E = offsets[nbuckets];
  const uint32_t per = (E + T - 1) / T;
  for (uint32_t q = blockIdx.x; q < count; q += gridDim.x) {   // block-uniform
    const uint32_t b = heavy[1 + q];
    const uint32_t lo = offsets[b], hi = offsets[b + 1];
    const uint32_t t_first = lo / per, t_last = (hi - 1) / per;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t t = t_first + threadIdx.x; t <= t_last; t += blockDim.x) {
      const uint32_t slot = ((uint64_t)t * per >= lo) ? 0 : 1;
      XYZZ<F> p = load_vec_rw(&partial[2 * (size_t)t + slot]);
      acc.add(p);
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t stride = blockDim.x / 2; stride > 0; stride >>= 1) {
      if (threadIdx.x < stride) {
        XYZZ<F> a = sm[threadIdx.x];
        a.add(sm[threadIdx.x + stride]);
        sm[threadIdx.x] = a;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) store_vec(&bucket_sums[b], sm[0]);
    __syncthreads();
  }
}

// k * p for a small (< 2^31) multiplier
template <class F>
__device__ __noinline__ XYZZ<F> small_mul(const XYZZ<F>& p, uint32_t k) {
  XYZZ<F> acc = XYZZ<F>::inf();
  if (k == 0 || p.is_inf()) return acc;
  int top = 31 - __clz(k);
  for (int bit = top; bit >= 0; bit--) {
    acc = acc.dbl();
    if ((k >> bit) & 1) acc.add(p);
  }
  return acc;
}

// Step 5a: thread (window wb, chunk ch) reduces L consecutive buckets with the running-sum
// trick:  V = sum_{j<L} (ch*L + j + 1) * bucket[ch*L + j].
// (tail kernels use 64-thread blocks so that they fit in the registers a pipelined accumulation
// leaves free on every SM)
constexpr int TAIL_THREADS = 64;
template <class F>
__global__ void __launch_bounds__(TAIL_THREADS)
k_reduce_chunks(const XYZZ<F>* __restrict__ bucket_sums_all, const uint32_t* __restrict__ offsets_all, MsmConfig cfg,
                XYZZ<F>* __restrict__ chunk_out_all) {
  uint32_t chunks_per_window = cfg.B / cfg.L;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= chunks_per_window * (uint32_t)cfg.nwb) return;
  const XYZZ<F>* bucket_sums = bucket_sums_all + (size_t)blockIdx.z * cfg.nbuckets();       // z = queued MSM
  const uint32_t* offsets = offsets_all + (size_t)blockIdx.z * cfg.nbuckets();
  XYZZ<F>* chunk_out = chunk_out_all + (size_t)blockIdx.z * chunks_per_window * cfg.nwb;
  uint32_t wb = t / chunks_per_window, ch = t % chunks_per_window;
  uint32_t first = wb * cfg.B + ch * cfg.L;
  XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
  uint32_t hi = offsets[first + cfg.L];
  for (int j = cfg.L - 1; j >= 0; j--) {
    const uint32_t lo = offsets[first + j];
    if (lo != hi) {                       // an empty bucket was never written: it is the identity
      XYZZ<F> p = load_vec_stream(&bucket_sums[first + j]);
      run.add(p);
    }
    hi = lo;
    acc.add(run);
  }
  // acc = sum (j+1) * bucket_j ; run = sum bucket_j ; add (ch*L) * run
  XYZZ<F> shifted = small_mul(run, ch * (uint32_t)cfg.L);
  acc.add(shifted);
  store_vec(&chunk_out[t], acc);
}

// Step 5b: one level of the sum tree: block (x, window y, queued MSM z) adds up to TAIL_THREADS
// consecutive elements of its window (count per window = n_in) and writes one:
// out[z * out_stride + y * gridDim.x + x].
template <class F>
__global__ void __launch_bounds__(TAIL_THREADS)
k_reduce_tree(const XYZZ<F>* __restrict__ in, uint32_t n_in, size_t in_stride, XYZZ<F>* __restrict__ out,
              size_t out_stride) {
  extern __shared__ uint4 smem_raw[];
  XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(smem_raw);
  const XYZZ<F>* src = in + (size_t)blockIdx.z * in_stride + (size_t)blockIdx.y * n_in;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  XYZZ<F> acc = XYZZ<F>::inf();
  if (i < n_in) acc = load_vec_rw(&src[i]);
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t stride = blockDim.x / 2; stride > 0; stride >>= 1) {
    if (threadIdx.x < stride && i + stride < n_in) {
      XYZZ<F> a = sm[threadIdx.x];
      a.add(sm[threadIdx.x + stride]);
      sm[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) store_vec(&out[(size_t)blockIdx.z * out_stride + (size_t)blockIdx.y * gridDim.x + blockIdx.x], sm[0]);
}

// Output slots of the queued MSMs (passed by value to the batched tail kernels).
template <class F>
struct TailOutputs {
  XYZZ<F>* result[MSM_QUEUE];   // XYZZ sum (always set: caller's slot or table scratch)
  uint8_t* bytes[MSM_QUEUE];    // RAW + COMP wire bytes, nullable
};

// Step 6 + 7 for every queued MSM (block z): result = sum_w 2^(c*w) * window_sums[w] (Horner; a copy
// when the table is precomputed), then XYZZ -> affine -> wire bytes.
template <class T>
__global__ void k_combine_finalize(const XYZZ<typename T::F>* __restrict__ window_sums_all, MsmConfig cfg,
                                   TailOutputs<typename T::F> outs) {
  typedef typename T::F F;
  if (threadIdx.x != 0) return;
  const int z = blockIdx.x;
  const XYZZ<F>* window_sums = window_sums_all + (size_t)z * (cfg.nwb + 1);
  XYZZ<F> acc = load_vec_rw(&window_sums[cfg.nwb - 1]);
  for (int w = cfg.nwb - 2; w >= 0; w--) {
    for (int i = 0; i < cfg.c; i++) acc = acc.dbl();
    XYZZ<F> p = load_vec_rw(&window_sums[w]);
    acc.add(p);
  }
  store_vec(outs.result[z], acc);
  if (outs.bytes[z]) {
    Affine<F> a = acc.to_affine();
    T::serialize(a, outs.bytes[z]);
  }
}

// `count` XYZZ results -> RAW + COMP bytes each
template <class T>
__global__ void k_finalize(const XYZZ<typename T::F>* __restrict__ results, int count, uint8_t* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  XYZZ<typename T::F> p = load_vec_rw(&results[i]);
  Affine<typename T::F> a = p.to_affine();
  T::serialize(a, out + (size_t)i * (T::RAW + T::COMP));
}

// Combine of the devices' partial sums: out[q] = sum_p gather[q * parts + p], as wire
// bytes at out + q * out_stride; one block per q.  The parts stored their XYZZ partials into
// `gather` on the primary device as peer-to-peer stores.
template <class T>
__global__ void k_sum_parts(const XYZZ<typename T::F>* __restrict__ gather, uint32_t parts, uint8_t* __restrict__ out,
                            uint32_t out_stride) {
  if (threadIdx.x) return;
  const uint32_t q = blockIdx.x;
  XYZZ<typename T::F> acc = load_vec_rw(&gather[(size_t)q * parts]);
  for (uint32_t p = 1; p < parts; p++) {
    XYZZ<typename T::F> x = load_vec_rw(&gather[(size_t)q * parts + p]);
    acc.add(x);
  }
  Affine<typename T::F> a = acc.to_affine();
  T::serialize(a, out + (size_t)q * out_stride);
}

// ---- table construction ---------------------------------------------------------
// raw wire bytes -> Montgomery affine, curve check, identity flags
template <class T>
__global__ void __launch_bounds__(128)
k_parse_bases(const uint8_t* __restrict__ raw, const uint8_t* __restrict__ inf_flags, uint32_t n,
              Affine<typename T::F>* __restrict__ out, uint8_t* __restrict__ skip, int* __restrict__ err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<typename T::F> p;
  int bad = 0;
  if (inf_flags && inf_flags[i]) p = Affine<typename T::F>::inf();
  else bad = T::parse(raw + (size_t)i * T::RAW, p);
  if (bad) { atomicExch(err, 1); p = Affine<typename T::F>::inf(); }
  store_vec(&out[i], p);
  skip[i] = p.is_inf() ? 1 : 0;
}

// window w of the precomputed table from window w-1:  P -> 2^c * P   (XYZZ scratch)
template <class F>
__global__ void __launch_bounds__(128)
k_precompute_shift(const Affine<F>* __restrict__ prev, uint32_t n, int c, XYZZ<F>* __restrict__ scratch) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = load_vec(&prev[i]);
  XYZZ<F> acc = XYZZ<F>::dbl_affine(p);
  for (int k = 1; k < c; k++) acc = acc.dbl();
  store_vec(&scratch[i], acc);
}

// XYZZ -> affine with Montgomery's simultaneous inversion over BATCH points per thread
template <class F, int BATCH>
__global__ void __launch_bounds__(128)
k_batch_to_affine(const XYZZ<F>* __restrict__ src, uint32_t n, Affine<F>* __restrict__ dst) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t first = t * BATCH;
  if (first >= n) return;
  uint32_t cnt = min((uint32_t)BATCH, n - first);
  // prefix products of ZZZ (identity points contribute 1); prefixes parked in dst[].x
  F run = F::one();
  for (uint32_t j = 0; j < cnt; j++) {
    F zzz = load_vec_rw(&src[first + j].ZZZ);
    store_vec(&dst[first + j].x, run);
    if (!zzz.is_zero()) run = run * zzz;
  }
  F inv = run.inverse();
  for (int j = (int)cnt - 1; j >= 0; j--) {
    XYZZ<F> p = load_vec_rw(&src[first + j]);
    Affine<F> a;
    if (p.is_inf()) {
      a = Affine<F>::inf();
    } else {
      F prefix = load_vec_rw(&dst[first + j].x);
      F izzz = inv * prefix;   // 1 / ZZZ_j
      inv = inv * p.ZZZ;       // drop ZZZ_j from the running inverse
      F iz = izzz * p.ZZ;
      F izz = iz.sqr();
      a.x = p.X * izz;
      a.y = p.Y * izzz;
    }
    store_vec(&dst[first + j], a);
  }
}

#endif  // __CUDACC__

// ------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------
template <class T>
struct BaseTable {
  typedef typename T::F F;
  uint32_t n = 0;
  bool precomputed = false;
  MsmConfig cfg{};
  DevBuf<Affine<F>> pts;     // n points, or W * n when precomputed (window-major)
  DevBuf<uint8_t> skip;      // 1 = identity base
  // workspace (reused by every join on this table; calls on one table are stream-ordered).  Sized
  // for queue_cap queued MSMs: queued MSM q owns buckets [q * nb, (q + 1) * nb).
  DevBuf<uint32_t> counts, cursor, tile_sums, entries;   // counts is all zero between joins (k_scan_apply clears it)
  DevBuf<uint32_t> offsets, heavy;   // bucket offsets (queue_cap * nb + 1); heavy[0] = queue length, then the queue
  DevBuf<uint32_t> open_bucket;      // per accumulation slice: the bucket it stopped inside of (or ~0)
  // Queued MSMs.  run() only records the MSM (scalar pointer, range, output slots); join() sorts,
  // accumulates and reduces everything queued with ONE launch sequence on the caller's stream: the
  // latency-bound end of an MSM (bucket reduction, window combine, affine conversion: ~40 dependent
  // point operations) costs the same wall time for one MSM as for a batch, and so do the partial
  // pieces and the launch gaps of the sort and the accumulation.  A table that is not pipelined
  // joins inside every run().  The scalars of a queued MSM must stay valid until its join has run.
  DevBuf<XYZZ<F>> bucket_sums, partial, chunk_out, tree_tmp, window_sums;
  uint32_t acc_blocks = 0;   // persistent grid of k_accumulate: resident blocks per SM x SMs
  int acc_variant = 0;       // ZKB200_ACC_VARIANT[_G2]: see build_tables (9 / 5 = cp.async-staged defaults for G1 / G2)
  template <class Fn> void acc_dispatch(Fn&& fn);
  int acc_occupancy();
  void acc_launch(uint32_t grid, uint32_t nbuckets, cudaStream_t st);
  int queued = 0;
  int pending = 0;           // MSMs sorted and accumulated whose tail has not been enqueued yet
  uint32_t pending_threads = 0;   // threads their accumulation ran with
  int queue_cap = 0;         // MSMs one join can take with the buffers currently allocated (1 unless pipelined)
  int queue_limit = 1;       // MSMs the caller asked to queue at most (<= queue_cap): run() joins when it is reached
  QueueSlots slots{};
  TailOutputs<F> outs{};
  bool pipelined = false;    // false: every run() joins immediately (plain stream order)

  static MsmConfig choose_config(uint32_t n, bool precompute, int force_c);
  void load(const uint8_t* host_raw, const uint8_t* host_inf, uint32_t n, bool precompute, int force_c,
            cudaStream_t st);
  void build_tables(cudaStream_t st);
  // d_scalars: count * 32 B canonical little-endian; uses bases [first, first + count).
  // d_result (nullable) receives the XYZZ sum, d_out_bytes (nullable) the RAW + COMP bytes.
  // d_err (nullable): set to 1 when a scalar is not canonical (>= r).
  void run(const uint32_t* d_scalars, uint32_t count, XYZZ<F>* d_result, uint8_t* d_out_bytes, cudaStream_t st,
           uint32_t first = 0, int* d_err = nullptr);
  // run every queued MSM on `st`; results are valid in stream order afterwards.  after_scatter
  // (nullable) is recorded once the scalars have been read for the last time (uploads into the same
  // staging buffers may then proceed).
  void join(cudaStream_t st, cudaEvent_t after_scatter = nullptr);
  // The two halves of join(), for callers that overlap the latency-bound tails of two tables on two
  // streams (Groth16: G1 and G2): sort_accumulate() runs the throughput-bound part of everything
  // queued, tail() the batched reduction of what sort_accumulate() left pending (any stream that is
  // ordered after it).
  void sort_accumulate(cudaStream_t st, cudaEvent_t after_scatter = nullptr);
  void tail(cudaStream_t st, int share = 1);   // share = tails of this many tables run side by side
  // pipelined = queue up to `depth` MSMs per join (allocates the buffers for that many on first use);
  // depth 0 = the default depth (ZKB200_QUEUE, 32), at most MSM_QUEUE
  void set_pipelined(bool on, int depth = 0);
  void ensure_queue(int slots);
  // forget every queued tail (after an error between run() and join(): the queued output pointers
  // may refer to buffers that are being unwound); the caller drains the streams first
  void abort_queue() {
    queued = 0;
    pending = 0;
    if (counts.p) cudaMemset(counts.p, 0, counts.bytes());   // a run cut short may have left the histogram dirty
  }
  // stage timing (bench.py's roofline leg): when `profile` is set, every join brackets its stages
  // with CUDA events (a ring of PROF_RING joins); stage_ms() / stage_totals() read them after the
  // streams have drained.
  // stages: 0 digits+scan+scatter, 1 accumulate, 2 partial fix-up + bucket reduce, 3 combine+finalize
  static constexpr int PROF_RING = 64;
  struct ProfRec { cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; int msms = 0; };
  bool profile = false;
  std::vector<ProfRec> prof;   // allocated when profiling is first switched on
  uint64_t prof_joins = 0;     // joins recorded since profiling was switched on
  void set_profile(bool on);
  void stage_ms(float out[4]);                                        // the last profiled join
  void stage_totals(float out[4], uint64_t* msms, uint64_t* joins);   // sums over the recorded joins
  ~BaseTable();
  size_t device_bytes() const;
};

// Pipelines a table for the lifetime of the scope.  If the scope is left by an exception while
// tails are still queued, the table's device is drained and the queue dropped, so that a later
// run() / join() on the handle never writes through output pointers of the failed call.
int current_ctx();
void set_ctx(int ctx);
template <class T>
struct PipelineScope {
  BaseTable<T>& t;
  bool was;
  int exc, ctx;
  cudaStream_t extra;
  // depth = tails the scope will queue before it joins (a proof knows: 2 for Groth16's A and C)
  PipelineScope(BaseTable<T>& t_, int ctx_, cudaStream_t extra_ = nullptr, int depth = 0)
      : t(t_), was(t_.pipelined), exc(std::uncaught_exceptions()), ctx(ctx_), extra(extra_) { t.set_pipelined(true, depth); }
  PipelineScope(const PipelineScope&) = delete;
  PipelineScope& operator=(const PipelineScope&) = delete;
  ~PipelineScope() {
    if (std::uncaught_exceptions() > exc) {
      const int keep = current_ctx();
      try { set_ctx(ctx); } catch (...) {}
      cudaDeviceSynchronize();
      if (extra) cudaStreamSynchronize(extra);
      t.abort_queue();
      try { set_ctx(keep); } catch (...) {}
    }
    t.pipelined = was;
  }
};

template <class T>
void finalize_points(const XYZZ<typename T::F>* d_results, int count, uint8_t* d_out, cudaStream_t st);

}  // namespace zk
