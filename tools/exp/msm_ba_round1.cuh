// Batched-affine bucket accumulation (alternative to k_accumulate's XYZZ mixed adds).
//
// An affine addition costs 1 inversion + 2M + 1S; sharing the inversion over a batch with
// Montgomery's trick brings an addition to ~6 field products (against 10 for an XYZZ mixed add),
// provided the batch is made of INDEPENDENT additions.  Inside a bucket the additions of a running
// sum depend on each other, so the bucket sums are formed by rounds of pairwise additions instead:
// in round r every bucket with m points adds them in pairs (point 2i with point 2i+1) and keeps
// ceil(m/2) points.  All pairs of a round, over all buckets, are independent; thread t takes BA_K
// consecutive output points, multiplies up its BA_K denominators, inverts once with the binary
// Euclid routine (additions and shifts only: it runs on the integer-add pipe while other warps keep
// the IMAD pipe busy) and unwinds.  Prefix products are parked in a scratch array laid out
// [j][thread] so the accesses coalesce.
//
// After the last round every bucket holds at most a couple of points; k_ba_finish turns them into
// the XYZZ bucket sums the common tail (bucket reduction, window combine) consumes.
#pragma once
#include "msm.cuh"

namespace zk {

constexpr int BA_K = 16;

// m -> ceil(m / 2) per bucket (input to the scan that yields the next round's offsets)
static __global__ void k_ba_next_counts(const uint32_t* __restrict__ off, uint32_t nb, uint32_t* __restrict__ counts) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  uint32_t m = off[b + 1] - off[b];
  counts[b] = (m + 1) >> 1;
}

template <class F>
struct BaPair {
  Affine<F> A, B;
  bool has_b;
};

// loads the pair feeding output index o of bucket b (i = o - out_off[b])
template <class F, bool FIRST>
__device__ __forceinline__ void ba_load(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ entries,
                                        const Affine<F>* __restrict__ in_pts, uint32_t in_lo, uint32_t m, uint32_t i,
                                        BaPair<F>& p) {
  const uint32_t ia = in_lo + 2 * i;
  p.has_b = 2 * i + 1 < m;
  if (FIRST) {
    uint32_t ea = entries[ia];
    p.A = load_vec(&bases[ea & 0x7fffffffu]);
    if (ea >> 31) p.A.y = p.A.y.neg();
    if (p.has_b) {
      uint32_t eb = entries[ia + 1];
      p.B = load_vec(&bases[eb & 0x7fffffffu]);
      if (eb >> 31) p.B.y = p.B.y.neg();
    }
  } else {
    p.A = load_vec_rw(&in_pts[ia]);
    if (p.has_b) p.B = load_vec_rw(&in_pts[ia + 1]);
  }
}

// denominator of the pair's addition: x2 - x1, or 2 y for a doubling, or 1 where no inverse is needed
template <class F>
__device__ __forceinline__ F ba_denominator(const BaPair<F>& p) {
  if (!p.has_b || p.A.is_inf() || p.B.is_inf()) return F::one();
  F d = p.B.x - p.A.x;
  if (!d.is_zero()) return d;
  if (p.A.y == p.B.y && !p.A.y.is_zero()) return p.A.y.dbl();
  return F::one();  // P + (-P)
}

template <class F>
__device__ __forceinline__ Affine<F> ba_add(const BaPair<F>& p, const F& dinv) {
  if (!p.has_b || p.B.is_inf()) return p.A;
  if (p.A.is_inf()) return p.B;
  F lam;
  if (p.A.x == p.B.x) {
    if (!(p.A.y == p.B.y) || p.A.y.is_zero()) return Affine<F>::inf();
    F xx = p.A.x.sqr();
    lam = (xx.dbl() + xx) * dinv;  // 3 x^2 / (2 y)
  } else {
    lam = (p.B.y - p.A.y) * dinv;
  }
  Affine<F> r;
  r.x = lam.sqr() - p.A.x - p.B.x;
  r.y = lam * (p.A.x - r.x) - p.A.y;
  return r;
}

// One round.  in_off / out_off: bucket offsets of the input / output point lists (nb + 1 entries).
template <class F, bool FIRST>
__global__ void __launch_bounds__(128, (sizeof(F) > 48 ? 2 : 3))
k_ba_round(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ entries,
           const Affine<F>* __restrict__ in_pts, const uint32_t* __restrict__ in_off,
           const uint32_t* __restrict__ out_off, uint32_t nb, Affine<F>* __restrict__ out_pts,
           F* __restrict__ scratch, uint32_t T) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = out_off[nb];
  const uint64_t o0_64 = (uint64_t)t * BA_K;
  if (t >= T || o0_64 >= total) return;
  const uint32_t o0 = (uint32_t)o0_64;
  const uint32_t cnt = min((uint32_t)BA_K, total - o0);
  // bucket of output o0
  uint32_t lo = 0, hi = nb;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (out_off[mid] <= o0) lo = mid; else hi = mid;
  }
  const uint32_t b0 = lo;
  // ---- pass 1: running product of the denominators ---------------------------------------
  F run = F::one();
  {
    uint32_t b = b0, b_end = out_off[b + 1];
    for (uint32_t j = 0; j < cnt; j++) {
      const uint32_t o = o0 + j;
      while (o >= b_end) { b++; b_end = out_off[b + 1]; }
      const uint32_t in_lo = in_off[b], m = in_off[b + 1] - in_lo;
      BaPair<F> p;
      ba_load<F, FIRST>(bases, entries, in_pts, in_lo, m, o - out_off[b], p);
      store_vec(&scratch[(size_t)j * T + t], run);
      run = run * ba_denominator(p);
    }
  }
  F inv = run.inverse();
  // ---- pass 2: unwind, form the sums -----------------------------------------------------------
  {
    // bucket of the last output, then walk downwards
    uint32_t b = b0;
    {
      const uint32_t o_last = o0 + cnt - 1;
      uint32_t b_end = out_off[b + 1];
      while (o_last >= b_end) { b++; b_end = out_off[b + 1]; }
    }
    for (int j = (int)cnt - 1; j >= 0; j--) {
      const uint32_t o = o0 + (uint32_t)j;
      while (o < out_off[b]) b--;
      const uint32_t in_lo = in_off[b], m = in_off[b + 1] - in_lo;
      BaPair<F> p;
      ba_load<F, FIRST>(bases, entries, in_pts, in_lo, m, o - out_off[b], p);
      F d = ba_denominator(p);
      F dinv = inv * load_vec_rw(&scratch[(size_t)j * T + t]);
      inv = inv * d;
      store_vec(&out_pts[o], ba_add(p, dinv));
    }
  }
}

// After the rounds: bucket b holds m = off[b+1] - off[b] (normally <= 1) affine points.
template <class F, bool FIRST>
__global__ void __launch_bounds__(128)
k_ba_finish(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ entries,
            const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ off, uint32_t nb,
            XYZZ<F>* __restrict__ bucket_sums) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t k = off[b]; k < off[b + 1]; k++) {
    Affine<F> p;
    if (FIRST) {
      uint32_t e = entries[k];
      p = load_vec(&bases[e & 0x7fffffffu]);
      if (e >> 31) p.y = p.y.neg();
    } else {
      p = load_vec_rw(&pts[k]);
    }
    acc.madd(p);
  }
  store_vec(&bucket_sums[b], acc);
}

}  // namespace zk
