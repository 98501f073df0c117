// Fr polynomial kernels and the QAP quotient (see fr_poly.cuh for the algorithm note).
#include "fr_poly.cuh"
#include "runtime.cuh"

namespace zk {

// ---------------------------------------------------------------------------
// small vector kernels
// ---------------------------------------------------------------------------
static __global__ void k_fr_to_mont(const uint32_t* __restrict__ raw, Fr* __restrict__ out, uint32_t n, int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr a = load_vec(reinterpret_cast<const Fr*>(raw) + i);
  if (!a.is_canonical_raw()) { if (err) atomicExch(err, 1); a = Fr::zero(); }
  store_vec(&out[i], a.to_mont());
}
static __global__ void k_fr_from_mont(const Fr* __restrict__ in, uint32_t* __restrict__ raw, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr a = load_vec_rw(&in[i]);
  store_vec(reinterpret_cast<Fr*>(raw) + i, a.from_mont());
}
void fr_to_mont(const uint32_t* d_raw, Fr* d_out, uint32_t n, int* d_err, cudaStream_t st) {
  if (n) k_fr_to_mont<<<cdiv(n, 256), 256, 0, st>>>(d_raw, d_out, n, d_err);
}
void fr_from_mont(const Fr* d_in, uint32_t* d_raw, uint32_t n, cudaStream_t st) {
  if (n) k_fr_from_mont<<<cdiv(n, 256), 256, 0, st>>>(d_in, d_raw, n);
}
__device__ __noinline__ Fr fr_pow_u32(const Fr& base, uint32_t e) {
  Fr acc = Fr::one();
  if (e == 0) return acc;
  int top = 31 - __clz(e);
  for (int bit = top; bit >= 0; bit--) {
    acc = acc * acc;
    if ((e >> bit) & 1) acc = acc * base;
  }
  return acc;
}

// consts[0] = omega_D, [1] = omega_D^-1, [2] = g, [3] = g^-1, [4] = 1/D   (Montgomery)
static __global__ void k_ntt_consts(int logD, Fr* consts) {
  if (threadIdx.x || blockIdx.x) return;
  Fr w, wi, g, gi, i2;
  for (int i = 0; i < 8; i++) {
    w.v[i] = FrParams::root32(i); wi.v[i] = FrParams::root32_inv(i);
    g.v[i] = FrParams::coset_g(i); gi.v[i] = FrParams::coset_g_inv(i); i2.v[i] = FrParams::inv2(i);
  }
  for (int i = logD; i < 32; i++) { w = w * w; wi = wi * wi; }
  Fr dinv = Fr::one();
  for (int i = 0; i < logD; i++) dinv = dinv * i2;
  consts[0] = w; consts[1] = wi; consts[2] = g; consts[3] = gi; consts[4] = dinv;
}

// out[i] = scale * base^i, 64 consecutive powers per thread
static __global__ void k_fr_powers(const Fr* base, const Fr* scale, Fr* __restrict__ out, uint32_t n) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t i0 = t * 64;
  if (i0 >= n) return;
  Fr b = load_vec_rw(base);
  Fr p = fr_pow_u32(b, i0);
  if (scale) p = p * load_vec_rw(scale);
  uint32_t end = min(n, i0 + 64);
  for (uint32_t i = i0; i < end; i++) {
    store_vec(&out[i], p);
    p = p * b;
  }
}

// ---------------------------------------------------------------------------
// radix-2 NTT stages (gridDim.y = number of vectors, each D long and contiguous)
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
k_ntt_dif_stage(Fr* __restrict__ x, const Fr* __restrict__ tw, uint32_t D, int logHalf, int s) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= D / 2) return;
  Fr* v = x + (size_t)blockIdx.y * D;
  uint32_t half = 1u << logHalf;
  uint32_t j = t & (half - 1);
  uint32_t i0 = ((t >> logHalf) << (logHalf + 1)) + j;
  uint32_t i1 = i0 + half;
  Fr a = load_vec_rw(&v[i0]), b = load_vec_rw(&v[i1]);
  Fr w = load_vec(&tw[(size_t)j << s]);
  store_vec(&v[i0], a + b);
  store_vec(&v[i1], (a - b) * w);
}
static __global__ void __launch_bounds__(256)
k_ntt_dit_stage(Fr* __restrict__ x, const Fr* __restrict__ tw_inv, uint32_t D, int logHalf, int s) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= D / 2) return;
  Fr* v = x + (size_t)blockIdx.y * D;
  uint32_t half = 1u << logHalf;
  uint32_t j = t & (half - 1);
  uint32_t i0 = ((t >> logHalf) << (logHalf + 1)) + j;
  uint32_t i1 = i0 + half;
  Fr a = load_vec_rw(&v[i0]);
  Fr b = load_vec_rw(&v[i1]) * load_vec(&tw_inv[(size_t)j << s]);
  store_vec(&v[i0], a + b);
  store_vec(&v[i1], a - b);
}

void NttPlan::build(int logD_, cudaStream_t st) {
  ZK_REQUIRE(logD_ >= 1 && logD_ <= 28, ZK_EARG, "NTT size out of range");
  logD = logD_;
  D = 1u << logD;
  DevBuf<Fr> consts(5);
  k_ntt_consts<<<1, 1, 0, st>>>(logD, consts.p);
  tw.alloc(D / 2);
  tw_inv.alloc(D / 2);
  coset.alloc(D);
  coset_inv.alloc(D);
  k_fr_powers<<<cdiv(cdiv(D / 2, 64), 128), 128, 0, st>>>(consts.p + 0, nullptr, tw.p, D / 2);
  k_fr_powers<<<cdiv(cdiv(D / 2, 64), 128), 128, 0, st>>>(consts.p + 1, nullptr, tw_inv.p, D / 2);
  k_fr_powers<<<cdiv(cdiv(D, 64), 128), 128, 0, st>>>(consts.p + 2, nullptr, coset.p, D);
  k_fr_powers<<<cdiv(cdiv(D, 64), 128), 128, 0, st>>>(consts.p + 3, consts.p + 4, coset_inv.p, D);
  ZK_CUDA(cudaGetLastError());
  ZK_CUDA(cudaStreamSynchronize(st));  // consts is freed on return
}

// ---------------------------------------------------------------------------
// fused passes: K consecutive radix-2 stages per launch, on a tile staged in shared memory
// ---------------------------------------------------------------------------
// A radix-2 stage moves 64 B per butterfly through HBM for one Fr product: at 2^21 points a stage is
// HBM-bound (and a launch each).  A pass keeps a tile of 2^TL elements in shared memory and runs K
// stages on it, so the data crosses HBM once per pass (3 passes instead of 21 stages at 2^21); what
// is left is the integer pipe (one Fr product per butterfly).
//
// Index bits of a pass over stages [s0, s0 + K): i = hi (s0 bits) | m (K bits) | low (L bits), stage
// s0 + t pairs m with m ^ 2^(K-1-t).  A tile holds every m for C = 2^(TL-K) consecutive values of
// `low` (C * 32 contiguous bytes per m), element (m, c) at slot m * C + c.  The two 16-byte halves of
// an element live in separate planes, so consecutive threads touch consecutive 16-byte words.
constexpr int NTT_TL = 10;                   // tile = 1024 elements = 32 KB
constexpr int NTT_THREADS = 256;
struct NttPass {
  int s0, K, logC;                           // first stage, stages, log2(C)
};

template <bool INVERSE>
__global__ void __launch_bounds__(NTT_THREADS)
k_ntt_pass(Fr* __restrict__ x, const Fr* __restrict__ tw, uint32_t D, int logD, NttPass P) {
  extern __shared__ uint4 ntt_tile[];        // [2][tile]
  const int K = P.K, logC = P.logC;
  const uint32_t tile = 1u << (K + logC), C = 1u << logC;
  const int L = logD - P.s0 - K;             // low bits below the K stage bits
  Fr* v = x + (size_t)blockIdx.y * D;
  // tile id -> (hi, low base): tiles per hi block = 2^L / C
  const uint32_t tiles_per_hi = 1u << (L - logC);
  const uint32_t hi = blockIdx.x / tiles_per_hi, lowbase = (blockIdx.x % tiles_per_hi) << logC;
  const size_t base = ((size_t)hi << (logD - P.s0)) + lowbase;
  auto gidx = [&](uint32_t e) {
    const size_t g = base + ((size_t)(e >> logC) << L) + (e & (C - 1));
    ZK_DCHECK(g < D);
    return g;
  };
  for (uint32_t e = threadIdx.x; e < tile; e += NTT_THREADS) {
    const uint4* src = reinterpret_cast<const uint4*>(&v[gidx(e)]);
    ntt_tile[e] = src[0];
    ntt_tile[tile + e] = src[1];
  }
  __syncthreads();
  for (int tt = 0; tt < K; tt++) {
    const int t = INVERSE ? K - 1 - tt : tt; // inverse: smallest distance first
    const int s = P.s0 + t;                  // global stage: pairs at distance D >> (s + 1)
    const int pb = K - 1 - t;                // bit of m that differs inside a pair
    for (uint32_t q = threadIdx.x; q < tile / 2; q += NTT_THREADS) {
      const uint32_t c = q & (C - 1), mm = q >> logC;
      const uint32_t m0 = ((mm >> pb) << (pb + 1)) | (mm & ((1u << pb) - 1));
      const uint32_t e0 = (m0 << logC) | c, e1 = e0 + (1u << (pb + logC));
      // twiddle exponent: (i0 mod half) << s, i0 mod half = (m0 mod 2^pb) << L | low
      const uint32_t j = ((m0 & ((1u << pb) - 1)) << L) | (lowbase + c);
      Fr a, b;
      uint4* a4 = reinterpret_cast<uint4*>(&a);
      uint4* b4 = reinterpret_cast<uint4*>(&b);
      a4[0] = ntt_tile[e0]; a4[1] = ntt_tile[tile + e0];
      b4[0] = ntt_tile[e1]; b4[1] = ntt_tile[tile + e1];
      ZK_DCHECK(((size_t)j << s) < D / 2 && e1 < tile);
      const Fr w = load_vec(&tw[(size_t)j << s]);
      Fr r0, r1;
      if (INVERSE) { b = b * w; r0 = a + b; r1 = a - b; }
      else { r0 = a + b; r1 = (a - b) * w; }
      const uint4* r04 = reinterpret_cast<const uint4*>(&r0);
      const uint4* r14 = reinterpret_cast<const uint4*>(&r1);
      ntt_tile[e0] = r04[0]; ntt_tile[tile + e0] = r04[1];
      ntt_tile[e1] = r14[0]; ntt_tile[tile + e1] = r14[1];
    }
    __syncthreads();
  }
  for (uint32_t e = threadIdx.x; e < tile; e += NTT_THREADS) {
    uint4* dst = reinterpret_cast<uint4*>(&v[gidx(e)]);
    dst[0] = ntt_tile[e];
    dst[1] = ntt_tile[tile + e];
  }
}

// Passes of a transform of 2^logD points, in forward (DIF) stage order: the last pass takes the
// final min(TL, logD) stages on contiguous tiles (C = 1 group each); the stages before it are split
// evenly into passes of at most TL - 3 stages (C >= 8: 256-byte contiguous runs).
static int ntt_plan_passes(int logD, NttPass out[8]) {
  int np = 0;
  const int klast = logD < NTT_TL ? logD : NTT_TL;
  const int rest = logD - klast;
  if (rest > 0) {
    const int km = NTT_TL - 3;
    const int cnt = (rest + km - 1) / km;
    int s0 = 0;
    for (int i = 0; i < cnt; i++) {
      int K = rest / cnt + (i < rest % cnt ? 1 : 0);
      out[np++] = NttPass{s0, K, NTT_TL - K};
      s0 += K;
    }
  }
  out[np++] = NttPass{rest, klast, 0};
  return np;
}

static void ntt_launch(const NttPlan& p, Fr* d, int batch, bool inverse, cudaStream_t st) {
  NttPass passes[8];
  const int np = ntt_plan_passes(p.logD, passes);
  for (int i = 0; i < np; i++) {
    const NttPass& P = passes[inverse ? np - 1 - i : i];
    const uint32_t tile = 1u << (P.K + P.logC);
    dim3 grid(p.D / tile, batch);
    const size_t smem = 2 * (size_t)tile * sizeof(uint4);
    if (inverse) k_ntt_pass<true><<<grid, NTT_THREADS, smem, st>>>(d, p.tw_inv.p, p.D, p.logD, P);
    else k_ntt_pass<false><<<grid, NTT_THREADS, smem, st>>>(d, p.tw.p, p.D, p.logD, P);
  }
}

static bool ntt_fused() { static int f = env_int("ZKB200_NTT_FUSED", 1); return f != 0; }

static void ntt_forward_batch(const NttPlan& p, Fr* d, int batch, cudaStream_t st) {
  if (ntt_fused()) return ntt_launch(p, d, batch, false, st);
  dim3 grid(cdiv(p.D / 2, 256), batch);
  for (int s = 0; s < p.logD; s++) k_ntt_dif_stage<<<grid, 256, 0, st>>>(d, p.tw.p, p.D, p.logD - 1 - s, s);
}
static void ntt_inverse_batch(const NttPlan& p, Fr* d, int batch, cudaStream_t st) {
  if (ntt_fused()) return ntt_launch(p, d, batch, true, st);
  dim3 grid(cdiv(p.D / 2, 256), batch);
  for (int s = p.logD - 1; s >= 0; s--) k_ntt_dit_stage<<<grid, 256, 0, st>>>(d, p.tw_inv.p, p.D, p.logD - 1 - s, s);
}
void NttPlan::forward(Fr* d, cudaStream_t st) const { ntt_forward_batch(*this, d, 1, st); }
void NttPlan::inverse(Fr* d, cudaStream_t st) const { ntt_inverse_batch(*this, d, 1, st); }

// ---------------------------------------------------------------------------
// QAP kernels
// ---------------------------------------------------------------------------
// QAP.ml:121-131 eval':  out[i] = sum_k sol[k] * M[k][i]; blockIdx.y selects V / W / Y.
// Writes the n coefficients to coeffs (kept for the MSM scalars) and a zero-padded,
// coset-shifted copy of length D to work (input of the forward NTT).
constexpr int COMBINE_X = 32, COMBINE_K = 8;   // block = 32 coefficients x 8 slices of the variable range
static __global__ void __launch_bounds__(COMBINE_X * COMBINE_K)
k_qap_combine(const Fr* __restrict__ vm, const Fr* __restrict__ wm, const Fr* __restrict__ ym,
              const Fr* __restrict__ sol, uint32_t m, uint32_t n, uint32_t D, const Fr* __restrict__ coset,
              Fr* __restrict__ coeffs, Fr* __restrict__ work) {
  __shared__ Fr part[COMBINE_K][COMBINE_X];
  const uint32_t i = blockIdx.x * COMBINE_X + threadIdx.x;
  const int which = blockIdx.y;
  Fr acc = Fr::zero();
  if (i < n) {
    const Fr* M = which == 0 ? vm : (which == 1 ? wm : ym);
    for (uint32_t k = threadIdx.y; k < m; k += COMBINE_K) {   // adjacent threadIdx.x read adjacent coefficients
      Fr s = load_vec(&sol[k]);
      if (s.is_zero()) continue;
      acc = acc + s * load_vec(&M[(size_t)k * n + i]);
    }
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y != 0 || i >= D) return;
#pragma unroll
  for (int q = 1; q < COMBINE_K; q++) acc = acc + part[q][threadIdx.x];
  if (i < n) {
    store_vec(&coeffs[(size_t)which * n + i], acc);
    acc = acc * load_vec(&coset[i]);
  }
  store_vec(&work[(size_t)which * D + i], acc);
}

// zero-padded coset shift of an n-coefficient vector into a D-long work vector
static __global__ void k_coset_load(const Fr* __restrict__ coeffs, uint32_t n, uint32_t D,
                                    const Fr* __restrict__ coset, Fr* __restrict__ work) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  Fr a = Fr::zero();
  if (i < n) a = load_vec_rw(&coeffs[i]) * load_vec(&coset[i]);
  store_vec(&work[i], a);
}

// in-place inversion, 16 elements per thread (Montgomery's trick); a zero raises the flag
static __global__ void k_fr_batch_inverse(Fr* __restrict__ x, uint32_t n, int* flag) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t i0 = t * 16;
  if (i0 >= n) return;
  uint32_t cnt = min(16u, n - i0);
  Fr pre[16];
  Fr run = Fr::one();
  for (uint32_t j = 0; j < cnt; j++) {
    Fr a = load_vec_rw(&x[i0 + j]);
    if (a.is_zero()) { atomicExch(flag, 1); a = Fr::one(); }
    pre[j] = run;
    run = run * a;
  }
  Fr inv = run.inverse();
  for (int j = (int)cnt - 1; j >= 0; j--) {
    Fr a = load_vec_rw(&x[i0 + j]);
    if (a.is_zero()) a = Fr::one();
    store_vec(&x[i0 + j], inv * pre[j]);
    inv = inv * a;
  }
}

// H[j] = (V[j] W[j] - Y[j]) / t[j] on the coset
static __global__ void k_quotient_pointwise(const Fr* __restrict__ work, uint32_t D, const Fr* __restrict__ t_inv,
                                            Fr* __restrict__ H) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  Fr v = load_vec_rw(&work[j]), w = load_vec_rw(&work[(size_t)D + j]), y = load_vec_rw(&work[2 * (size_t)D + j]);
  store_vec(&H[j], (v * w - y) * load_vec(&t_inv[j]));
}

// Divisibility check (QAP.ml:134 `assert (is_zero rem)`).  h is interpolated from D >= n + 1 values of
// (V W - Y) / t on the coset and truncated to n - 1 coefficients, so two things are verified:
//  (1) the discarded coefficients n-1 .. D-1 are all zero (k_h_unshift_check) — exact;
//  (2) h(x) t(x) = V(x) W(x) - Y(x) at TWO points off the coset: a fixed x0, and x1 derived from
//      the values of all five polynomials at x0, so the second point depends on the witness and
//      cannot be aimed at in advance.  Both sides have degree <= 2n - 2: a non-divisible input
//      passes with probability <= (2n / r)^2 — the check is probabilistic, and DESIGN.md says so.
// The evaluation-form prover's check (k_eval_prepare: V(j) W(j) = Y(j) on every domain point) is exact.
struct EvalJob {
  const Fr* coeffs[5];
  uint32_t len[5];
};
constexpr int EVAL_CHUNK = 32, EVAL_THREADS = 256;
static __device__ __forceinline__ Fr eval_point() {
  // an arbitrary fixed element (Montgomery image of a constant); any point off the domain works
  Fr x;
  const uint32_t c[8] = {0x7f4a7c15u, 0x9e3779b9u, 0xf39cc060u, 0x5cedc834u, 0x1082276bu, 0xf3a8b2c1u, 0x2545f491u, 0x0f6c7d3au};
  for (int i = 0; i < 8; i++) x.v[i] = c[i];
  return x;
}
// xs[0] = x0 (fixed)
static __global__ void k_eval_point_init(Fr* xs) {
  if (threadIdx.x || blockIdx.x) return;
  xs[0] = eval_point();
}
static __global__ void __launch_bounds__(EVAL_THREADS)
k_poly_eval_partial(EvalJob job, const Fr* __restrict__ xp, Fr* __restrict__ partials, uint32_t blocks_per_poly) {
  __shared__ Fr sm[EVAL_THREADS];
  int which = blockIdx.y;
  const Fr* c = job.coeffs[which];
  uint32_t n = job.len[which];
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t i0 = t * EVAL_CHUNK;
  Fr acc = Fr::zero();
  if (i0 < n) {
    Fr x0 = load_vec_rw(xp);
    uint32_t end = min(n, i0 + EVAL_CHUNK);
    for (int i = (int)end - 1; i >= (int)i0; i--) acc = acc * x0 + load_vec_rw(&c[i]);
    acc = acc * fr_pow_u32(x0, i0);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = EVAL_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] = sm[threadIdx.x] + sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[(size_t)which * blocks_per_poly + blockIdx.x] = sm[0];
}
// Compares the two sides at the point just evaluated; next (nullable) receives the following point:
// x0 + sum_w e_w x0^(w+1), a polynomial mix of the five values.
static __global__ void k_poly_eval_check(const Fr* __restrict__ partials, uint32_t blocks_per_poly, const Fr* xp, Fr* next,
                                         int* flag) {
  if (threadIdx.x || blockIdx.x) return;
  Fr e[5];
  for (int w = 0; w < 5; w++) {
    Fr acc = Fr::zero();
    for (uint32_t b = 0; b < blocks_per_poly; b++) acc = acc + load_vec_rw(&partials[(size_t)w * blocks_per_poly + b]);
    e[w] = acc;
  }
  // order: V, W, Y, t, h
  if (e[4] * e[3] != e[0] * e[1] - e[2]) atomicExch(flag, 1);
  if (next) {
    Fr x = load_vec_rw(xp), mix = x, pw = x;
    for (int w = 0; w < 5; w++) { mix = mix + e[w] * pw; pw = pw * x; }
    *next = mix;
  }
}
// H <- H .* coset_inv (undo the coset shift); coefficients n - 1 .. D - 1 must vanish
static __global__ void k_h_unshift_check(Fr* __restrict__ H, const Fr* __restrict__ coset_inv, uint32_t D, uint32_t keep,
                                         int* flag) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  Fr h = load_vec_rw(&H[i]);
  if (i >= keep) {
    if (!h.is_zero()) atomicExch(flag, 1);
    return;
  }
  store_vec(&H[i], h * load_vec(&coset_inv[i]));
}

// ---------------------------------------------------------------------------
// QapDevice
// ---------------------------------------------------------------------------
static int log2_ceil_u32(uint32_t n) {
  int l = 0;
  while ((1ull << l) < n) l++;
  return l;
}

// ---------------------------------------------------------------------------
// EvalDomain (evaluation-form quotient)
// ---------------------------------------------------------------------------
// out[which][j] = sum over row j of val * sol[col]   (blockIdx.y = matrix)
struct CsrPtrs {
  const uint32_t* row_ptr[3];
  const uint32_t* col[3];
  const Fr* val[3];
};
static __global__ void __launch_bounds__(128)
k_csr_matvec(CsrPtrs m, const Fr* __restrict__ sol, uint32_t n, uint32_t n_vars, Fr* __restrict__ out) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int which = blockIdx.y;
  const uint32_t* rp = m.row_ptr[which];
  Fr acc = Fr::zero();
  for (uint32_t k = rp[j]; k < rp[j + 1]; k++) {
    ZK_DCHECK(m.col[which][k] < n_vars);
    acc = acc + load_vec(&m.val[which][k]) * load_vec(&sol[m.col[which][k]]);
  }
  store_vec(&out[(size_t)which * n + j], acc);
}
// gate check V(j) W(j) == Y(j) and a = w .* evals zero-padded to D (3 vectors)
static __global__ void __launch_bounds__(128)
k_eval_prepare(const Fr* __restrict__ evals, const Fr* __restrict__ w, uint32_t n, uint32_t D, Fr* __restrict__ work,
               int* flag) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  Fr a[3] = {Fr::zero(), Fr::zero(), Fr::zero()};
  if (i < n) {
    Fr wi = load_vec(&w[i]);
    Fr v = load_vec_rw(&evals[i]), ww = load_vec_rw(&evals[(size_t)n + i]), y = load_vec_rw(&evals[2 * (size_t)n + i]);
    if (v * ww != y) atomicExch(flag, 1);
    a[0] = v * wi; a[1] = ww * wi; a[2] = y * wi;
  }
  for (int q = 0; q < 3; q++) store_vec(&work[(size_t)q * D + i], a[q]);
}
static __global__ void k_mul_by_ghat(Fr* __restrict__ work, const Fr* __restrict__ ghat, uint32_t D) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  Fr g = load_vec(&ghat[i]);
  for (int q = 0; q < 3; q++) store_vec(&work[(size_t)q * D + i], load_vec_rw(&work[(size_t)q * D + i]) * g);
}
// H[k] = c1[k] * SV[n+k] * SW[n+k] - c2 * SY[n+k]
static __global__ void k_eval_quotient(const Fr* __restrict__ work, const Fr* __restrict__ c1, const Fr* __restrict__ c2,
                                       uint32_t n, uint32_t D, Fr* __restrict__ H) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  Fr sv = load_vec_rw(&work[n + k]), sw = load_vec_rw(&work[(size_t)D + n + k]), sy = load_vec_rw(&work[2 * (size_t)D + n + k]);
  store_vec(&H[k], load_vec(&c1[k]) * sv * sw - load_vec_rw(c2) * sy);
}
// kernel 1/d: g[0] = 0, g[i] = to_mont(i) (inverted afterwards), zero beyond 2n
static __global__ void k_fill_index(Fr* __restrict__ g, uint32_t limit, uint32_t D) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  Fr a = Fr::zero();
  if (i >= 1 && i < limit) { a.v[0] = i; a = a.to_mont(); } else a = Fr::one();   // placeholders inverted to 1
  store_vec(&g[i], a);
}
static __global__ void k_clear_outside(Fr* __restrict__ g, uint32_t limit, uint32_t D) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  if (i == 0 || i >= limit) store_vec(&g[i], Fr::zero());
}
// c1[k] = t_shift[k] * dinv^2
static __global__ void k_scale_c1(Fr* __restrict__ c1, const Fr* __restrict__ dinv, uint32_t n) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  Fr d = load_vec_rw(dinv);
  store_vec(&c1[k], load_vec_rw(&c1[k]) * d * d);
}

void EvalDomain::load(uint32_t n_, const uint8_t* w_raw, const uint8_t* t_shift_raw, cudaStream_t st) {
  ZK_REQUIRE(n_ >= 2 && n_ < (1u << 26), ZK_EARG, "eval_domain_load: bad size");
  n = n_;
  int logD = 1;
  while ((1ull << logD) < 2ull * n) logD++;
  plan.build(logD, st);
  const uint32_t D = plan.D;
  flag.alloc(2);
  ZK_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), st));
  w.alloc(n); c1.alloc(n); c2.alloc(1); ghat.alloc(D);
  {
    DevBuf<uint32_t> raw((size_t)n * 8);
    ZK_CUDA(cudaMemcpyAsync(raw.p, w_raw, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    fr_to_mont(raw.p, w.p, n, flag.p, st);
    ZK_CUDA(cudaMemcpyAsync(raw.p, t_shift_raw, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    fr_to_mont(raw.p, c1.p, n, flag.p, st);
    ZK_CUDA(cudaStreamSynchronize(st));
  }
  {
    DevBuf<Fr> consts(5);
    k_ntt_consts<<<1, 1, 0, st>>>(logD, consts.p);
    ZK_CUDA(cudaMemcpyAsync(c2.p, consts.p + 4, sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    k_scale_c1<<<cdiv(n, 256), 256, 0, st>>>(c1.p, c2.p, n);
    ZK_CUDA(cudaStreamSynchronize(st));
  }
  k_fill_index<<<cdiv(D, 256), 256, 0, st>>>(ghat.p, 2 * n, D);
  k_fr_batch_inverse<<<cdiv(cdiv(D, 16), 64), 64, 0, st>>>(ghat.p, D, flag.p + 1);
  k_clear_outside<<<cdiv(D, 256), 256, 0, st>>>(ghat.p, 2 * n, D);
  plan.forward(ghat.p, st);
  evals.alloc(3 * (size_t)n);
  work.alloc(3 * (size_t)D);
  H.alloc(n);
  int fl[2];
  ZK_CUDA(cudaMemcpyAsync(fl, flag.p, sizeof(fl), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(fl[0] == 0, ZK_EPOINT, "eval_domain_load: scalar is not canonical (>= r)");
}

void EvalDomain::load_matrix(int which, uint32_t m_, const uint32_t* row_ptr, const uint32_t* col, const uint8_t* val_raw,
                             cudaStream_t st) {
  ZK_REQUIRE(which >= 0 && which < 3 && row_ptr && m_ > 0, ZK_EARG, "r1cs_load: bad arguments");
  ZK_REQUIRE(m == 0 || m == m_, ZK_EARG, "r1cs_load: matrices disagree on the number of variables");
  m = m_;
  uint32_t nnz = row_ptr[n];
  for (uint32_t j = 0; j < n; j++) ZK_REQUIRE(row_ptr[j] <= row_ptr[j + 1], ZK_EARG, "r1cs_load: row_ptr not monotone");
  for (uint32_t k = 0; k < nnz; k++) ZK_REQUIRE(col[k] < m, ZK_EARG, "r1cs_load: column index out of range");
  SparseMat& M = mat[which];
  M.row_ptr.alloc(n + 1);
  M.col.alloc(nnz ? nnz : 1);
  M.val.alloc(nnz ? nnz : 1);
  ZK_CUDA(cudaMemcpyAsync(M.row_ptr.p, row_ptr, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
  if (nnz) {
    DevBuf<uint32_t> raw((size_t)nnz * 8);
    ZK_CUDA(cudaMemcpyAsync(M.col.p, col, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    ZK_CUDA(cudaMemcpyAsync(raw.p, val_raw, (size_t)nnz * 32, cudaMemcpyHostToDevice, st));
    fr_to_mont(raw.p, M.val.p, nnz, flag.p, st);
    ZK_CUDA(cudaStreamSynchronize(st));
  }
  sol_m.ensure(m);
}

void EvalDomain::eval(const uint32_t* d_sol_raw, cudaStream_t st) {
  values(d_sol_raw, st);
  quotient(st);
}

void EvalDomain::values(const uint32_t* d_sol_raw, cudaStream_t st) {
  const uint32_t D = plan.D;
  ZK_REQUIRE(mat[0].row_ptr.p && mat[1].row_ptr.p && mat[2].row_ptr.p, ZK_EARG, "prove_r1cs: matrices not loaded");
  ZK_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), st));
  fr_to_mont(d_sol_raw, sol_m.p, m, flag.p, st);
  CsrPtrs P;
  for (int q = 0; q < 3; q++) { P.row_ptr[q] = mat[q].row_ptr.p; P.col[q] = mat[q].col.p; P.val[q] = mat[q].val.p; }
  k_csr_matvec<<<dim3(cdiv(n, 128), 3), 128, 0, st>>>(P, sol_m.p, n, m, evals.p);
  k_eval_prepare<<<cdiv(D, 128), 128, 0, st>>>(evals.p, w.p, n, D, work.p, flag.p + 1);
  ZK_CUDA(cudaGetLastError());
}

void EvalDomain::quotient(cudaStream_t st) {
  const uint32_t D = plan.D;
  ntt_forward_batch(plan, work.p, 3, st);
  k_mul_by_ghat<<<cdiv(D, 256), 256, 0, st>>>(work.p, ghat.p, D);
  ntt_inverse_batch(plan, work.p, 3, st);
  k_eval_quotient<<<cdiv(n, 256), 256, 0, st>>>(work.p, c1.p, c2.p, n, D, H.p);
  ZK_CUDA(cudaGetLastError());
}

struct QuotientScratch {
  DevBuf<Fr> work;  // 3 * D
};

void QapDevice::load(const uint8_t* v, const uint8_t* w, const uint8_t* y, const uint8_t* tgt, uint32_t m_, uint32_t n_,
                     cudaStream_t st) {
  ZK_REQUIRE(n_ >= 1 && (uint64_t)m_ * n_ < (1ull << 31), ZK_EARG, "qap_load: bad dimensions");
  m = m_;
  n = n_;
  plan.build(log2_ceil_u32(n + 1) < 1 ? 1 : log2_ceil_u32(n + 1), st);
  const uint32_t D = plan.D;
  flag.alloc(2);
  ZK_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), st));
  size_t mn = (size_t)m * n;
  {
    DevBuf<uint32_t> raw(mn * 8 > (size_t)(n + 1) * 8 ? mn * 8 : (size_t)(n + 1) * 8);
    auto up = [&](const uint8_t* src, DevBuf<Fr>& dst, size_t count) {
      dst.alloc(count);
      if (!count) return;
      ZK_CUDA(cudaMemcpyAsync(raw.p, src, count * 32, cudaMemcpyHostToDevice, st));
      fr_to_mont(raw.p, dst.p, (uint32_t)count, flag.p, st);
    };
    if (m) { up(v, vm, mn); up(w, wm, mn); up(y, ym, mn); }
    up(tgt, target, n + 1);
    ZK_CUDA(cudaStreamSynchronize(st));
  }
  // 1 / t on the coset (bit-reversed, matching the forward transform's output order)
  t_inv_evals.alloc(D);
  k_coset_load<<<cdiv(D, 256), 256, 0, st>>>(target.p, n + 1, D, plan.coset.p, t_inv_evals.p);
  plan.forward(t_inv_evals.p, st);
  k_fr_batch_inverse<<<cdiv(cdiv(D, 16), 64), 64, 0, st>>>(t_inv_evals.p, D, flag.p + 1);
  ZK_CUDA(cudaGetLastError());
  int fl[2] = {0, 0};
  ZK_CUDA(cudaMemcpyAsync(fl, flag.p, sizeof(fl), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(fl[0] == 0, ZK_EPOINT, "qap_load: coefficient is not canonical (>= r)");
  ZK_REQUIRE(fl[1] == 0, ZK_EARG, "qap_load: target vanishes on the evaluation coset");
  sol_m.alloc(m ? m : 1);
  V.alloc(3 * (size_t)D);       // work: V | W | Y evaluations
  H.alloc(D);
  Vc.alloc(3 * (size_t)n);      // coefficient copies: V | W | Y, n each
  uint32_t bpp = cdiv(cdiv(n + 1, EVAL_CHUNK), EVAL_THREADS);
  partials.alloc(5 * (size_t)bpp);
  eval_xs.alloc(2);
  ZK_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), st));
}

void QapDevice::eval(const uint32_t* d_sol_raw, cudaStream_t st) {
  combine(d_sol_raw, st);
  quotient_from_work(st);
}

void QapDevice::combine(const uint32_t* d_sol_raw, cudaStream_t st) {
  const uint32_t D = plan.D;
  ZK_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), st));
  fr_to_mont(d_sol_raw, sol_m.p, m, flag.p, st);
  k_qap_combine<<<dim3(cdiv(D, COMBINE_X), 3), dim3(COMBINE_X, COMBINE_K), 0, st>>>(vm.p, wm.p, ym.p, sol_m.p, m, n, D,
                                                                                     plan.coset.p, Vc.p, V.p);
  ZK_CUDA(cudaGetLastError());
}

void QapDevice::quotient_from_work(cudaStream_t st) {
  const uint32_t D = plan.D;
  ntt_forward_batch(plan, V.p, 3, st);
  k_quotient_pointwise<<<cdiv(D, 256), 256, 0, st>>>(V.p, D, t_inv_evals.p, H.p);
  plan.inverse(H.p, st);
  k_h_unshift_check<<<cdiv(D, 256), 256, 0, st>>>(H.p, plan.coset_inv.p, D, n > 0 ? n - 1 : 0, flag.p + 1);
  // divisibility check at x0 and at a witness-dependent x1
  EvalJob job;
  job.coeffs[0] = Vc.p; job.len[0] = n;
  job.coeffs[1] = Vc.p + n; job.len[1] = n;
  job.coeffs[2] = Vc.p + 2 * (size_t)n; job.len[2] = n;
  job.coeffs[3] = target.p; job.len[3] = n + 1;
  job.coeffs[4] = H.p; job.len[4] = n > 0 ? n - 1 : 0;
  uint32_t bpp = cdiv(cdiv(n + 1, EVAL_CHUNK), EVAL_THREADS);
  k_eval_point_init<<<1, 1, 0, st>>>(eval_xs.p);
  for (int round = 0; round < 2; round++) {
    k_poly_eval_partial<<<dim3(bpp, 5), EVAL_THREADS, 0, st>>>(job, eval_xs.p + round, partials.p, bpp);
    k_poly_eval_check<<<1, 1, 0, st>>>(partials.p, bpp, eval_xs.p + round, round == 0 ? eval_xs.p + 1 : nullptr, flag.p + 1);
  }
  ZK_CUDA(cudaGetLastError());
}

void QapDevice::set_coeffs(const uint32_t* d_vwy_raw, cudaStream_t st) {
  const uint32_t D = plan.D;
  ZK_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), st));
  fr_to_mont(d_vwy_raw, Vc.p, 3 * n, flag.p, st);
  for (int which = 0; which < 3; which++)
    k_coset_load<<<cdiv(D, 256), 256, 0, st>>>(Vc.p + (size_t)which * n, n, D, plan.coset.p, V.p + (size_t)which * D);
}

}  // namespace zk

extern "C" {

int zk_qap_load(const uint8_t* v, const uint8_t* w, const uint8_t* y, const uint8_t* target, size_t m, size_t n,
                uint64_t* handle) {
  ZK_API_BEGIN
  using namespace zk;
  ZK_REQUIRE(v && w && y && target && handle && m > 0 && n > 0 && m < (1u << 28) && n < (1u << 27), ZK_EARG,
             "qap_load: bad arguments");
  auto h = std::make_unique<QapHandle>();
  h->q.load(v, w, y, target, (uint32_t)m, (uint32_t)n, default_stream());
  *handle = register_handle(std::move(h));
  ZK_API_END
}

int zk_quotient_domain_load(const uint8_t* target, size_t n, uint64_t* handle) {
  ZK_API_BEGIN
  using namespace zk;
  ZK_REQUIRE(target && handle && n >= 2 && n < (1u << 27), ZK_EARG, "quotient_domain_load: bad arguments");
  auto h = std::make_unique<QapHandle>();
  h->q.load(nullptr, nullptr, nullptr, target, 0, (uint32_t)n, default_stream());
  *handle = register_handle(std::move(h));
  ZK_API_END
}

int zk_qap_free(uint64_t handle) {
  ZK_API_BEGIN
  ZK_CUDA(cudaDeviceSynchronize());
  zk::HandleBase* h = zk::lookup_handle(handle, 0);
  ZK_REQUIRE(h->kind == 3 || h->kind == 6, ZK_EARG, "qap_free: not a QAP / domain handle");
  zk::drop_handle(handle);
  ZK_API_END
}

int zk_eval_domain_load(size_t n, const uint8_t* w, const uint8_t* t_shift, uint64_t* handle) {
  ZK_API_BEGIN
  using namespace zk;
  ZK_REQUIRE(w && t_shift && handle, ZK_EARG, "eval_domain_load: null argument");
  auto h = std::make_unique<EvalDomainHandle>();
  h->d.load((uint32_t)n, w, t_shift, default_stream());
  *handle = register_handle(std::move(h));
  ZK_API_END
}

int zk_r1cs_load(uint64_t domain, int which, size_t m, const uint32_t* row_ptr, const uint32_t* col, const uint8_t* val) {
  ZK_API_BEGIN
  using namespace zk;
  auto* h = static_cast<EvalDomainHandle*>(lookup_handle(domain, 6));
  h->d.load_matrix(which, (uint32_t)m, row_ptr, col, val, default_stream());
  ZK_API_END
}

// h: (n - 1) * 32 bytes out (nullable); vwy: 3 * n * 32 bytes out (nullable)
int zk_qap_eval(uint64_t handle, const uint8_t* sol, uint8_t* h_out, uint8_t* vwy_out) {
  ZK_API_BEGIN
  using namespace zk;
  auto* h = static_cast<QapHandle*>(lookup_handle(handle, 3));
  QapDevice& q = h->q;
  ZK_REQUIRE(sol, ZK_EARG, "qap_eval: null witness");
  cudaStream_t st = default_stream();
  size_t need = std::max((size_t)q.m * 8, 3 * (size_t)q.n * 8);
  h->d_raw.ensure(need);
  ZK_CUDA(cudaMemcpyAsync(h->d_raw.p, sol, (size_t)q.m * 32, cudaMemcpyHostToDevice, st));
  q.eval(h->d_raw.p, st);
  int fl[2];
  ZK_CUDA(cudaMemcpyAsync(fl, q.flag.p, sizeof(fl), cudaMemcpyDeviceToHost, st));
  if (h_out && q.n > 1) {
    fr_from_mont(q.H.p, h->d_raw.p, q.n - 1, st);
    ZK_CUDA(cudaMemcpyAsync(h_out, h->d_raw.p, (size_t)(q.n - 1) * 32, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
  }
  if (vwy_out) {
    fr_from_mont(q.Vc.p, h->d_raw.p, 3 * q.n, st);
    ZK_CUDA(cudaMemcpyAsync(vwy_out, h->d_raw.p, 3 * (size_t)q.n * 32, cudaMemcpyDeviceToHost, st));
  }
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(fl[0] == 0, ZK_EPOINT, "qap_eval: witness scalar is not canonical (>= r)");
  ZK_REQUIRE(fl[1] == 0, ZK_EREMAINDER, "qap_eval: V*W - Y is not divisible by the target (QAP.ml:134)");
  ZK_API_END
}

// Standalone quotient from coefficient vectors: V, W, Y of n coefficients, T of n + 1; h gets n - 1.
int zk_fr_quotient(const uint8_t* V, const uint8_t* W, const uint8_t* Y, const uint8_t* T, size_t n, uint8_t* h_out) {
  ZK_API_BEGIN
  using namespace zk;
  ZK_REQUIRE(V && W && Y && T && h_out && n >= 2 && n < (1u << 27), ZK_EARG, "fr_quotient: bad arguments");
  cudaStream_t st = default_stream();
  QapDevice q;
  q.load(nullptr, nullptr, nullptr, T, 0, (uint32_t)n, st);
  DevBuf<uint32_t> raw(3 * n * 8);
  ZK_CUDA(cudaMemcpyAsync(raw.p, V, n * 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(raw.p + n * 8, W, n * 32, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(raw.p + 2 * n * 8, Y, n * 32, cudaMemcpyHostToDevice, st));
  q.set_coeffs(raw.p, st);
  q.quotient_from_work(st);
  int fl[2];
  ZK_CUDA(cudaMemcpyAsync(fl, q.flag.p, sizeof(fl), cudaMemcpyDeviceToHost, st));
  fr_from_mont(q.H.p, raw.p, (uint32_t)n - 1, st);
  ZK_CUDA(cudaMemcpyAsync(h_out, raw.p, (n - 1) * 32, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(cudaStreamSynchronize(st));
  ZK_REQUIRE(fl[0] == 0, ZK_EPOINT, "fr_quotient: coefficient is not canonical (>= r)");
  ZK_REQUIRE(fl[1] == 0, ZK_EREMAINDER, "fr_quotient: V*W - Y is not divisible by T");
  ZK_API_END
}

}  // extern "C"
