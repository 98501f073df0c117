/* CPU oracle in C — TEST / BASELINE INFRASTRUCTURE ONLY (never linked into libzkb200).
 *
 * PARITY UNPINNED at the bls12-381 boundary (see oracle/bls12_381.py); this file is checked
 * against the Python oracle in tests/test_cpu_c_oracle.py.
 *
 * Restates, with plain 64-bit limb arithmetic, the reference's G1 "MSM":
 *   curve.ml:91      sum_map  = Var.Map.fold (fun k v acc -> f k v + acc) m zero
 *   curve.ml:94-103  dot      = sum_map m (fun k mk -> mk * c_k)
 *   curve.ml:112-118 apply_powers: loop (x * c + acc)
 * i.e. a left fold of one full scalar multiplication (double-and-add, as G1.mul of the
 * bls12-381 package does per call) plus one addition per term.  The reference is
 * single-threaded; `threads` > 1 splits the index range into contiguous chunks, folds each
 * chunk on its own thread and adds the partial sums — the same arithmetic, all host cores.
 *
 * Second entry point, NOT the reference's algorithm: zkoracle_g1_msm_pippenger, a plain
 * multi-threaded bucket method (signed c-bit digits, mixed additions, running-sum bucket
 * reduction) — the "fair CPU" line of SURVEY.md §8(d): what a CPU does with the same algorithm
 * family the GPU path uses.  Reported beside cpu_baseline, never as it.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[6]; } fp;

static const fp P = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL, 0x64774b84f38512bfULL,
                      0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
static const uint64_t N0 = 0x89f3fffcfffcfffdULL;                 /* -p^-1 mod 2^64 */
static const fp R2 = {{0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL, 0x67eb88a9939d83c0ULL,
                       0x9a793e85b519952dULL, 0x11988fe592cae3aaULL}};
static const fp ONE = {{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL, 0x77ce585370525745ULL,
                        0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}};

static int fp_is_zero(const fp *a) { uint64_t o = 0; for (int i = 0; i < 6; i++) o |= a->v[i]; return o == 0; }
static int fp_eq(const fp *a, const fp *b) { return memcmp(a, b, sizeof(fp)) == 0; }
static int fp_geq_p(const uint64_t *t) {
  for (int i = 5; i >= 0; i--) { if (t[i] > P.v[i]) return 1; if (t[i] < P.v[i]) return 0; }
  return 1;
}
static void fp_sub_p(uint64_t *t) {
  u128 br = 0;
  for (int i = 0; i < 6; i++) { u128 d = (u128)t[i] - P.v[i] - br; t[i] = (uint64_t)d; br = (d >> 64) & 1; }
}
static void fp_add(fp *r, const fp *a, const fp *b) {
  u128 c = 0; uint64_t t[6];
  for (int i = 0; i < 6; i++) { c += (u128)a->v[i] + b->v[i]; t[i] = (uint64_t)c; c >>= 64; }
  if (c || fp_geq_p(t)) fp_sub_p(t);
  memcpy(r->v, t, sizeof t);
}
static void fp_sub(fp *r, const fp *a, const fp *b) {
  u128 br = 0; uint64_t t[6];
  for (int i = 0; i < 6; i++) { u128 d = (u128)a->v[i] - b->v[i] - br; t[i] = (uint64_t)d; br = (d >> 64) & 1; }
  if (br) { u128 c = 0; for (int i = 0; i < 6; i++) { c += (u128)t[i] + P.v[i]; t[i] = (uint64_t)c; c >>= 64; } }
  memcpy(r->v, t, sizeof t);
}
static void fp_mul(fp *r, const fp *a, const fp *b) {           /* CIOS Montgomery */
  uint64_t t[8] = {0};
  for (int i = 0; i < 6; i++) {
    u128 c = 0;
    for (int j = 0; j < 6; j++) { c += (u128)a->v[j] * b->v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[6]; t[6] = (uint64_t)c; t[7] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * N0;
    c = ((u128)m * P.v[0] + t[0]) >> 64;
    for (int j = 1; j < 6; j++) { c += (u128)m * P.v[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[6]; t[5] = (uint64_t)c; t[6] = t[7] + (uint64_t)(c >> 64);
  }
  if (t[6] || fp_geq_p(t)) fp_sub_p(t);
  memcpy(r->v, t, 6 * sizeof(uint64_t));
}
static void fp_sqr(fp *r, const fp *a) { fp_mul(r, a, a); }
static void fp_inv(fp *r, const fp *a) {                          /* a^(p-2) */
  fp e = P, acc = ONE;
  e.v[0] -= 2;
  for (int i = 5; i >= 0; i--)
    for (int b = 63; b >= 0; b--) { fp_sqr(&acc, &acc); if ((e.v[i] >> b) & 1) fp_mul(&acc, &acc, a); }
  *r = acc;
}
static void fp_from_be(fp *r, const uint8_t *b) {
  fp raw;
  for (int i = 0; i < 6; i++) { uint64_t w = 0; for (int k = 0; k < 8; k++) w = (w << 8) | b[(5 - i) * 8 + k]; raw.v[i] = w; }
  fp_mul(r, &raw, &R2);
}
static void fp_to_be(uint8_t *b, const fp *a) {
  fp one_raw = {{1, 0, 0, 0, 0, 0}}, raw;
  fp_mul(&raw, a, &one_raw);
  for (int i = 0; i < 6; i++) for (int k = 0; k < 8; k++) b[(5 - i) * 8 + k] = (uint8_t)(raw.v[i] >> (56 - 8 * k));
}

/* Jacobian points, identity = Z == 0 */
typedef struct { fp X, Y, Z; } g1;
static void g1_set_inf(g1 *p) { memset(p, 0, sizeof *p); }
static int g1_is_inf(const g1 *p) { return fp_is_zero(&p->Z); }
static void g1_dbl(g1 *r, const g1 *p) {                          /* dbl-2009-l, a = 0 */
  if (g1_is_inf(p)) { *r = *p; return; }
  fp A, B, C, D, E, F, t;
  fp_sqr(&A, &p->X); fp_sqr(&B, &p->Y); fp_sqr(&C, &B);
  fp_add(&t, &p->X, &B); fp_sqr(&t, &t); fp_sub(&t, &t, &A); fp_sub(&t, &t, &C); fp_add(&D, &t, &t);
  fp_add(&E, &A, &A); fp_add(&E, &E, &A); fp_sqr(&F, &E);
  fp Z3; fp_mul(&Z3, &p->Y, &p->Z); fp_add(&Z3, &Z3, &Z3);
  fp X3; fp_sub(&X3, &F, &D); fp_sub(&X3, &X3, &D);
  fp C8; fp_add(&C8, &C, &C); fp_add(&C8, &C8, &C8); fp_add(&C8, &C8, &C8);
  fp Y3; fp_sub(&t, &D, &X3); fp_mul(&Y3, &E, &t); fp_sub(&Y3, &Y3, &C8);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void g1_add(g1 *r, const g1 *p, const g1 *q) {             /* add-2007-bl */
  if (g1_is_inf(p)) { *r = *q; return; }
  if (g1_is_inf(q)) { *r = *p; return; }
  fp Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t;
  fp_sqr(&Z1Z1, &p->Z); fp_sqr(&Z2Z2, &q->Z);
  fp_mul(&U1, &p->X, &Z2Z2); fp_mul(&U2, &q->X, &Z1Z1);
  fp_mul(&S1, &p->Y, &q->Z); fp_mul(&S1, &S1, &Z2Z2);
  fp_mul(&S2, &q->Y, &p->Z); fp_mul(&S2, &S2, &Z1Z1);
  if (fp_eq(&U1, &U2)) { if (fp_eq(&S1, &S2)) { g1_dbl(r, p); } else g1_set_inf(r); return; }
  fp_sub(&H, &U2, &U1); fp_add(&I, &H, &H); fp_sqr(&I, &I); fp_mul(&J, &H, &I);
  fp_sub(&rr, &S2, &S1); fp_add(&rr, &rr, &rr); fp_mul(&V, &U1, &I);
  fp X3, Y3, Z3;
  fp_sqr(&X3, &rr); fp_sub(&X3, &X3, &J); fp_sub(&X3, &X3, &V); fp_sub(&X3, &X3, &V);
  fp_sub(&t, &V, &X3); fp_mul(&Y3, &rr, &t); fp_mul(&t, &S1, &J); fp_add(&t, &t, &t); fp_sub(&Y3, &Y3, &t);
  fp_add(&Z3, &p->Z, &q->Z); fp_sqr(&Z3, &Z3); fp_sub(&Z3, &Z3, &Z1Z1); fp_sub(&Z3, &Z3, &Z2Z2); fp_mul(&Z3, &Z3, &H);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}
/* G1.mul: MSB-first double-and-add over the 255-bit scalar (32 B little-endian) */
static void g1_mul(g1 *r, const g1 *p, const uint8_t *k) {
  g1 acc; g1_set_inf(&acc);
  for (int i = 31; i >= 0; i--)
    for (int b = 7; b >= 0; b--) { g1_dbl(&acc, &acc); if ((k[i] >> b) & 1) g1_add(&acc, &acc, p); }
  *r = acc;
}
static int g1_from_raw(g1 *p, const uint8_t *b) {
  if (b[0] & 0x40) { g1_set_inf(p); return 0; }
  fp_from_be(&p->X, b); fp_from_be(&p->Y, b + 48); p->Z = ONE;
  return 0;
}
static void g1_to_raw(uint8_t *b, const g1 *p) {
  if (g1_is_inf(p)) { memset(b, 0, 96); b[0] = 0x40; return; }
  fp zi, zi2, zi3, x, y;
  fp_inv(&zi, &p->Z); fp_sqr(&zi2, &zi); fp_mul(&zi3, &zi2, &zi);
  fp_mul(&x, &p->X, &zi2); fp_mul(&y, &p->Y, &zi3);
  fp_to_be(b, &x); fp_to_be(b + 48, &y);
}

typedef struct { const uint8_t *bases, *scalars; size_t lo, hi; g1 acc; } job;
static void *fold_range(void *arg) {
  job *j = (job *)arg;
  g1 acc; g1_set_inf(&acc);
  for (size_t i = j->lo; i < j->hi; i++) {                       /* curve.ml:91: f k v + acc */
    g1 p, t;
    g1_from_raw(&p, j->bases + 96 * i);
    g1_mul(&t, &p, j->scalars + 32 * i);
    g1_add(&acc, &t, &acc);
  }
  j->acc = acc;
  return NULL;
}

/* out = sum_i scalars[i] * bases[i] by the reference's fold; returns 0 */
int zkoracle_g1_msm_fold(const uint8_t *bases, const uint8_t *scalars, size_t n, int threads, uint8_t out[96]) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  job *jobs = (job *)calloc(threads, sizeof(job));
  pthread_t *th = (pthread_t *)calloc(threads, sizeof(pthread_t));
  for (int t = 0; t < threads; t++) {
    jobs[t].bases = bases; jobs[t].scalars = scalars;
    jobs[t].lo = n * (size_t)t / threads; jobs[t].hi = n * (size_t)(t + 1) / threads;
    if (threads == 1) fold_range(&jobs[t]); else pthread_create(&th[t], NULL, fold_range, &jobs[t]);
  }
  g1 acc; g1_set_inf(&acc);
  for (int t = 0; t < threads; t++) { if (threads > 1) pthread_join(th[t], NULL); g1_add(&acc, &jobs[t].acc, &acc); }
  g1_to_raw(out, &acc);
  free(jobs); free(th);
  return 0;
}

/* ---- "fair CPU" bucket method ------------------------------------------------------------ */
/* r = p + (x2, y2) with the addend affine (madd-2007-bl: 7M + 4S); handles p = inf, p = +-q */
static void g1_madd(g1 *r, const g1 *p, const fp *x2, const fp *y2) {
  if (g1_is_inf(p)) { r->X = *x2; r->Y = *y2; r->Z = ONE; return; }
  fp Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
  fp_sqr(&Z1Z1, &p->Z); fp_mul(&U2, x2, &Z1Z1);
  fp_mul(&S2, y2, &p->Z); fp_mul(&S2, &S2, &Z1Z1);
  if (fp_eq(&U2, &p->X)) {
    if (fp_eq(&S2, &p->Y)) { g1 q; q.X = *x2; q.Y = *y2; q.Z = ONE; g1_dbl(r, &q); } else g1_set_inf(r);
    return;
  }
  fp_sub(&H, &U2, &p->X); fp_sqr(&HH, &H); fp_add(&I, &HH, &HH); fp_add(&I, &I, &I); fp_mul(&J, &H, &I);
  fp_sub(&rr, &S2, &p->Y); fp_add(&rr, &rr, &rr); fp_mul(&V, &p->X, &I);
  fp X3, Y3, Z3;
  fp_sqr(&X3, &rr); fp_sub(&X3, &X3, &J); fp_sub(&X3, &X3, &V); fp_sub(&X3, &X3, &V);
  fp_sub(&t, &V, &X3); fp_mul(&Y3, &rr, &t); fp_mul(&t, &p->Y, &J); fp_add(&t, &t, &t); fp_sub(&Y3, &Y3, &t);
  fp_add(&Z3, &p->Z, &H); fp_sqr(&Z3, &Z3); fp_sub(&Z3, &Z3, &Z1Z1); fp_sub(&Z3, &Z3, &HH);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}

typedef struct { fp x, y; int inf; } g1aff;
typedef struct {
  const g1aff *pts; const int32_t *digits;   /* digits[i * W + w], signed, |d| <= 2^(c-1) */
  size_t lo, hi; int w, W, c; g1 acc;
} pjob;

/* one (window, point range) job: fill 2^(c-1) buckets, reduce them with the running sum */
static void *pippenger_job(void *arg) {
  pjob *j = (pjob *)arg;
  const size_t nb = (size_t)1 << (j->c - 1);
  g1 *bk = (g1 *)calloc(nb + 1, sizeof(g1));                    /* Z = 0: identity */
  for (size_t i = j->lo; i < j->hi; i++) {
    int32_t d = j->digits[i * j->W + j->w];
    if (d == 0 || j->pts[i].inf) continue;
    fp y = j->pts[i].y;
    if (d < 0) { fp z; memset(&z, 0, sizeof z); fp_sub(&y, &z, &y); d = -d; }
    g1_madd(&bk[d], &bk[d], &j->pts[i].x, &y);
  }
  g1 run, sum; g1_set_inf(&run); g1_set_inf(&sum);
  for (size_t b = nb; b >= 1; b--) { g1_add(&run, &run, &bk[b]); g1_add(&sum, &sum, &run); }
  free(bk);
  j->acc = sum;
  return NULL;
}

typedef struct { pjob *jobs; int first, step, count; } pworker;
static void *pippenger_worker(void *arg) {
  pworker *w = (pworker *)arg;
  for (int k = w->first; k < w->count; k += w->step) pippenger_job(&w->jobs[k]);
  return NULL;
}

/* out = sum_i scalars[i] * bases[i] by the bucket method with c-bit signed windows (2 <= c <= 20);
 * returns 0, or -1 on a bad argument */
int zkoracle_g1_msm_pippenger(const uint8_t *bases, const uint8_t *scalars, size_t n, int c, int threads,
                              uint8_t out[96]) {
  if (c < 2 || c > 20 || n == 0) return -1;
  if (threads < 1) threads = 1;
  const int W = (255 + c) / c;                                  /* room for the carry out of bit 254 */
  g1aff *pts = (g1aff *)malloc(n * sizeof(g1aff));
  int32_t *digits = (int32_t *)malloc(n * (size_t)W * sizeof(int32_t));
  for (size_t i = 0; i < n; i++) {
    const uint8_t *b = bases + 96 * i;
    pts[i].inf = (b[0] & 0x40) != 0;
    if (!pts[i].inf) { fp_from_be(&pts[i].x, b); fp_from_be(&pts[i].y, b + 48); }
    const uint8_t *k = scalars + 32 * i;
    int carry = 0;
    for (int w = 0; w < W; w++) {
      int64_t v = carry;
      for (int t = 0; t < c; t++) {
        int bit = w * c + t;
        if (bit < 256) v += (int64_t)((k[bit >> 3] >> (bit & 7)) & 1) << t;
      }
      carry = 0;
      if (v > ((int64_t)1 << (c - 1))) { v -= (int64_t)1 << c; carry = 1; }
      digits[i * W + w] = (int32_t)v;
    }
  }
  /* W windows x `chunks` point ranges, so that every thread has work when threads > W */
  int chunks = (threads + W - 1) / W;
  if ((size_t)chunks > n) chunks = (int)n;
  const int njobs = W * chunks;
  pjob *jobs = (pjob *)calloc(njobs, sizeof(pjob));
  for (int w = 0; w < W; w++)
    for (int ch = 0; ch < chunks; ch++) {
      pjob *j = &jobs[w * chunks + ch];
      j->pts = pts; j->digits = digits; j->w = w; j->W = W; j->c = c;
      j->lo = n * (size_t)ch / chunks; j->hi = n * (size_t)(ch + 1) / chunks;
    }
  if (threads > njobs) threads = njobs;
  pthread_t *th = (pthread_t *)calloc(threads, sizeof(pthread_t));
  pworker *wk = (pworker *)calloc(threads, sizeof(pworker));
  for (int t = 0; t < threads; t++) {
    wk[t].jobs = jobs; wk[t].first = t; wk[t].step = threads; wk[t].count = njobs;
    if (threads == 1) pippenger_worker(&wk[t]); else pthread_create(&th[t], NULL, pippenger_worker, &wk[t]);
  }
  if (threads > 1) for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  /* Horner over the windows, most significant first */
  g1 acc; g1_set_inf(&acc);
  for (int w = W - 1; w >= 0; w--) {
    for (int t = 0; t < c; t++) g1_dbl(&acc, &acc);
    for (int ch = 0; ch < chunks; ch++) g1_add(&acc, &acc, &jobs[w * chunks + ch].acc);
  }
  g1_to_raw(out, &acc);
  free(jobs); free(th); free(wk); free(pts); free(digits);
  return 0;
}

/* out[i] = scalars[i] * generator-like base given in `base` (test helper) */
int zkoracle_g1_mul(const uint8_t *base, const uint8_t *scalar, uint8_t out[96]) {
  g1 p, r; g1_from_raw(&p, base); g1_mul(&r, &p, scalar); g1_to_raw(out, &r); return 0;
}
