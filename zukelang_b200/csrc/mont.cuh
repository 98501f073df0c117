// Montgomery prime-field arithmetic on 32-bit limbs, carries held in registers.
//
// One template serves Fp (N = 12, 381-bit) and Fr (N = 8, 255-bit).  The
// parameter class P supplies N and constexpr limb accessors (mod, r2, one, n0)
// so that after full unrolling every modulus limb is an immediate operand of
// its IMAD.
//
// Replaces the Fp/Fr arithmetic the reference reaches through the bls12-381
// package (call sites /root/reference/src/lib/zk/curve.ml:121-171).
//
// Multiplication is CIOS Montgomery with the even/odd accumulator split: the
// 64-bit products a[j]*b_i for even j tile limbs (j, j+1) without overlap, and
// likewise for odd j one limb higher, so each half is a single uninterrupted
// mad.lo.cc / madc.hi.cc carry chain (IMAD pipe) with no carry fix-ups between
// rows.  4N+1 IMADs per row, N rows.
//
// Provenance: the even/odd CIOS split and the row-helper vocabulary (mul_n, cmad_n,
// madc_n_rshift, mad_row, final_sub) follow the scheme publicly documented for GPU Montgomery
// arithmetic in the sppark / yrrid `mont_t` family (Apache-2.0); nothing is copied from those
// sources — the code below was written against the PTX ISA.  mul2 (two interleaved products),
// sqr_row (dedicated square) and the constexpr-immediate parameter classes are this project's own.
#pragma once
#include "ptx.cuh"

template <class P>
struct alignas(16) Mont {
  static constexpr int N = P::N;
  uint32_t v[N];

  // ---- construction -------------------------------------------------------
  static ZK_HD Mont zero() {
    Mont r;
    ZK_UNROLL for (int i = 0; i < N; i++) r.v[i] = 0;
    return r;
  }
  static ZK_HD Mont one() {  // R mod p
    Mont r;
    ZK_UNROLL for (int i = 0; i < N; i++) r.v[i] = P::one(i);
    return r;
  }
  static ZK_HD Mont r2() {
    Mont r;
    ZK_UNROLL for (int i = 0; i < N; i++) r.v[i] = P::r2(i);
    return r;
  }

  ZK_HD bool is_zero() const {
    uint32_t o = 0;
    ZK_UNROLL for (int i = 0; i < N; i++) o |= v[i];
    return o == 0;
  }
  ZK_HD bool operator==(const Mont& b) const {
    uint32_t o = 0;
    ZK_UNROLL for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  ZK_HD bool operator!=(const Mont& b) const { return !(*this == b); }

  // ---- add / sub ----------------------------------------------------------
  // r = (a >= p) ? a - p : a, for a < 2p given with an extra carry bit `hi`
  static ZK_HD void final_sub(Mont& a, uint32_t hi) {
    uint32_t t[N];
    t[0] = ptx::sub_cc(a.v[0], P::mod(0));
    ZK_UNROLL for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(a.v[i], P::mod(i));
    uint32_t borrow = ptx::subc(hi, 0);  // 0 if a - p >= 0, else 0xffffffff
    ZK_UNROLL for (int i = 0; i < N; i++) a.v[i] = borrow ? a.v[i] : t[i];
  }

  friend ZK_HD Mont operator+(const Mont& a, const Mont& b) {
    Mont r;
    r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
    ZK_UNROLL for (int i = 1; i < N; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
    uint32_t hi = ptx::addc(0, 0);
    final_sub(r, hi);
    return r;
  }
  friend ZK_HD Mont operator-(const Mont& a, const Mont& b) {
    Mont r;
    r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
    ZK_UNROLL for (int i = 1; i < N; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
    uint32_t borrow = ptx::subc(0, 0);  // 0 or 0xffffffff
    uint32_t t[N];
    t[0] = ptx::add_cc(r.v[0], P::mod(0) & borrow);
    ZK_UNROLL for (int i = 1; i < N; i++) t[i] = ptx::addc_cc(r.v[i], P::mod(i) & borrow);
    ZK_UNROLL for (int i = 0; i < N; i++) r.v[i] = t[i];
    return r;
  }
  ZK_HD Mont neg() const {
    Mont r;
    if (is_zero()) return *this;
    r.v[0] = ptx::sub_cc(P::mod(0), v[0]);
    ZK_UNROLL for (int i = 1; i < N; i++) r.v[i] = ptx::subc_cc(P::mod(i), v[i]);
    return r;
  }
  ZK_HD Mont dbl() const { return *this + *this; }

  // ---- Montgomery product ---------------------------------------------------
  // acc[j], acc[j+1] = lo, hi of a[j] * bi for j = 0, 2, .. n-2
  template <int n>
  static ZK_HD void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
    ZK_UNROLL for (int j = 0; j < n; j += 2) {
      acc[j] = ptx::mul_lo(a[j], bi);
      acc[j + 1] = ptx::mul_hi(a[j], bi);
    }
  }
  // acc[0..n) += (a[0], a[2], ..) * bi as one carry chain; carry-out left in CC
  template <int n>
  static ZK_HD void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
    acc[0] = ptx::mad_lo_cc(a[0], bi, acc[0]);
    acc[1] = ptx::madc_hi_cc(a[0], bi, acc[1]);
    ZK_UNROLL for (int j = 2; j < n; j += 2) {
      acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j]);
      acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 1]);
    }
  }
  // same with the modulus limbs (immediates) starting at limb `off`
  template <int n, int off>
  static ZK_HD void cmad_mod(uint32_t* acc, uint32_t mi) {
    acc[0] = ptx::mad_lo_cc(P::mod(off), mi, acc[0]);
    acc[1] = ptx::madc_hi_cc(P::mod(off), mi, acc[1]);
    ZK_UNROLL for (int j = 2; j < n; j += 2) {
      acc[j] = ptx::madc_lo_cc(P::mod(off + j), mi, acc[j]);
      acc[j + 1] = ptx::madc_hi_cc(P::mod(off + j), mi, acc[j + 1]);
    }
  }
  // odd[] = (odd[] >> 64) + (a[0], a[2], ..) * bi, consuming the incoming CC
  static ZK_HD void madc_n_rshift(uint32_t* odd, const uint32_t* a, uint32_t bi) {
    ZK_UNROLL for (int j = 0; j < N - 2; j += 2) {
      odd[j] = ptx::madc_lo_cc(a[j], bi, odd[j + 2]);
      odd[j + 1] = ptx::madc_hi_cc(a[j], bi, odd[j + 3]);
    }
    odd[N - 2] = ptx::madc_lo_cc(a[N - 2], bi, 0);
    odd[N - 1] = ptx::madc_hi(a[N - 2], bi, 0);
  }
  // T += m * p with m chosen so that the low limb cancels
  static ZK_HD void reduce_row(uint32_t* even, uint32_t* odd) {
    uint32_t mi = even[0] * P::n0();
    cmad_mod<N, 1>(odd, mi);
    cmad_mod<N, 0>(even, mi);
    odd[N - 1] = ptx::addc(odd[N - 1], 0);
  }
  // one CIOS row: (even, odd) <- ((even, odd) >> 32) + a * bi, then reduce.
  // On entry `odd` is the array that was even-aligned in the previous row.
  static ZK_HD void mad_row(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
    even[0] = ptx::add_cc(even[0], odd[1]);
    madc_n_rshift(odd, a + 1, bi);
    cmad_n<N>(even, a, bi);
    odd[N - 1] = ptx::addc(odd[N - 1], 0);
    reduce_row(even, odd);
  }

  friend ZK_HD Mont operator*(const Mont& a, const Mont& b) {
    uint32_t e[N], o[N];
    mul_n<N>(e, a.v, b.v[0]);
    mul_n<N>(o, a.v + 1, b.v[0]);
    reduce_row(e, o);
    ZK_UNROLL for (int i = 1; i < N; i += 2) {
      mad_row(o, e, a.v, b.v[i]);
      if (i + 1 < N) mad_row(e, o, a.v, b.v[i + 1]);
    }
    // value = (e >> 32) + o
    Mont r;
    r.v[0] = ptx::add_cc(e[0], o[1]);
    ZK_UNROLL for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(e[i], o[i + 1]);
    r.v[N - 1] = ptx::addc(e[N - 1], 0);
    final_sub(r, 0);
    return r;
  }
  // Two independent products with their CIOS rows interleaved: r1 = a b, r2 = c d.  Each product
  // is a serial chain of rows (row i+1 needs row i's accumulator); issuing the rows of two products
  // alternately gives the scheduler two independent chains to overlap, which hides the dependent-
  // issue latency of IMAD.WIDE.X at low occupancy.
  static ZK_HD void mul2(const Mont& a, const Mont& b, const Mont& c, const Mont& d, Mont& r1, Mont& r2) {
    uint32_t e1[N], o1[N], e2[N], o2[N];
    mul_n<N>(e1, a.v, b.v[0]);
    mul_n<N>(e2, c.v, d.v[0]);
    mul_n<N>(o1, a.v + 1, b.v[0]);
    mul_n<N>(o2, c.v + 1, d.v[0]);
    reduce_row(e1, o1);
    reduce_row(e2, o2);
    ZK_UNROLL for (int i = 1; i < N; i += 2) {
      mad_row(o1, e1, a.v, b.v[i]);
      mad_row(o2, e2, c.v, d.v[i]);
      if (i + 1 < N) {
        mad_row(e1, o1, a.v, b.v[i + 1]);
        mad_row(e2, o2, c.v, d.v[i + 1]);
      }
    }
    r1.v[0] = ptx::add_cc(e1[0], o1[1]);
    ZK_UNROLL for (int i = 1; i < N - 1; i++) r1.v[i] = ptx::addc_cc(e1[i], o1[i + 1]);
    r1.v[N - 1] = ptx::addc(e1[N - 1], 0);
    r2.v[0] = ptx::add_cc(e2[0], o2[1]);
    ZK_UNROLL for (int i = 1; i < N - 1; i++) r2.v[i] = ptx::addc_cc(e2[i], o2[i + 1]);
    r2.v[N - 1] = ptx::addc(e2[N - 1], 0);
    final_sub(r1, 0);
    final_sub(r2, 0);
  }

  // mul2 with the row loop ROLLED (two rows of each product per iteration): ~8 KB of code instead of
  // ~22 KB, and a loop body that fits the 6 KB L0 instruction cache.  For the latency-bound callers
  // (bucket reduction, fix-ups, fixed-base ladders, affine conversion): their fully inlined formulas
  // were 200 KB per kernel, far beyond the 32 KB L1.5 instruction cache, and ncu shows them waiting
  // for instructions (`no_instruction` 3.2 warps per issue) more than for operands.  The row operands
  // b_i, d_i are taken from rotating register copies, so no index is dynamic.
  static ZK_HD void mul2_rolled(const Mont& a, const Mont& b, const Mont& c, const Mont& d, Mont& r1, Mont& r2) {
    uint32_t e1[N], o1[N], e2[N], o2[N], bb[N], dd[N];
    ZK_UNROLL for (int j = 0; j < N; j++) { bb[j] = b.v[j]; dd[j] = d.v[j]; }
    mul_n<N>(e1, a.v, bb[0]);
    mul_n<N>(e2, c.v, dd[0]);
    mul_n<N>(o1, a.v + 1, bb[0]);
    mul_n<N>(o2, c.v + 1, dd[0]);
    reduce_row(e1, o1);
    reduce_row(e2, o2);
    ZK_NOUNROLL for (int it = 0; it < (N - 1) / 2; it++) {   // rows (1, 2), (3, 4), ...
      mad_row(o1, e1, a.v, bb[1]);
      mad_row(o2, e2, c.v, dd[1]);
      mad_row(e1, o1, a.v, bb[2]);
      mad_row(e2, o2, c.v, dd[2]);
      ZK_UNROLL for (int j = 1; j + 2 < N; j++) { bb[j] = bb[j + 2]; dd[j] = dd[j + 2]; }
    }
    if ((N - 1) & 1) {                                       // the last row when N is even
      mad_row(o1, e1, a.v, bb[1]);
      mad_row(o2, e2, c.v, dd[1]);
    }
    r1.v[0] = ptx::add_cc(e1[0], o1[1]);
    ZK_UNROLL for (int i = 1; i < N - 1; i++) r1.v[i] = ptx::addc_cc(e1[i], o1[i + 1]);
    r1.v[N - 1] = ptx::addc(e1[N - 1], 0);
    r2.v[0] = ptx::add_cc(e2[0], o2[1]);
    ZK_UNROLL for (int i = 1; i < N - 1; i++) r2.v[i] = ptx::addc_cc(e2[i], o2[i + 1]);
    r2.v[N - 1] = ptx::addc(e2[N - 1], 0);
    final_sub(r1, 0);
    final_sub(r2, 0);
  }

  // ---- Montgomery square ---------------------------------------------------------
  // a^2 = sum_i a_i * e^(i) * 2^(32 i) with e^(i) = (0, .., 0, a_i, 2 a_{i+1}, .., 2 a_{N-1}): row i of
  // the CIOS loop only multiplies the limbs j >= i, 78 partial products for N = 12 instead of 144
  // (needs 2a + p < 2^(32 N): true for Fp; Fr, one bit short, keeps the plain product).  Limbs of the row
  // operand below `start` are zero: the odd-aligned chain just shifts there (carry adds on the ALU
  // pipe) and the even-aligned chain starts at the first non-zero limb.
  static ZK_HD void sqr_row(uint32_t* even, uint32_t* odd, const uint32_t* e, uint32_t bi, int start) {
    even[0] = ptx::add_cc(even[0], odd[1]);
    ZK_UNROLL for (int j = 0; j < N - 2; j += 2) {
      if (j + 1 >= start) {
        odd[j] = ptx::madc_lo_cc(e[j + 1], bi, odd[j + 2]);
        odd[j + 1] = ptx::madc_hi_cc(e[j + 1], bi, odd[j + 3]);
      } else {
        odd[j] = ptx::addc_cc(odd[j + 2], 0);
        odd[j + 1] = ptx::addc_cc(odd[j + 3], 0);
      }
    }
    odd[N - 2] = ptx::madc_lo_cc(e[N - 1], bi, 0);
    odd[N - 1] = ptx::madc_hi(e[N - 1], bi, 0);
    const int fe = start + (start & 1);  // first even limb >= start
    if (fe < N) {
      even[fe] = ptx::mad_lo_cc(e[fe], bi, even[fe]);
      even[fe + 1] = ptx::madc_hi_cc(e[fe], bi, even[fe + 1]);
      ZK_UNROLL for (int j = 0; j < N; j += 2) {
        if (j > fe) {
          even[j] = ptx::madc_lo_cc(e[j], bi, even[j]);
          even[j + 1] = ptx::madc_hi_cc(e[j], bi, even[j + 1]);
        }
      }
      odd[N - 1] = ptx::addc(odd[N - 1], 0);
    }
    reduce_row(even, odd);
  }
  ZK_HD Mont sqr() const {
    // (valid only when 2a + p fits the N-limb accumulator: Fp has 3 spare bits, Fr has one)
    if (!P::FAST_SQR) return *this * *this;
    // d = 2a limb-wise with the carries between limbs; row i uses e = (.., a_i, a_{i+1} << 1, d_{i+2}, ..):
    // the bit that 2 a_i pushes into limb i+1 belongs to the diagonal term, not to row i's cross terms
    uint32_t d[N], e[N], ev[N], od[N];
    d[0] = v[0] << 1;
    ZK_UNROLL for (int i = 1; i < N; i++) d[i] = (v[i] << 1) | (v[i - 1] >> 31);
    ZK_UNROLL for (int j = 0; j < N; j++) e[j] = d[j];
    e[0] = v[0];
    e[1] = v[1] << 1;
    mul_n<N>(ev, e, v[0]);
    mul_n<N>(od, e + 1, v[0]);
    reduce_row(ev, od);
    ZK_UNROLL for (int i = 1; i < N; i += 2) {
      e[i] = v[i];
      if (i + 1 < N) e[i + 1] = v[i + 1] << 1;
      sqr_row(od, ev, e, v[i], i);
      if (i + 1 < N) {
        e[i + 1] = v[i + 1];
        if (i + 2 < N) e[i + 2] = v[i + 2] << 1;
        sqr_row(ev, od, e, v[i + 1], i + 1);
        if (i + 2 < N) e[i + 2] = d[i + 2];   // restored: the next row patches it again as its own a_i
      }
    }
    Mont r;
    r.v[0] = ptx::add_cc(ev[0], od[1]);
    ZK_UNROLL for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(ev[i], od[i + 1]);
    r.v[N - 1] = ptx::addc(ev[N - 1], 0);
    final_sub(r, 0);
    return r;
  }
  // out-of-line product for code that is not throughput critical (keeps kernels small)
  static ZK_NI Mont mul_call(const Mont& a, const Mont& b) { return a * b; }

  // ---- conversions -----------------------------------------------------------
  ZK_HD Mont to_mont() const { return *this * r2(); }
  ZK_HD Mont from_mont() const {
    Mont u = zero();
    u.v[0] = 1;
    return *this * u;
  }
  // canonical-integer comparison helpers (operate on from_mont'ed values)
  static ZK_HD bool geq_raw(const uint32_t* a, const uint32_t* b) {  // a >= b
    ptx::sub_cc(a[0], b[0]);
    ZK_UNROLL for (int i = 1; i < N; i++) ptx::subc_cc(a[i], b[i]);
    return ptx::subc(0, 0) == 0;
  }
  ZK_HD bool is_canonical_raw() const {  // value < p, for an unreduced limb array
    uint32_t m[N];
    ZK_UNROLL for (int i = 0; i < N; i++) m[i] = P::mod(i);
    return !geq_raw(v, m);
  }

  // a^e for a little-endian limb exponent (square-and-multiply, MSB first)
  template <int EN>
  ZK_NI Mont pow_limbs(const uint32_t (&e)[EN]) const {
    Mont acc = one();
    bool started = false;
    for (int i = EN - 1; i >= 0; i--) {
      for (int bit = 31; bit >= 0; bit--) {
        if (started) acc = mul_call(acc, acc);
        if ((e[i] >> bit) & 1) {
          acc = started ? mul_call(acc, *this) : *this;
          started = true;
        }
      }
    }
    return acc;
  }
  // ---- inversion ------------------------------------------------------------------
  // Binary extended Euclid on the raw limbs (adds, subtractions and shifts only: ~2 log2 p
  // iterations of IADD3 / SHF chains instead of ~1.5 log2 p Montgomery products).  Input and
  // output are Montgomery residues; inverse of 0 is 0.
  static ZK_HD void shr1(uint32_t* a, uint32_t top) {
    ZK_UNROLL for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[N - 1] = (a[N - 1] >> 1) | (top << 31);
  }
  static ZK_HD void halve_mod(uint32_t* x) {  // x <- x / 2 mod p
    uint32_t carry = 0;
    if (x[0] & 1) {
      x[0] = ptx::add_cc(x[0], P::mod(0));
      ZK_UNROLL for (int i = 1; i < N; i++) x[i] = ptx::addc_cc(x[i], P::mod(i));
      carry = ptx::addc(0, 0);
    }
    shr1(x, carry);
  }
  static ZK_HD void sub_raw(uint32_t* a, const uint32_t* b) {  // a -= b (a >= b)
    a[0] = ptx::sub_cc(a[0], b[0]);
    ZK_UNROLL for (int i = 1; i < N; i++) a[i] = ptx::subc_cc(a[i], b[i]);
  }
  static ZK_HD void sub_mod_raw(uint32_t* a, const uint32_t* b) {  // a <- a - b mod p
    a[0] = ptx::sub_cc(a[0], b[0]);
    ZK_UNROLL for (int i = 1; i < N; i++) a[i] = ptx::subc_cc(a[i], b[i]);
    uint32_t borrow = ptx::subc(0, 0);
    if (borrow) {
      a[0] = ptx::add_cc(a[0], P::mod(0));
      ZK_UNROLL for (int i = 1; i < N; i++) a[i] = ptx::addc_cc(a[i], P::mod(i));
    }
  }
  static ZK_HD bool is_one_raw(const uint32_t* a) {
    uint32_t o = a[0] ^ 1u;
    ZK_UNROLL for (int i = 1; i < N; i++) o |= a[i];
    return o == 0;
  }
  ZK_NI Mont inverse() const {
    if (is_zero()) return *this;
    uint32_t u[N], w[N], x1[N], x2[N];
    ZK_UNROLL for (int i = 0; i < N; i++) { u[i] = v[i]; w[i] = P::mod(i); x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    while (!is_one_raw(u) && !is_one_raw(w)) {
      while (!(u[0] & 1)) { shr1(u, 0); halve_mod(x1); }
      while (!(w[0] & 1)) { shr1(w, 0); halve_mod(x2); }
      if (geq_raw(u, w)) { sub_raw(u, w); sub_mod_raw(x1, x2); }
      else { sub_raw(w, u); sub_mod_raw(x2, x1); }
    }
    Mont r, r3;
    const bool pick = is_one_raw(u);
    ZK_UNROLL for (int i = 0; i < N; i++) { r.v[i] = pick ? x1[i] : x2[i]; r3.v[i] = P::r3(i); }
    return mul_call(r, r3);  // (aR)^-1 * R^3 / R = a^-1 R
  }
  // multiplicative inverse by Fermat (a^(p-2)); kept as the cross-check of inverse()
  ZK_NI Mont inverse_fermat() const {
    uint32_t e[N];
    e[0] = ptx::sub_cc(P::mod(0), 2);
    ZK_UNROLL for (int i = 1; i < N; i++) e[i] = ptx::subc_cc(P::mod(i), 0);
    return pow_limbs<N>(e);
  }
};
