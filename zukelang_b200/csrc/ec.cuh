// Fp2 tower and short-Weierstrass (a = 0) point arithmetic in XYZZ coordinates,
// generic over the base field so G1 (Fp) and G2 (Fp2) share one implementation.
//
// Replaces Bls12_381.G1/G2.{add, mul, negate} as wrapped by ExtendG at
// /root/reference/src/lib/zk/curve.ml:159-191.
//
// XYZZ: x = X/ZZ, y = Y/ZZZ with ZZ^3 = ZZZ^2; the identity is ZZ = 0.
// Mixed addition (affine addend) costs 8M + 2S, general addition 12M + 2S,
// doubling 6M + 4S-ish (EFD: madd-2008-s, add-2008-s, dbl-2008-s-1, mdbl-2008-s-1).
#pragma once
#include "mont.cuh"
#include "params.cuh"

typedef Mont<FpParams> Fp;
typedef Mont<FrParams> Fr;

// The one out-of-line Fp product: arguments and result stay in registers under the device ABI.
// Used by Fp2 and FpCall, where fully inlined formulas would be tens of kilobytes of code.
struct FpPair { Fp lo, hi; };
#ifndef ZK_HOST_SIM
static __device__ __noinline__ Fp fp_mul_outofline(Fp a, Fp b) { return a * b; }
// two independent products with interleaved rows (Mont::mul2), out of line
static __device__ __noinline__ FpPair fp_mul2_outofline(Fp a, Fp b, Fp c, Fp d) {
  FpPair r;
  Fp::mul2(a, b, c, d, r.lo, r.hi);
  return r;
}
// The same pair of products with the row loop rolled (Mont::mul2_rolled): ~8 KB of code.  What the
// out-of-line point formulas below call — the latency-bound kernels spend their time waiting for
// instructions, not for the multiplier (see mont.cuh).
static __device__ __noinline__ FpPair fp_mul2_compact(Fp a, Fp b, Fp c, Fp d) {
  FpPair r;
  Fp::mul2_rolled(a, b, c, d, r.lo, r.hi);
  return r;
}
#else
static inline FpPair fp_mul2_compact(Fp a, Fp b, Fp c, Fp d) {
  FpPair r;
  Fp::mul2_rolled(a, b, c, d, r.lo, r.hi);
  return r;
}
static inline Fp fp_mul_outofline(Fp a, Fp b) { return a * b; }
static inline FpPair fp_mul2_outofline(Fp a, Fp b, Fp c, Fp d) {
  FpPair r;
  Fp::mul2(a, b, c, d, r.lo, r.hi);
  return r;
}
#endif

// -------------------------------------------------------------------------
// Fp2 = Fp[u] / (u^2 + 1)
// -------------------------------------------------------------------------
struct Fp2 {
  Fp c0, c1;
  static ZK_HD Fp2 zero() { return Fp2{Fp::zero(), Fp::zero()}; }
  static ZK_HD Fp2 one() { return Fp2{Fp::one(), Fp::zero()}; }
  ZK_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  ZK_HD bool operator==(const Fp2& b) const { return c0 == b.c0 && c1 == b.c1; }
  ZK_HD bool operator!=(const Fp2& b) const { return !(*this == b); }
  friend ZK_HD Fp2 operator+(const Fp2& a, const Fp2& b) { return Fp2{a.c0 + b.c0, a.c1 + b.c1}; }
  friend ZK_HD Fp2 operator-(const Fp2& a, const Fp2& b) { return Fp2{a.c0 - b.c0, a.c1 - b.c1}; }
  ZK_HD Fp2 neg() const { return Fp2{c0.neg(), c1.neg()}; }
  ZK_HD Fp2 dbl() const { return Fp2{c0.dbl(), c1.dbl()}; }
  // Karatsuba: 3 Fp products
  friend ZK_HD Fp2 operator*(const Fp2& a, const Fp2& b) {
    FpPair t = fp_mul2_outofline(a.c0, b.c0, a.c1, b.c1);
    Fp t2 = fp_mul_outofline(a.c0 + a.c1, b.c0 + b.c1);
    return Fp2{t.lo - t.hi, t2 - t.lo - t.hi};
  }
  // Two independent Fp2 products: their six Fp products go out as three interleaved pairs
  // (Mont::mul2), so that every carry chain has an independent neighbour to overlap with.
  static ZK_HD void mul2(const Fp2& a, const Fp2& b, const Fp2& c, const Fp2& d, Fp2& r1, Fp2& r2) {
    FpPair p0 = fp_mul2_outofline(a.c0, b.c0, c.c0, d.c0);
    FpPair p1 = fp_mul2_outofline(a.c1, b.c1, c.c1, d.c1);
    FpPair pm = fp_mul2_outofline(a.c0 + a.c1, b.c0 + b.c1, c.c0 + c.c1, d.c0 + d.c1);
    r1 = Fp2{p0.lo - p1.lo, pm.lo - p0.lo - p1.lo};
    r2 = Fp2{p0.hi - p1.hi, pm.hi - p0.hi - p1.hi};
  }
  // (c0 + c1 u)^2 = (c0 + c1)(c0 - c1) + 2 c0 c1 u : 2 Fp products
  ZK_HD Fp2 sqr() const {
    Fp s = c0 + c1;
    Fp d = c0 - c1;
    FpPair t = fp_mul2_outofline(c0, c1, s, d);
    return Fp2{t.hi, t.lo.dbl()};
  }
  ZK_NI Fp2 inverse() const {
    Fp n = Fp::mul_call(c0, c0) + Fp::mul_call(c1, c1);
    Fp ni = n.inverse();
    return Fp2{Fp::mul_call(c0, ni), Fp::mul_call(c1, ni).neg()};
  }
};

// Fp with the product routed through ONE out-of-line routine (arguments and result stay in
// registers under the device ABI).  Same representation as Fp; used where the fully inlined
// point formulas would overflow the instruction cache.
struct FpCall {
  Fp f;
  static ZK_HD FpCall zero() { return FpCall{Fp::zero()}; }
  static ZK_HD FpCall one() { return FpCall{Fp::one()}; }
  ZK_HD bool is_zero() const { return f.is_zero(); }
  ZK_HD bool operator==(const FpCall& b) const { return f == b.f; }
  friend ZK_HD FpCall operator+(const FpCall& a, const FpCall& b) { return FpCall{a.f + b.f}; }
  friend ZK_HD FpCall operator-(const FpCall& a, const FpCall& b) { return FpCall{a.f - b.f}; }
#ifndef ZK_HOST_SIM
  friend ZK_HD FpCall operator*(const FpCall& a, const FpCall& b) { return FpCall{fp_mul_outofline(a.f, b.f)}; }
  ZK_HD FpCall sqr() const { return FpCall{fp_mul_outofline(f, f)}; }
#else
  friend ZK_HD FpCall operator*(const FpCall& a, const FpCall& b) { return FpCall{a.f * b.f}; }
  ZK_HD FpCall sqr() const { return FpCall{f * f}; }
#endif
  static ZK_HD void mul2(const FpCall& a, const FpCall& b, const FpCall& c, const FpCall& d, FpCall& r1, FpCall& r2) {
    FpPair p = fp_mul2_outofline(a.f, b.f, c.f, d.f);
    r1.f = p.lo;
    r2.f = p.hi;
  }
  ZK_HD FpCall dbl() const { return FpCall{f.dbl()}; }
  ZK_HD FpCall neg() const { return FpCall{f.neg()}; }
  ZK_NI FpCall inverse() const { return FpCall{f.inverse()}; }
};

// -------------------------------------------------------------------------
// Products of the OUT-OF-LINE point formulas (add, dbl, to_affine: bucket reduction, fix-ups,
// window combine, ladders).  Every product goes through the one compact routine above, in
// independent pairs, so that a whole latency-bound kernel is a few tens of KB of code: r1 = a b,
// r2 = c d; lat_sqr2 = two independent squares.
// -------------------------------------------------------------------------
ZK_HD void lat_mul2(const Fp& a, const Fp& b, const Fp& c, const Fp& d, Fp& r1, Fp& r2) {
  FpPair p = fp_mul2_compact(a, b, c, d);
  r1 = p.lo;
  r2 = p.hi;
}
ZK_HD void lat_mul2(const FpCall& a, const FpCall& b, const FpCall& c, const FpCall& d, FpCall& r1, FpCall& r2) {
  FpPair p = fp_mul2_compact(a.f, b.f, c.f, d.f);
  r1.f = p.lo;
  r2.f = p.hi;
}
// (out of line themselves: the ten Fp additions around the three calls are 6 KB of code, and a G2
// addition makes seven of these)
static ZK_NI void lat_mul2(const Fp2& a, const Fp2& b, const Fp2& c, const Fp2& d, Fp2& r1, Fp2& r2) {
  FpPair p0 = fp_mul2_compact(a.c0, b.c0, c.c0, d.c0);
  FpPair p1 = fp_mul2_compact(a.c1, b.c1, c.c1, d.c1);
  FpPair pm = fp_mul2_compact(a.c0 + a.c1, b.c0 + b.c1, c.c0 + c.c1, d.c0 + d.c1);
  r1 = Fp2{p0.lo - p1.lo, pm.lo - p0.lo - p1.lo};
  r2 = Fp2{p0.hi - p1.hi, pm.hi - p0.hi - p1.hi};
}
ZK_HD void lat_sqr2(const Fp& a, const Fp& b, Fp& ra, Fp& rb) { lat_mul2(a, a, b, b, ra, rb); }
ZK_HD void lat_sqr2(const FpCall& a, const FpCall& b, FpCall& ra, FpCall& rb) { lat_mul2(a, a, b, b, ra, rb); }
static ZK_NI void lat_sqr2(const Fp2& a, const Fp2& b, Fp2& ra, Fp2& rb) {
  // (c0 + c1 u)^2 = (c0 + c1)(c0 - c1) + 2 c0 c1 u
  FpPair x = fp_mul2_compact(a.c0 + a.c1, a.c0 - a.c1, a.c0, a.c1);
  FpPair y = fp_mul2_compact(b.c0 + b.c1, b.c0 - b.c1, b.c0, b.c1);
  ra = Fp2{x.lo, x.hi.dbl()};
  rb = Fp2{y.lo, y.hi.dbl()};
}
template <class F>
ZK_HD F lat_mul(const F& a, const F& b) {
  F r, unused;
  lat_mul2(a, b, a, b, r, unused);
  return r;
}

// -------------------------------------------------------------------------
// points
// -------------------------------------------------------------------------
template <class F>
struct Affine {
  F x, y;  // identity is encoded as (0, 0), which is not on either curve (b != 0)
  ZK_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  static ZK_HD Affine inf() { return Affine{F::zero(), F::zero()}; }
  ZK_HD Affine neg() const { return Affine{x, y.neg()}; }
};

template <class F>
struct XYZZ {
  F X, Y, ZZ, ZZZ;

  static ZK_HD XYZZ inf() { return XYZZ{F::zero(), F::zero(), F::zero(), F::zero()}; }
  ZK_HD bool is_inf() const { return ZZ.is_zero(); }
  static ZK_HD XYZZ from_affine(const Affine<F>& p) {
    if (p.is_inf()) return inf();
    return XYZZ{p.x, p.y, F::one(), F::one()};
  }
  ZK_HD XYZZ neg() const { return XYZZ{X, Y.neg(), ZZ, ZZZ}; }

  // The out-of-line formulas below (latency-bound callers: bucket reduction, fix-ups, window
  // combine, fixed-base ladders) issue their independent products as interleaved pairs through ONE
  // compact out-of-line routine (lat_mul2): a lone warp overlaps two carry chains instead of waiting
  // on one, and the kernels stay small enough for the instruction caches.

  // 2 * (affine p)   — mdbl-2008-s-1
  static ZK_NI XYZZ dbl_affine(const Affine<F>& p) {
    if (p.is_inf() || p.y.is_zero()) return inf();
    F U = p.y.dbl();
    F V, xx;
    lat_sqr2(U, p.x, V, xx);
    F W, S;
    lat_mul2(U, V, p.x, V, W, S);
    F M = xx.dbl() + xx;
    XYZZ r;
    F MM, t2;
    lat_mul2(M, M, W, p.y, MM, t2);
    r.X = MM - S.dbl();
    F t1 = lat_mul(M, S - r.X);
    r.Y = t1 - t2;
    r.ZZ = V;
    r.ZZZ = W;
    return r;
  }

  // dbl-2008-s-1
  ZK_NI XYZZ dbl() const {
    if (is_inf() || Y.is_zero()) return inf();
    F U = Y.dbl();
    F V, xx;
    lat_sqr2(U, X, V, xx);
    F W, S;
    lat_mul2(U, V, X, V, W, S);
    F M = xx.dbl() + xx;
    XYZZ r;
    F MM, t2;
    lat_mul2(M, M, W, Y, MM, t2);
    r.X = MM - S.dbl();
    F t1;
    lat_mul2(M, S - r.X, V, ZZ, t1, r.ZZ);
    r.Y = t1 - t2;
    r.ZZZ = lat_mul(W, ZZZ);
    return r;
  }

  // madd with the independent products issued in interleaved pairs (F::mul2): what k_accumulate runs
  ZK_HD void madd_paired(const Affine<F>& p) {
    if (p.is_inf()) return;
    if (is_inf()) {
      X = p.x; Y = p.y; ZZ = F::one(); ZZZ = F::one();
      return;
    }
    F U2, S2;
    F::mul2(p.x, ZZ, p.y, ZZZ, U2, S2);
    F Pd = U2 - X;
    F Rd = S2 - Y;
    if (Pd.is_zero()) {
      if (Rd.is_zero()) *this = dbl_affine(p);
      else *this = inf();
      return;
    }
    F PP = Pd.sqr();
    F RR = Rd.sqr();
    F PPP, Q;
    F::mul2(Pd, PP, X, PP, PPP, Q);
    F X3 = RR - PPP - Q.dbl();
    F t1, t2;
    F::mul2(Rd, Q - X3, Y, PPP, t1, t2);
    F::mul2(ZZ, PP, ZZZ, PPP, ZZ, ZZZ);
    Y = t1 - t2;
    X = X3;
  }

  // this += affine p   — madd-2008-s, with the exceptional cases handled
  ZK_HD void madd(const Affine<F>& p) {
    if (p.is_inf()) return;
    if (is_inf()) {
      X = p.x; Y = p.y; ZZ = F::one(); ZZZ = F::one();
      return;
    }
    F U2 = p.x * ZZ;
    F S2 = p.y * ZZZ;
    F Pd = U2 - X;
    F Rd = S2 - Y;
    if (Pd.is_zero()) {
      if (Rd.is_zero()) *this = dbl_affine(p);
      else *this = inf();
      return;
    }
    F PP = Pd.sqr();
    F PPP = Pd * PP;
    F Q = X * PP;
    F X3 = Rd.sqr() - PPP - Q.dbl();
    Y = Rd * (Q - X3) - Y * PPP;
    X = X3;
    ZZ = ZZ * PP;
    ZZZ = ZZZ * PPP;
  }

  // this += q   — add-2008-s, products in independent pairs (7 pair steps for 12M + 2S)
  ZK_NI void add(const XYZZ& q) {
    if (q.is_inf()) return;
    if (is_inf()) { *this = q; return; }
    F U1, U2, S1, S2;
    lat_mul2(X, q.ZZ, q.X, ZZ, U1, U2);
    lat_mul2(Y, q.ZZZ, q.Y, ZZZ, S1, S2);
    F Pd = U2 - U1;
    F Rd = S2 - S1;
    if (Pd.is_zero()) {
      if (Rd.is_zero()) *this = dbl();
      else *this = inf();
      return;
    }
    F PP, RR;
    lat_sqr2(Pd, Rd, PP, RR);
    F PPP, Q;
    lat_mul2(Pd, PP, U1, PP, PPP, Q);
    F X3 = RR - PPP - Q.dbl();
    F t1, t2, z2, z3;
    lat_mul2(Rd, Q - X3, S1, PPP, t1, t2);
    lat_mul2(ZZ, q.ZZ, ZZZ, q.ZZZ, z2, z3);
    lat_mul2(z2, PP, z3, PPP, ZZ, ZZZ);
    Y = t1 - t2;
    X = X3;
  }

  // affine coordinates (one field inversion); identity -> (0, 0)
  ZK_NI Affine<F> to_affine() const {
    if (is_inf()) return Affine<F>::inf();
    // 1/ZZZ * ZZ = 1/Z  ->  1/ZZ = (1/Z)^2
    F izzz = ZZZ.inverse();
    F iz, y;
    lat_mul2(izzz, ZZ, Y, izzz, iz, y);
    F izz = lat_mul(iz, iz);
    return Affine<F>{lat_mul(X, izz), y};
  }
};

// k * p for a canonical little-endian 8-limb scalar (double-and-add, MSB first)
template <class F>
ZK_NI XYZZ<F> scalar_mul(const XYZZ<F>& p, const uint32_t* k, int nlimbs = 8) {
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int i = nlimbs - 1; i >= 0; i--) {
    for (int bit = 31; bit >= 0; bit--) {
      acc = acc.dbl();
      if ((k[i] >> bit) & 1) acc.add(p);
    }
  }
  return acc;
}

typedef Affine<Fp> G1Affine;
typedef Affine<Fp2> G2Affine;
typedef XYZZ<Fp> G1XYZZ;
typedef XYZZ<Fp2> G2XYZZ;
