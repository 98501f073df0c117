// Verifier-side arithmetic: the Fp12 tower, the optimal-ate pairing and the
// square roots behind point decompression.
//
// Replaces `Bls12_381.Pairing.pairing` as the reference calls it at
// /root/reference/src/groth16/groth16.ml:103,168 and src/pinocchio/pinocchio.ml:269, `GT.( + )`
// (the GT product, curve.ml:212-220) and `G1/G2.of_compressed_bytes_exn` (curve.ml:201,210).
//
// Tower:  Fp2 = Fp[u]/(u^2 + 1),  Fp6 = Fp2[v]/(v^3 - xi) with xi = 1 + u,  Fp12 = Fp6[w]/(w^2 - v).
// The sextic twist sends (x, y) in E'(Fp2) to (x / w^2, y / w^3); line values are scaled by w^3
// (an element of a proper subfield, removed by the final exponentiation), so a line through the
// twisted point T with slope lam, evaluated at P = (xp, yp) in E(Fp), is
//
//     l = (yT - lam xT)  +  (lam xp) v  +  (-yp) v w.
//
// The Miller loop runs over |z| = 0xd201000000010000 in affine coordinates (one Fp inversion per
// step: the binary Euclid routine, adds and shifts only).  The final exponentiation is the easy
// part (p^6 - 1)(p^2 + 1) followed by the hard part in the form
//
//     3 (p^4 - p^2 + 1) / r  =  (z - 1)^2 (z + p) (z^2 + p^2 - 1) + 3,
//
// i.e. the result is the CUBE of f^((p^12-1)/r) — still a non-degenerate bilinear map (3 is
// prime to r).  GT values never leave the library except as opaque bytes that are compared or
// multiplied, exactly how the reference uses them.
//
// Everything here is latency code run by a handful of threads per call (the verifier is not on
// the prover hot path); it compiles under ZK_HOST_SIM so tests/host_sim can execute the same
// formulas on the CPU.
#pragma once
#include "ec.cuh"

// ---- Fp2 helpers ---------------------------------------------------------------------------------
ZK_NI Fp2 f2_mul(const Fp2& a, const Fp2& b) { return a * b; }
ZK_NI Fp2 f2_sqr(const Fp2& a) { return a.sqr(); }
ZK_HD Fp2 f2_conj(const Fp2& a) { return Fp2{a.c0, a.c1.neg()}; }
ZK_HD Fp2 f2_mul_xi(const Fp2& a) { return Fp2{a.c0 - a.c1, a.c0 + a.c1}; }  // (c0 + c1 u)(1 + u)
ZK_HD Fp2 f2_mul_fp(const Fp2& a, const Fp& s) { return Fp2{Fp::mul_call(a.c0, s), Fp::mul_call(a.c1, s)}; }

// Fp2 constant / 12-limb exponent out of the generated tables of params.cuh
#define ZK_FP2_CONST(dst, C0, C1) \
  ZK_UNROLL for (int i_ = 0; i_ < 12; i_++) { (dst).c0.v[i_] = FpParams::C0(i_); (dst).c1.v[i_] = FpParams::C1(i_); }
#define ZK_FP_LIMBS(dst, NAME) \
  ZK_UNROLL for (int i_ = 0; i_ < 12; i_++) (dst)[i_] = FpParams::NAME(i_);

// ---- Fp6 -----------------------------------------------------------------------------------------
struct Fp6 {
  Fp2 a0, a1, a2;
  static ZK_HD Fp6 zero() { return Fp6{Fp2::zero(), Fp2::zero(), Fp2::zero()}; }
  static ZK_HD Fp6 one() { return Fp6{Fp2::one(), Fp2::zero(), Fp2::zero()}; }
  ZK_HD bool operator==(const Fp6& b) const { return a0 == b.a0 && a1 == b.a1 && a2 == b.a2; }
  friend ZK_HD Fp6 operator+(const Fp6& a, const Fp6& b) { return Fp6{a.a0 + b.a0, a.a1 + b.a1, a.a2 + b.a2}; }
  friend ZK_HD Fp6 operator-(const Fp6& a, const Fp6& b) { return Fp6{a.a0 - b.a0, a.a1 - b.a1, a.a2 - b.a2}; }
  ZK_HD Fp6 neg() const { return Fp6{a0.neg(), a1.neg(), a2.neg()}; }
  ZK_HD Fp6 mul_v() const { return Fp6{f2_mul_xi(a2), a0, a1}; }
};

// Karatsuba, 6 Fp2 products
ZK_NI Fp6 f6_mul(const Fp6& a, const Fp6& b) {
  Fp2 t0 = f2_mul(a.a0, b.a0), t1 = f2_mul(a.a1, b.a1), t2 = f2_mul(a.a2, b.a2);
  Fp6 r;
  r.a0 = t0 + f2_mul_xi(f2_mul(a.a1 + a.a2, b.a1 + b.a2) - t1 - t2);
  r.a1 = f2_mul(a.a0 + a.a1, b.a0 + b.a1) - t0 - t1 + f2_mul_xi(t2);
  r.a2 = f2_mul(a.a0 + a.a2, b.a0 + b.a2) - t0 - t2 + t1;
  return r;
}

ZK_NI Fp6 f6_inverse(const Fp6& a) {
  Fp2 A = f2_sqr(a.a0) - f2_mul_xi(f2_mul(a.a1, a.a2));
  Fp2 B = f2_mul_xi(f2_sqr(a.a2)) - f2_mul(a.a0, a.a1);
  Fp2 C = f2_sqr(a.a1) - f2_mul(a.a0, a.a2);
  Fp2 F = f2_mul(a.a0, A) + f2_mul_xi(f2_mul(a.a2, B) + f2_mul(a.a1, C));
  Fp2 Fi = F.inverse();
  return Fp6{f2_mul(A, Fi), f2_mul(B, Fi), f2_mul(C, Fi)};
}

// ---- Fp12 ----------------------------------------------------------------------------------------
struct Fp12 {
  Fp6 c0, c1;
  static ZK_HD Fp12 one() { return Fp12{Fp6::one(), Fp6::zero()}; }
  ZK_HD bool operator==(const Fp12& b) const { return c0 == b.c0 && c1 == b.c1; }
  ZK_HD Fp12 conj() const { return Fp12{c0, c1.neg()}; }   // = x^(p^6); the inverse of a unitary x
};

ZK_NI Fp12 f12_mul(const Fp12& a, const Fp12& b) {
  Fp6 t0 = f6_mul(a.c0, b.c0), t1 = f6_mul(a.c1, b.c1);
  Fp12 r;
  r.c1 = f6_mul(a.c0 + a.c1, b.c0 + b.c1) - t0 - t1;
  r.c0 = t0 + t1.mul_v();
  return r;
}

// complex squaring: 2 Fp6 products
ZK_NI Fp12 f12_sqr(const Fp12& a) {
  Fp6 t = f6_mul(a.c0, a.c1);
  Fp12 r;
  r.c0 = f6_mul(a.c0 + a.c1, a.c0 + a.c1.mul_v()) - t - t.mul_v();
  r.c1 = t + t;
  return r;
}

ZK_NI Fp12 f12_inverse(const Fp12& a) {
  Fp6 d = f6_mul(a.c0, a.c0) - f6_mul(a.c1, a.c1).mul_v();
  Fp6 di = f6_inverse(d);
  return Fp12{f6_mul(a.c0, di), f6_mul(a.c1, di).neg()};
}

// x -> x^p.  As a polynomial in w (w^6 = xi) the coefficient of w^k is conjugated and multiplied
// by xi^(k (p-1)/6); k = 2 j + i for the slot a_j of c_i.
ZK_NI Fp12 f12_frobenius(const Fp12& a) {
  Fp2 g1, g2, g3, g4, g5;
  ZK_FP2_CONST(g1, frob1_c0, frob1_c1)
  ZK_FP2_CONST(g2, frob2_c0, frob2_c1)
  ZK_FP2_CONST(g3, frob3_c0, frob3_c1)
  ZK_FP2_CONST(g4, frob4_c0, frob4_c1)
  ZK_FP2_CONST(g5, frob5_c0, frob5_c1)
  Fp12 r;
  r.c0.a0 = f2_conj(a.c0.a0);                 // w^0
  r.c1.a0 = f2_mul(f2_conj(a.c1.a0), g1);     // w^1
  r.c0.a1 = f2_mul(f2_conj(a.c0.a1), g2);     // w^2
  r.c1.a1 = f2_mul(f2_conj(a.c1.a1), g3);     // w^3
  r.c0.a2 = f2_mul(f2_conj(a.c0.a2), g4);     // w^4
  r.c1.a2 = f2_mul(f2_conj(a.c1.a2), g5);     // w^5
  return r;
}

// x^z for UNITARY x (z = -0xd201000000010000 is negative: conjugate = invert)
ZK_NI Fp12 f12_exp_z(const Fp12& a) {
  const uint32_t zhi = 0xd2010000u, zlo = 0x00010000u;
  Fp12 acc = a;  // top bit (63) is set
  for (int i = 62; i >= 0; i--) {
    acc = f12_sqr(acc);
    const uint32_t bit = i >= 32 ? (zhi >> (i - 32)) & 1u : (zlo >> i) & 1u;
    if (bit) acc = f12_mul(acc, a);
  }
  return acc.conj();
}

// ---- pairing -------------------------------------------------------------------------------------
ZK_NI Fp12 pairing_line(const Fp2& lam, const Fp2& xt, const Fp2& yt, const Fp& xp, const Fp& yp) {
  Fp12 l;
  l.c0 = Fp6{yt - f2_mul(lam, xt), f2_mul_fp(lam, xp), Fp2::zero()};
  l.c1 = Fp6{Fp2::zero(), Fp2{yp.neg(), Fp::zero()}, Fp2::zero()};
  return l;
}

// f_{|z|,Q}(P); neither argument may be the identity, both must have order r
ZK_NI Fp12 miller_loop(const Affine<Fp>& p, const Affine<Fp2>& q) {
  const uint32_t zhi = 0xd2010000u, zlo = 0x00010000u;
  Fp12 f = Fp12::one();
  Fp2 rx = q.x, ry = q.y;
  for (int i = 62; i >= 0; i--) {
    // tangent at T
    Fp2 xx = f2_sqr(rx);
    Fp2 lam = f2_mul(xx.dbl() + xx, ry.dbl().inverse());
    f = f12_mul(f12_sqr(f), pairing_line(lam, rx, ry, p.x, p.y));
    Fp2 nx = f2_sqr(lam) - rx - rx;
    ry = f2_mul(lam, rx - nx) - ry;
    rx = nx;
    const uint32_t bit = i >= 32 ? (zhi >> (i - 32)) & 1u : (zlo >> i) & 1u;
    if (bit) {
      lam = f2_mul(q.y - ry, (q.x - rx).inverse());
      f = f12_mul(f, pairing_line(lam, rx, ry, p.x, p.y));
      nx = f2_sqr(lam) - rx - q.x;
      ry = f2_mul(lam, rx - nx) - ry;
      rx = nx;
    }
  }
  return f;
}

ZK_NI Fp12 final_exponentiation(const Fp12& f) {
  // easy part
  Fp12 f1 = f12_mul(f.conj(), f12_inverse(f));                 // f^(p^6 - 1)
  Fp12 f2 = f12_mul(f12_frobenius(f12_frobenius(f1)), f1);     // ^(p^2 + 1): unitary from here on
  // hard part (times 3)
  Fp12 t0 = f12_mul(f12_exp_z(f2), f2.conj());                 // f2^(z - 1)
  Fp12 t1 = f12_mul(f12_exp_z(t0), t0.conj());                 // f2^((z - 1)^2)
  Fp12 b = f12_mul(f12_exp_z(t1), f12_frobenius(t1));          // ^(z + p)
  Fp12 c = f12_mul(f12_mul(f12_exp_z(f12_exp_z(b)), f12_frobenius(f12_frobenius(b))), b.conj());  // ^(z^2 + p^2 - 1)
  return f12_mul(c, f12_mul(f12_sqr(f2), f2));                 // * f2^3
}

// ---- subgroup membership -------------------------------------------------------------------------
template <class F>
ZK_NI bool in_prime_subgroup(const Affine<F>& p) {
  if (p.is_inf()) return true;
  uint32_t k[8];
  ZK_UNROLL for (int i = 0; i < 8; i++) k[i] = FpParams::group_order(i);
  return scalar_mul(XYZZ<F>::from_affine(p), k).is_inf();
}

// ---- square roots (decompression) ----------------------------------------------------------------
// p = 3 mod 4:  sqrt(a) = a^((p+1)/4) when a is a square
ZK_NI bool fp_sqrt(const Fp& a, Fp& out) {
  uint32_t e[12];
  ZK_FP_LIMBS(e, exp_sqrt)
  out = a.pow_limbs<12>(e);
  return Fp::mul_call(out, out) == a;
}

ZK_NI Fp2 f2_pow(const Fp2& a, const uint32_t* e) {   // 12-limb little-endian exponent
  Fp2 acc = Fp2::one();
  for (int i = 11; i >= 0; i--) {
    const uint32_t w = e[i];
    for (int bit = 31; bit >= 0; bit--) {
      acc = f2_sqr(acc);
      if ((w >> bit) & 1) acc = f2_mul(acc, a);
    }
  }
  return acc;
}

// Square root in Fp2 for p = 3 mod 4 (Adj, Rodriguez-Henriquez, "Square root computation over even
// extension fields", algorithm 9): a1 = a^((p-3)/4), alpha = a1^2 a, x0 = a1 a;
// alpha^(p+1) = -1 means "no root"; alpha = -1 gives u x0, otherwise (1 + alpha)^((p-1)/2) x0.
ZK_NI bool fp2_sqrt(const Fp2& a, Fp2& out) {
  if (a.is_zero()) { out = a; return true; }
  uint32_t e[12];
  ZK_FP_LIMBS(e, exp_p_minus_3_over_4)
  Fp2 a1 = f2_pow(a, e);
  Fp2 x0 = f2_mul(a1, a);
  Fp2 alpha = f2_mul(a1, x0);
  const Fp2 minus_one = Fp2::one().neg();
  if (f2_mul(f2_conj(alpha), alpha) == minus_one) return false;
  if (alpha == minus_one) {
    out = Fp2{x0.c1.neg(), x0.c0};  // u * x0
  } else {
    ZK_FP_LIMBS(e, exp_p_minus_1_over_2)
    Fp2 b = f2_pow(Fp2::one() + alpha, e);
    out = f2_mul(b, x0);
  }
  return f2_sqr(out) == a;
}
