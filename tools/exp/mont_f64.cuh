// EXPERIMENT for round 2 — NOT included by any translation unit of libzkb200.
//
// Montgomery product in Fp (BLS12-381, R = 2^384 — the residues the rest of the library uses) with
// the limb products on the FP64 pipe instead of the IMAD pipe.
//
// Representation: 8 limbs of 48 bits in uint64_t (little-endian limb order).  A limb product
// a * b < 2^96 is split exactly by two round-toward-zero FMAs (N. Emmart's construction):
//
//     hi = fma_rz(a, b, 2^104)                 = 2^104 + floor(a b / 2^52) * 2^52     (ulp(2^104) = 2^52)
//     lo = fma_rz(a, b, (2^104 + 2^52) - hi)   = 2^52 + (a b mod 2^52)                 (exact)
//
// so the 52-bit mantissa fields of hi and lo ARE floor(ab / 2^52) and ab mod 2^52.  Columns are
// summed as 64-bit integers over the raw bit patterns (the constant exponent fields are subtracted
// once per column), normalised to 48-bit limbs with shifts, and reduced word by word
// (q_i = t_i * n0' mod 2^48, then q_i * p is added to the columns).
//
// Instruction budget per product: 136 limb products = 272 DFMA, ~290 64-bit integer additions
// (ideally 3-input IADD3 / IADD3.X pairs), ~60 shifts / masks, 16 int<->double re-biasings.
//
// ZK_HOST_SIM: fma_rz is emulated with exact 128-bit integer arithmetic (every operand here is an
// integer-valued double), so tests/host_sim runs the identical algorithm on the CPU.
#pragma once
#include <stdint.h>
#include <string.h>
#include "../../zukelang_b200/csrc/ptx.cuh"

namespace f64mont {

#ifdef ZK_HOST_SIM
inline uint64_t d2bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
inline double bits2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
// a, b: non-negative integers < 2^53; c: integer-valued, |c| < 2^106.  Exact a*b + c, truncated
// toward zero to 53 significant bits.
inline double fma_rz(double a, double b, double c) {
  __int128 s = (__int128)(uint64_t)a * (uint64_t)b;
  const bool cneg = c < 0;
  double ca = cneg ? -c : c;
  uint64_t cb = d2bits(ca);
  int e = (int)((cb >> 52) & 0x7ff) - 1075;                  // ca = m * 2^e
  unsigned __int128 m = (cb & 0xfffffffffffffULL) | (ca != 0 ? 0x10000000000000ULL : 0);
  __int128 ci = e >= 0 ? (__int128)(m << e) : (__int128)(m >> -e);   // c is an integer: no bits lost
  s = cneg ? s - ci : s + ci;
  const bool neg = s < 0;
  unsigned __int128 u = neg ? (unsigned __int128)(-s) : (unsigned __int128)s;
  int bl = 0;
  for (unsigned __int128 t = u; t; t >>= 1) bl++;
  if (bl > 53) u = (u >> (bl - 53)) << (bl - 53);            // toward zero
  double r = 0, scale = 1;
  for (int i = 0; i < 128; i += 32) { r += (double)(uint32_t)(u >> i) * scale; scale *= 4294967296.0; }   // exact: <= 53 significant bits
  return neg ? -r : r;
}
#else
__device__ __forceinline__ uint64_t d2bits(double x) { return (uint64_t)__double_as_longlong(x); }
__device__ __forceinline__ double bits2d(uint64_t u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ double fma_rz(double a, double b, double c) {
  double d;
  asm("fma.rz.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
  return d;
}
#endif

constexpr int NL = 8;                                  // limbs
constexpr uint64_t MASK48 = 0xffffffffffffULL;
constexpr uint64_t BITS_2P52 = 0x4330000000000000ULL;  // bit pattern of 2^52
constexpr uint64_t BITS_2P104 = 0x4670000000000000ULL; // bit pattern of 2^104

// integer < 2^52 -> the same value as a double (re-biasing: one OR and one FP add)
ZK_HD double u2d(uint64_t x) { return bits2d(BITS_2P52 | x) - 4503599627370496.0; }

struct Limbs { uint64_t v[NL]; };

// r = a * b / 2^384 mod p for fully reduced 48-bit-limb operands; p and n0' = -p^-1 mod 2^48 are
// passed in so that the
// caller decides where they live (registers, constant memory)
ZK_HD Limbs mul(const Limbs& a, const Limbs& b, const Limbs& p, uint64_t n0inv48) {
  const double C1 = 20282409603651670423947251286016.0;                    // 2^104
  const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;  // 2^104 + 2^52
  double ad[NL], bd[NL], pd[NL];
  ZK_UNROLL for (int i = 0; i < NL; i++) { ad[i] = u2d(a.v[i]); bd[i] = u2d(b.v[i]); pd[i] = u2d(p.v[i]); }
  // column sums of the raw bit patterns
  uint64_t HI[2 * NL], LO[2 * NL];
  ZK_UNROLL for (int k = 0; k < 2 * NL; k++) { HI[k] = 0; LO[k] = 0; }
  ZK_UNROLL for (int i = 0; i < NL; i++) {
    ZK_UNROLL for (int j = 0; j < NL; j++) {
      const double hi = fma_rz(ad[i], bd[j], C1);
      const double lo = fma_rz(ad[i], bd[j], C2 - hi);
      HI[i + j] += d2bits(hi);
      LO[i + j] += d2bits(lo);
    }
  }
  // remove the exponent fields once per column: column k receives cnt(k) products of a * b and,
  // by the time it is read, cnt(k) products of q * p as well (arithmetic is mod 2^64 until then)
  ZK_UNROLL for (int k = 0; k < 2 * NL - 1; k++) {
    const uint64_t cnt = 2 * (uint64_t)(k < NL ? k + 1 : 2 * NL - 1 - k);
    HI[k] -= cnt * BITS_2P104;
    LO[k] -= cnt * BITS_2P52;
  }
  // word-serial Montgomery reduction; `carry` is the part of the lower columns above their 48 bits
  uint64_t carry = 0;
  ZK_UNROLL for (int i = 0; i < NL; i++) {
    const uint64_t t = (LO[i] + carry) & MASK48;                 // low limb of column i so far
    // q = t * n0' mod 2^48: only the low half of the product is needed
    const double td = u2d(t), nd = u2d(n0inv48);
    const double qh = fma_rz(td, nd, C1);
    const double ql = fma_rz(td, nd, C2 - qh);
    const uint64_t q = d2bits(ql) & MASK48;                      // (mantissa = t n0' mod 2^52) mod 2^48
    const double qd = u2d(q);
    ZK_UNROLL for (int j = 0; j < NL; j++) {
      const double hi = fma_rz(qd, pd[j], C1);
      const double lo = fma_rz(qd, pd[j], C2 - hi);
      HI[i + j] += d2bits(hi);
      LO[i + j] += d2bits(lo);
    }
    // column i is now 0 mod 2^48: push everything above bit 48 into the carry
    carry = ((LO[i] + carry) >> 48) + (HI[i] << 4);
  }
  Limbs r;
  ZK_UNROLL for (int k = 0; k < NL; k++) {
    const uint64_t s = LO[NL + k] + carry;
    r.v[k] = s & MASK48;
    carry = (s >> 48) + (HI[NL + k] << 4);
  }
  // result < 2 p: one conditional subtraction
  uint64_t d[NL], borrow = 0;
  ZK_UNROLL for (int k = 0; k < NL; k++) {
    const uint64_t x = r.v[k] - p.v[k] - borrow;
    d[k] = x & MASK48;
    borrow = (x >> 63) & 1;
  }
  if (!borrow) { ZK_UNROLL for (int k = 0; k < NL; k++) r.v[k] = d[k]; }
  return r;
}

}  // namespace f64mont
