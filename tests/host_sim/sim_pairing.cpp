// Host simulation of the Fp12 tower, pairing and square-root formulas (development check, tests only).
// All operands are Montgomery limbs, little-endian 32-bit.
#include "../../zukelang_b200/csrc/pairing.cuh"
#include <string.h>

extern "C" {
// op: 0 mul, 1 sqr, 2 inverse, 3 frobenius, 4 conj, 5 exp_z
void sim_f12_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fp12 A, B, r;
  memcpy(&A, a, sizeof(A));
  memcpy(&B, b, sizeof(B));
  switch (op) {
    case 0: r = f12_mul(A, B); break;
    case 1: r = f12_sqr(A); break;
    case 2: r = f12_inverse(A); break;
    case 3: r = f12_frobenius(A); break;
    case 4: r = A.conj(); break;
    default: r = f12_exp_z(A); break;
  }
  memcpy(out, &r, sizeof(r));
}
void sim_miller(const uint32_t* p, const uint32_t* q, uint32_t* out) {
  Affine<Fp> P;
  Affine<Fp2> Q;
  memcpy(&P, p, sizeof(P));
  memcpy(&Q, q, sizeof(Q));
  Fp12 f = miller_loop(P, Q);
  memcpy(out, &f, sizeof(f));
}
void sim_final_exp(const uint32_t* a, uint32_t* out) {
  Fp12 A;
  memcpy(&A, a, sizeof(A));
  Fp12 r = final_exponentiation(A);
  memcpy(out, &r, sizeof(r));
}
int sim_fp_sqrt(const uint32_t* a, uint32_t* out) {
  Fp A, r;
  memcpy(&A, a, sizeof(A));
  bool ok = fp_sqrt(A, r);
  memcpy(out, &r, sizeof(r));
  return ok ? 1 : 0;
}
int sim_fp2_sqrt(const uint32_t* a, uint32_t* out) {
  Fp2 A, r = Fp2::zero();
  memcpy(&A, a, sizeof(A));
  bool ok = fp2_sqrt(A, r);
  memcpy(out, &r, sizeof(r));
  return ok ? 1 : 0;
}
int sim_in_subgroup(int g2, const uint32_t* p) {
  if (g2) {
    Affine<Fp2> Q;
    memcpy(&Q, p, sizeof(Q));
    return in_prime_subgroup(Q) ? 1 : 0;
  }
  Affine<Fp> P;
  memcpy(&P, p, sizeof(P));
  return in_prime_subgroup(P) ? 1 : 0;
}
}
