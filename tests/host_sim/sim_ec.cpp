// Host simulation of the XYZZ point formulas (development check, tests only).
#include "../../zukelang_b200/csrc/ec.cuh"
#include <string.h>

template <class F>
static void ec_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* k, uint32_t* out) {
  XYZZ<F> A, B;
  Affine<F> Q;
  memcpy(&A, a, sizeof(A));
  memcpy(&B, b, sizeof(B));
  memcpy(&Q, b, sizeof(Q));
  XYZZ<F> r = XYZZ<F>::inf();
  switch (op) {
    case 0: r = A; r.madd(Q); break;          // XYZZ += affine
    case 1: r = A; r.add(B); break;           // XYZZ += XYZZ
    case 2: r = A.dbl(); break;
    case 3: r = scalar_mul(A, k); break;
    case 4: r = XYZZ<F>::dbl_affine(Q); break;
    case 5: r = A; r.madd_paired(Q); break;   // interleaved product pairs (Mont::mul2 / Fp2::mul2)
  }
  Affine<F> aff = r.to_affine();
  memcpy(out, &aff, sizeof(aff));
  memcpy(out + sizeof(aff) / 4, &r, sizeof(r));
}

extern "C" {
void sim_g1_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* k, uint32_t* out) { ec_op<Fp>(op, a, b, k, out); }
void sim_g2_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* k, uint32_t* out) { ec_op<Fp2>(op, a, b, k, out); }
}
