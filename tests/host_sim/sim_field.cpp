// Host simulation of the device limb algorithms (development check, tests only).
// Compiled with -DZK_HOST_SIM: ptx.cuh then emulates each PTX carry instruction.
#include "../../zukelang_b200/csrc/params.cuh"
#include "../../zukelang_b200/csrc/mont.cuh"
#include <string.h>

typedef Mont<FpParams> Fp;
typedef Mont<FrParams> Fr;

template <class F>
static void binop(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  F x, y, r;
  memcpy(x.v, a, sizeof(x.v));
  memcpy(y.v, b, sizeof(y.v));
  switch (op) {
    case 0: r = x * y; break;
    case 1: r = x + y; break;
    case 2: r = x - y; break;
    case 3: r = x.neg(); break;
    case 4: r = x.to_mont(); break;
    case 5: r = x.from_mont(); break;
    case 6: r = x.inverse(); break;
    case 7: r = x.dbl(); break;
    case 8: r = x.inverse_fermat(); break;
    case 9: r = x.sqr(); break;
    case 10: { F r2; F::mul2(x, y, y, y, r, r2); r = r + r2; break; }           // x y + y y, interleaved rows
    case 11: { F r2; F::mul2_rolled(x, y, y, y, r, r2); r = r + r2; break; }    // same, rolled row loop
    case 12: { F r2; F::mul2_rolled(y, y, x, y, r2, r); r = r + r2; break; }    // operands swapped between the two products
    default: r = F::zero();
  }
  memcpy(out, r.v, sizeof(r.v));
}

extern "C" {
void sim_fp_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) { binop<Fp>(op, a, b, out); }
void sim_fr_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) { binop<Fr>(op, a, b, out); }
}
