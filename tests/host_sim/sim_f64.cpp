// Host simulation of the experimental FP64-pipe Montgomery product (csrc/exp/mont_f64.cuh).
#include "../../tools/exp/mont_f64.cuh"

extern "C" void sim_f64_mul(const uint64_t* a, const uint64_t* b, const uint64_t* p, uint64_t n0inv, uint64_t* out) {
  f64mont::Limbs A, B, P;
  for (int i = 0; i < 8; i++) { A.v[i] = a[i]; B.v[i] = b[i]; P.v[i] = p[i]; }
  f64mont::Limbs r = f64mont::mul(A, B, P, n0inv);
  for (int i = 0; i < 8; i++) out[i] = r.v[i];
}
extern "C" double sim_f64_fma_rz(double a, double b, double c) { return f64mont::fma_rz(a, b, c); }
