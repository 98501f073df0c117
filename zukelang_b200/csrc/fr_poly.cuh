// Fr polynomial kernels: QAP linear combination, coset NTT, exact quotient h = (V W - Y) / t.
//
// Replaces QAP.Make(F).eval (/root/reference/src/lib/zk/QAP.ml:120-135) and the list
// polynomial operations it calls (src/lib/zk/polynomial.ml:109-131 add/sum/mul_scalar/mul,
// :142-169 div_rem).  The reference multiplies and divides schoolbook style in O(n^2); the
// quotient is unique, so computing it by evaluation on a coset of the 2^k subgroup gives the
// same coefficients.  The subgroup generator is 5^((r-1)/2^32), the root the reference
// records at src/lib/zk/FFT.ml:208,219.
//
// Domain note: the reference's QAP lives on the integer points 0..n-1 (QAP.ml:84,92), so
// t(x) = prod (x - j) is an arbitrary degree-n polynomial, not x^n - 1.  We evaluate
// V, W, Y and t on g*H with |H| = D >= n + 1 (t has no root there), divide pointwise and
// interpolate back: deg h <= n - 2 < D, so the result is exact.
#pragma once
#include "common.cuh"
#include "runtime.cuh"

namespace zk {

struct NttPlan {
  int logD = 0;
  uint32_t D = 0;
  DevBuf<Fr> tw, tw_inv;          // omega^k, omega^-k   for k < D/2 (Montgomery form)
  DevBuf<Fr> coset, coset_inv;    // g^i ;  g^-i / D     for i < D
  void build(int logD, cudaStream_t st);
  // in-place forward transform of a length-D vector: natural order in, bit-reversed out
  void forward(Fr* d, cudaStream_t st) const;
  // in-place inverse (unscaled): bit-reversed in, natural out; multiply by coset_inv afterwards
  void inverse(Fr* d, cudaStream_t st) const;
};

// Dense QAP resident on the device (config 1 / 2 path).  With m = 0 it is
// just the quotient domain for a given target (zk_fr_quotient).
struct QapDevice {
  uint32_t m = 0, n = 0;          // variables, gates (= degree of target)
  DevBuf<Fr> vm, wm, ym;          // m x n coefficient matrices, row k = polynomial of variable k
  DevBuf<Fr> target;              // n + 1 coefficients (Montgomery)
  DevBuf<Fr> t_inv_evals;         // 1 / t on the coset, bit-reversed order, length D
  NttPlan plan;
  DevBuf<Fr> sol_m;               // witness in Montgomery form
  DevBuf<Fr> V;                   // work: V | W | Y on the coset, 3 * D
  DevBuf<Fr> Vc;                  // coefficients V | W | Y, n each (the A / B / B1 MSM scalars)
  DevBuf<Fr> H;                   // quotient coefficients (n - 1 used), Montgomery
  DevBuf<Fr> partials;
  DevBuf<Fr> eval_xs;     // the two evaluation points of the divisibility check
  DevBuf<int> flag;               // [0] non-canonical input, [1] remainder != 0
  void load(const uint8_t* v, const uint8_t* w, const uint8_t* y, const uint8_t* target, uint32_t m, uint32_t n,
            cudaStream_t st);
  // d_sol_raw: m canonical scalars on the device.  Enqueues QAP.eval: fills Vc and H.
  void eval(const uint32_t* d_sol_raw, cudaStream_t st);
  // its two halves: combine() fills Vc (the A / B scalars) and the coset-shifted work vectors,
  // quotient_from_work() then computes H — a prover can start the B query in between
  void combine(const uint32_t* d_sol_raw, cudaStream_t st);
  // V | W | Y given directly as 3 * n canonical scalars on the device
  void set_coeffs(const uint32_t* d_vwy_raw, cudaStream_t st);
  void quotient_from_work(cudaStream_t st);
};

// Evaluation-form quotient for circuits whose dense QAP.t cannot exist.
// The QAP polynomials are handled by their values on the reference's own domain 0..n-1
// (QAP.ml:84): V(j) = <gate_j.l, sol> is a sparse mat-vec, and h = (V W - Y) / t, of degree
// <= n - 2, is represented by its values at the n shifted points u_k = n + k:
//     V(u_k) = t(u_k) * sum_j w_j V(j) / (u_k - j),   w_j = 1 / prod_{i != j} (j - i)
// i.e. one cyclic convolution (size D >= 2n) of w .* V with the kernel 1/d, so that
//     h(u_k) = t(u_k) * SV_k * SW_k - SY_k.
// The matching proving key holds the Lagrange-basis points [L_j(tau)]G and
// [L'_k(tau) Z(tau) / delta]G instead of the monomial ones; the proof elements are the same
// group elements as the reference's (tests compare the bytes of both paths).
struct SparseMat {                 // CSR over Fr: n rows (gates) x m columns (variables)
  DevBuf<uint32_t> row_ptr, col;
  DevBuf<Fr> val;                  // Montgomery
};
struct EvalDomain {
  uint32_t n = 0, m = 0;
  NttPlan plan;                    // cyclic NTT of size D >= 2n
  DevBuf<Fr> w;                    // barycentric weights on 0..n-1
  DevBuf<Fr> c1;                   // t(u_k) / D^2
  DevBuf<Fr> c2;                   // 1 / D   (one element)
  DevBuf<Fr> ghat;                 // NTT of the kernel 1/d, bit-reversed
  SparseMat mat[3];                // l, r, lhs coefficient matrices (Gate.l / .r / .lhs, circuit.ml:75)
  DevBuf<Fr> sol_m, evals, work, H;  // witness; V|W|Y on the domain (3n); 3D scratch; h(u_k) (n)
  DevBuf<int> flag;                // [0] non-canonical input, [1] V(j) W(j) != Y(j) for some gate
  void load(uint32_t n, const uint8_t* w_raw, const uint8_t* t_shift_raw, cudaStream_t st);
  void load_matrix(int which, uint32_t m, const uint32_t* row_ptr, const uint32_t* col, const uint8_t* val_raw,
                   cudaStream_t st);
  // d_sol_raw: m canonical scalars.  Fills evals and H; sets flag[1] if a gate is violated.
  void eval(const uint32_t* d_sol_raw, cudaStream_t st);
  // its two halves: values() fills evals (V | W | Y on the domain: the A / B scalars) and checks the
  // gates, quotient() extrapolates and fills H — a prover can start the B query in between
  void values(const uint32_t* d_sol_raw, cudaStream_t st);
  void quotient(cudaStream_t st);
};
struct EvalDomainHandle : HandleBase {
  EvalDomain d;
  EvalDomainHandle() { kind = 6; }
};

struct QapHandle : HandleBase {
  QapDevice q;
  DevBuf<uint32_t> d_raw;   // staging for host-facing calls
  QapHandle() { kind = 3; }
};

// canonical little-endian bytes <-> Montgomery vectors (device buffers)
void fr_to_mont(const uint32_t* d_raw, Fr* d_out, uint32_t n, int* d_err, cudaStream_t st);
void fr_from_mont(const Fr* d_in, uint32_t* d_raw, uint32_t n, cudaStream_t st);

}  // namespace zk
