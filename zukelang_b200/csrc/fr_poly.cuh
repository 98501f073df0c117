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
  DevBuf<int> flag;               // [0] non-canonical input, [1] remainder != 0
  void load(const uint8_t* v, const uint8_t* w, const uint8_t* y, const uint8_t* target, uint32_t m, uint32_t n,
            cudaStream_t st);
  // d_sol_raw: m canonical scalars on the device.  Enqueues QAP.eval: fills Vc and H.
  void eval(const uint32_t* d_sol_raw, cudaStream_t st);
  // V | W | Y given directly as 3 * n canonical scalars on the device
  void set_coeffs(const uint32_t* d_vwy_raw, cudaStream_t st);
  void quotient_from_work(cudaStream_t st);
};

struct QapHandle : HandleBase {
  QapDevice q;
  DevBuf<uint32_t> d_raw;   // staging for host-facing calls
  QapHandle() { kind = 3; }
};

// canonical little-endian bytes <-> Montgomery vectors (device buffers)
void fr_to_mont(const uint32_t* d_raw, Fr* d_out, uint32_t n, int* d_err, cudaStream_t st);
void fr_from_mont(const Fr* d_in, uint32_t* d_raw, uint32_t n, cudaStream_t st);
// out[i] = a * x[i] + b * y[i]   (a, b device scalars in Montgomery form; y may be null)
void fr_axpby(const Fr* a, const Fr* x, const Fr* b, const Fr* y, Fr* out, uint32_t n, cudaStream_t st);

}  // namespace zk
