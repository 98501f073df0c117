/* libzkb200 — C ABI of the B200 (sm_100a) prover hot path for zukelang.
 *
 * This is the boundary a zukelang build binds through dune foreign_stubs
 * (see INTEGRATION.md).  Plain pointers and sizes only; the caller owns every
 * buffer; nothing is retained after a call returns except device-side copies
 * behind handles.  There is no CPU fallback: without a CUDA device every entry
 * point fails with ZK_ECUDA.
 *
 * Byte formats (those of the bls12-381 opam package the reference links,
 * i.e. blst / zcash serialisation):
 *   Fr   32 B little-endian canonical integer < r          (Fr.to_bytes)
 *   G1   96 B  x || y, each 48 B big-endian; identity = 0x40 then zeros
 *        48 B  compressed: x big-endian, bit7 = 1, bit6 = identity,
 *              bit5 = (y > (p-1)/2)                         (G1.to_compressed_bytes,
 *              /root/reference/src/lib/zk/curve.ml:199)
 *   G2   192 B x.c1 || x.c0 || y.c1 || y.c0; 96 B compressed, sign bit from
 *              y.c1 then y.c0                               (curve.ml:208)
 * A "point result" buffer is the uncompressed form followed by the compressed
 * form: ZK_G1_OUT = 96 + 48 bytes, ZK_G2_OUT = 192 + 96 bytes.
 *
 * Return value: 0 = ok, negative = error; zk_last_error() gives the message.
 */
#ifndef ZKB200_H
#define ZKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZK_OK 0
#define ZK_EARG (-1)       /* bad argument / length; OCaml side raises Invalid_argument */
#define ZK_EPOINT (-2)     /* non-canonical scalar or point not on the curve */
#define ZK_ECUDA (-3)      /* CUDA / NCCL failure (never rerouted to a CPU path) */
#define ZK_EREMAINDER (-4) /* QAP.eval: p is not divisible by target (QAP.ml:134 assert) */

#define ZK_FR_BYTES 32
#define ZK_G1_RAW 96
#define ZK_G1_COMP 48
#define ZK_G1_OUT 144
#define ZK_G2_RAW 192
#define ZK_G2_COMP 96
#define ZK_G2_OUT 288

/* ---- lifecycle -------------------------------------------------------------- */
/* Select the CUDA device (-1 = keep the current one) and create the library stream. */
int zk_init(int device);
int zk_shutdown(void);
const char *zk_last_error(void);
/* "name;sm_count;cc_major.cc_minor" */
int zk_device_info(char *buf, size_t cap);

/* ---- one-shot MSM -------------------------------------------------------------
 * out = sum_i scalars[i] * bases[i].
 * Replaces Curve.G.dot (curve.ml:94-103), sum_map (:91) and apply_powers (:112-118)
 * for G = G1 / G2; the OCaml stub flattens the Var.Map / list arguments.
 * inf_flags (nullable): inf_flags[i] != 0 marks bases[i] as the identity. */
int zk_g1_msm(const uint8_t *bases, const uint8_t *inf_flags, const uint8_t *scalars, size_t n,
              uint8_t out[ZK_G1_OUT]);
int zk_g2_msm(const uint8_t *bases, const uint8_t *inf_flags, const uint8_t *scalars, size_t n,
              uint8_t out[ZK_G2_OUT]);

/* ---- resident base tables (proving-key queries stay in HBM) -------------------
 * precompute != 0 additionally stores 2^(c*w) * P_i for every window w so that all
 * windows share one bucket set (W times the memory, no window combine).
 * window_bits = 0 picks c from n. */
int zk_g1_table_load(const uint8_t *bases, const uint8_t *inf_flags, size_t n, int precompute,
                     int window_bits, uint64_t *handle);
int zk_g2_table_load(const uint8_t *bases, const uint8_t *inf_flags, size_t n, int precompute,
                     int window_bits, uint64_t *handle);
/* MSM of the first n table points with host scalars (copied in, result copied out). */
int zk_g1_table_msm(uint64_t handle, const uint8_t *scalars, size_t n, uint8_t out[ZK_G1_OUT]);
int zk_g2_table_msm(uint64_t handle, const uint8_t *scalars, size_t n, uint8_t out[ZK_G2_OUT]);
/* Same with device-resident scalars / result, enqueued on `cuda_stream` (a cudaStream_t,
 * NULL = the library stream) without synchronising. */
int zk_g1_table_msm_dev(uint64_t handle, const void *d_scalars, size_t n, void *d_out, void *cuda_stream);
int zk_g2_table_msm_dev(uint64_t handle, const void *d_scalars, size_t n, void *d_out, void *cuda_stream);
/* info[0] = window bits c, [1] = windows W, [2] = bucket windows, [3] = buckets per window,
 * [4] = segments, [5] = device bytes, [6] = points, [7] = precomputed */
int zk_table_info(uint64_t handle, uint64_t info[8]);
int zk_table_free(uint64_t handle);

/* ---- fixed-base batch: out[i] = scalars[i] * generator -------------------------
 * Replaces Curve.G.of_Fr (curve.ml:180) / powers (:106-109) when applied to a vector.
 * out: n uncompressed points (ZK_G1_RAW / ZK_G2_RAW bytes each). */
int zk_g1_fixed_base_mul(const uint8_t *scalars, size_t n, uint8_t *out);
int zk_g2_fixed_base_mul(const uint8_t *scalars, size_t n, uint8_t *out);

/* ---- measurement helpers --------------------------------------------------------
 * Integer-pipe microbenchmarks.  kind: 0 = mad.lo.u32 chains,
 * 1 = mad.lo.cc / madc.hi.cc carry chains, 2 = mad.wide.u32, 3 = Fp Montgomery products
 * (register resident), 4 = Fr products, 5 = G1 XYZZ mixed adds.
 * Returns operations per second of that kind in *ops_per_s. */
int zk_bench_intpipe(int kind, int iters, double *ops_per_s, double *elapsed_ms);

/* Device unit-test hook: applies op to n operand pairs of field elements given as raw
 * little-endian 32-bit limbs (tests/test_device_field.py).  field: 0 = Fp (12 limbs),
 * 1 = Fr (8 limbs).  op: 0 mul, 1 add, 2 sub, 3 neg, 4 to_mont, 5 from_mont, 6 inverse, 7 dbl. */
int zk_test_field_op(int field, int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
