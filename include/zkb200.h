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
/* Drive ONE CUDA device (-1 = keep the current one) and create the library stream. */
int zk_init(int device);
/* Drive ndev (1..8) devices of one box from this single process:
 * devs[0] is the primary device (QAP / domain handles live there, results are combined there).
 * Key and table handles loaded afterwards spread their base-point ranges over all the devices by
 * themselves, and zk_groth16_prove* / zk_g*_table_msm* return FINISHED results: every device
 * reduces its range on its own stream and stores its partial sum into the primary device's memory
 * (peer-to-peer over NVLink; peer access between devs[0] and every other device is required,
 * ZK_ECUDA otherwise), where the partials are added.  This is what lets the OCaml `prove`
 * (protocol.mli:23) use the whole box.  The device list can only change after zk_shutdown. */
int zk_init_devices(const int *devs, int ndev);
int zk_device_count(void); /* devices driven; 0 before zk_init */
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
/* `count` MSMs over the first n table points: scalars[i] points to the i-th host scalar vector,
 * out receives count point results.  The MSMs run in groups (a short first group, then up to the
 * queue depth): a group is uploaded on a copy stream while the previous group is accumulated, and
 * sorted, accumulated and reduced as one launch sequence, so a caller with many MSMs over one key
 * (a stream of proofs) gets the device-resident rate end to end.  Pinned host memory makes the
 * uploads asynchronous. */
int zk_g1_table_msm_batch(uint64_t handle, const uint8_t *const *scalars, size_t n, size_t count, uint8_t *out);
int zk_g2_table_msm_batch(uint64_t handle, const uint8_t *const *scalars, size_t n, size_t count, uint8_t *out);
/* Same with device-resident scalars / result, enqueued on `cuda_stream` (a cudaStream_t,
 * NULL = the library stream) without synchronising.  Device pointers belong to one device: these
 * entry points (and zk_table_pipeline / zk_table_join) need a table that lives on ONE device
 * (zk_init, or fewer than 4096 points per device), ZK_EARG otherwise. */
int zk_g1_table_msm_dev(uint64_t handle, const void *d_scalars, size_t n, void *d_out, void *cuda_stream);
int zk_g2_table_msm_dev(uint64_t handle, const void *d_scalars, size_t n, void *d_out, void *cuda_stream);
/* info[0] = window bits c, [1] = windows W, [2] = bucket windows, [3] = buckets per window,
 * [4] = 1 (legacy field), [5] = device bytes (all devices), [6] = points, [7] = precomputed;
 * c / W are those of the primary device's part when the table is spread over several devices */
int zk_table_info(uint64_t handle, uint64_t info[8]);
/* Pipelining of consecutive *_msm_dev calls on one table.  With enable = 0 (default) every call
 * runs its MSM at once: plain stream order.  With enable != 0 a call only QUEUES its MSM (scalar
 * pointer, count, output slot) and zk_table_join(handle, cuda_stream) — or a call that finds the
 * queue full — runs everything queued as ONE launch sequence: one counting sort and one balanced
 * accumulation over all the queued MSMs' buckets, then one batched tail (bucket reduction, affine
 * conversion: ~40 dependent point operations whose wall time is the same for one MSM as for a
 * batch).  enable = 1: default depth (32 MSMs per join, or ZKB200_QUEUE; fewer for tables beyond
 * 2^21 points, whose sorted entry lists are kept below 2 GB); enable = 2..32: that depth.
 * d_scalars of a queued call must stay valid and unchanged, and d_out is valid on cuda_stream,
 * after the join. */
int zk_table_pipeline(uint64_t handle, int enable);
int zk_table_join(uint64_t handle, void *cuda_stream);
/* Stage timing for the roofline leg of bench.py: enable != 0 makes the following joins on this
 * table (one per MSM unless pipelined) bracket their stages with CUDA events on the launching
 * stream; stage_ms (nullable) receives the last profiled join's times once that stream has been
 * synchronised:
 * [0] digits + scan + scatter, [1] bucket accumulation, [2] bucket reduction, [3] window combine. */
int zk_table_profile(uint64_t handle, int enable, float stage_ms[4]);
/* Sums of those stage times over every join recorded since profiling was switched on (at most the
 * last 64 joins): counts[0] = MSMs the joins held, counts[1] = joins.  With a pipelined table one join
 * sorts and accumulates all its queued MSMs in one launch sequence, so the accumulation kernel's
 * time per MSM inside a timed region is totals[1] / counts[0]. */
int zk_table_profile_totals(uint64_t handle, float totals[4], uint64_t counts[2]);
/* Per-step timing of zk_g*_table_msm_batch (bench.py's e2e leg): enable != 0 makes the following
 * batch calls on this table record, for each of their first 64 steps, when its upload started and
 * ended (copy stream) and when the kernels of its GROUP (the steps joined together) started and
 * ended (compute stream).  out (nullable, with steps) receives 4 floats per step of the LAST batch
 * call, in ms since its first upload started. */
int zk_table_batch_timing(uint64_t handle, int enable, float *out, size_t cap, size_t *steps);
int zk_table_free(uint64_t handle);

/* ---- sum of k points (combining the shards' partial results) ------
 * points: k uncompressed points; out: a point result.  Replaces G.sum / repeated G.( + )
 * (curve.ml:163,178).  The _dev forms take device pointers and enqueue on cuda_stream; being
 * asynchronous they cannot return ZK_EPOINT: an input point that does not parse makes them write
 * 0xff into every byte of the affected result, which no parser of this library accepts. */
int zk_g1_sum(const uint8_t *points, size_t k, uint8_t out[ZK_G1_OUT]);
int zk_g2_sum(const uint8_t *points, size_t k, uint8_t out[ZK_G2_OUT]);
int zk_g1_sum_dev(const void *d_points, size_t k, void *d_out, void *cuda_stream);
int zk_g2_sum_dev(const void *d_points, size_t k, void *d_out, void *cuda_stream);
/* Batched form for one gather of `batch` queued results from k shards: d_points is laid out
 * [shard][batch] (what all_gather_into_tensor produces), out[q] = sum over shards of point q. */
int zk_g1_sum_strided_dev(const void *d_points, size_t k, size_t batch, void *d_out, void *cuda_stream);
int zk_g2_sum_strided_dev(const void *d_points, size_t k, size_t batch, void *d_out, void *cuda_stream);

/* ---- fixed-base batch: out[i] = scalars[i] * generator -------------------------
 * Replaces Curve.G.of_Fr (curve.ml:180) / powers (:106-109) when applied to a vector.
 * out: n uncompressed points (ZK_G1_RAW / ZK_G2_RAW bytes each). */
int zk_g1_fixed_base_mul(const uint8_t *scalars, size_t n, uint8_t *out);
int zk_g2_fixed_base_mul(const uint8_t *scalars, size_t n, uint8_t *out);

/* ---- QAP evaluation -----------------------------------------------------------------
 * Replaces QAP.Make(F).eval (/root/reference/src/lib/zk/QAP.ml:120-135): with
 * V = sum_k sol_k v_k, W, Y likewise, returns h with h * target = V W - Y.
 * zk_qap_load uploads the dense QAP.t (QAP.ml:11-16): v, w, y are m x n row-major
 * matrices of Fr (row k = coefficients of variable k's polynomial, lowest degree first,
 * zero padded to n = degree(target)); target has n + 1 coefficients.
 * zk_qap_eval: sol = m scalars in the same variable order; h_out (nullable) receives n - 1
 * coefficients; vwy_out (nullable) receives V | W | Y, n coefficients each.
 * Returns ZK_EREMAINDER where the reference's `assert (is_zero rem)` (QAP.ml:134) fails. */
int zk_qap_load(const uint8_t *v, const uint8_t *w, const uint8_t *y, const uint8_t *target, size_t m,
                size_t n, uint64_t *handle);
int zk_qap_eval(uint64_t handle, const uint8_t *sol, uint8_t *h_out, uint8_t *vwy_out);
int zk_qap_free(uint64_t handle);
/* Polynomial.div_rem specialised to the exact quotient (polynomial.ml:142-169):
 * h = (V W - Y) / T for coefficient vectors V, W, Y (n each) and T (n + 1); h_out gets n - 1. */
int zk_fr_quotient(const uint8_t *V, const uint8_t *W, const uint8_t *Y, const uint8_t *T, size_t n,
                   uint8_t *h_out);

/* Target-only domain (no dense matrices) for zk_groth16_prove_coeffs; freed with zk_qap_free. */
int zk_quotient_domain_load(const uint8_t *target, size_t n, uint64_t *handle);

/* ---- Groth16 ---------------------------------------------------------------------------
 * The proving key of /root/reference/src/groth16/groth16.ml:24-34, flattened.  Var.Map fields
 * are given as arrays in increasing Var order; mid_index[j] is the position of the j-th
 * mid variable in the witness vector (which lists ALL variables in increasing Var order). */
typedef struct {
  size_t n;                  /* degree of target = number of gates */
  size_t m;                  /* number of variables = witness length */
  size_t n_mid;              /* |Dom(ltd_mid)| */
  size_t n_h;                /* points in tiztd; 0 means n - 1 (the reference's length) */
  const uint32_t *mid_index; /* n_mid entries */
  const uint8_t *a, *b1, *d1;   /* G1: alpha, beta, delta            (ZK_G1_RAW each) */
  const uint8_t *b2, *d2;       /* G2: beta, delta                   (ZK_G2_RAW each) */
  const uint8_t *ti1;           /* G1: tau^i, at least n points      (groth16.ml:73 has n + 2) */
  const uint8_t *ti2;           /* G2: tau^i, at least n points      (groth16.ml:87) */
  const uint8_t *tiztd;         /* G1: tau^i Z(tau)/delta, n - 1     (groth16.ml:80-83) */
  const uint8_t *ltd_mid;       /* G1: L_k(tau)/delta, n_mid points  (groth16.ml:74-79) */
} zk_groth16_pkey;

#define ZK_GROTH16_PROOF_OUT (ZK_G1_OUT + ZK_G2_OUT + ZK_G1_OUT) /* a | b | c point results */

/* shard_index / shard_count = (0, 1): the whole key.  After zk_init_devices its list-valued fields
 * are spread over the devices and zk_groth16_prove* still return the finished proof.
 * shard_count > 1 (one PROCESS per GPU, e.g. torchrun): this process keeps slice shard_index of
 * every list-valued field on its device and zk_groth16_prove* return this
 * shard's PARTIAL sums, which the caller adds across shards (zk_g*_sum). */
int zk_groth16_pk_load(const zk_groth16_pkey *pk, int shard_index, int shard_count, uint64_t *handle);
/* Groth16.Make(C).prove (groth16.ml:235-237 then 123-161).  sol: m witness scalars; r, s: the two
 * Fr.gen draws of groth16.ml:124-125 (r first), supplied by the host RNG. */
int zk_groth16_prove(uint64_t pk, uint64_t qap, const uint8_t *sol, const uint8_t r[32], const uint8_t s[32],
                     uint8_t proof_out[ZK_GROTH16_PROOF_OUT]);
/* Same with the combinations V | W | Y given directly (3 * n scalars); `qap` may be a
 * zk_quotient_domain_load handle.  Used when the dense QAP.t cannot exist. */
int zk_groth16_prove_coeffs(uint64_t pk, uint64_t qap, const uint8_t *vwy, const uint8_t *sol,
                            const uint8_t r[32], const uint8_t s[32], uint8_t proof_out[ZK_GROTH16_PROOF_OUT]);

/* ---- evaluation-form Groth16 for circuits too large for a dense QAP.t ----
 * The QAP polynomials are handled by their values on the reference's own domain 0..n-1
 * (QAP.ml:84) and h by its values on n..2n-1; the key given to zk_groth16_pk_load then holds the
 * Lagrange-basis points ti1 = [L_j(tau)]1, ti2 = [L_j(tau)]2, tiztd = [L'_k(tau) Z(tau)/delta]1
 * (n_h = n) derived at key generation.  Proof elements are the same group elements as
 * zk_groth16_prove's.
 * zk_eval_domain_load: w = barycentric weights 1 / prod_{i != j} (j - i), t_shift[k] = t(n + k).
 * zk_r1cs_load: CSR matrix `which` (0 = Gate.l, 1 = Gate.r, 2 = Gate.lhs; circuit.ml:75) with n
 * rows and m columns (variables in increasing Var order).  Handles are freed with zk_qap_free.
 * zk_groth16_prove_r1cs: `sol` = m canonical scalars in host memory (pinned memory makes the
 * upload one asynchronous DMA) — or in the primary device's memory, complete before the call: a
 * sharded caller (one process per GPU) uploads 1/N of the witness per rank and all-gathers it over
 * NVLink instead of sending the whole witness down every PCIe link. */
int zk_eval_domain_load(size_t n, const uint8_t *w, const uint8_t *t_shift, uint64_t *handle);
int zk_r1cs_load(uint64_t domain, int which, size_t m, const uint32_t *row_ptr, const uint32_t *col,
                 const uint8_t *val);
int zk_groth16_prove_r1cs(uint64_t pk, uint64_t domain, const uint8_t *sol, const uint8_t r[32],
                          const uint8_t s[32], uint8_t proof_out[ZK_GROTH16_PROOF_OUT]);

/* ---- Pinocchio (Protocol 2) --------------------------------------------------------------
 * The proving key of /root/reference/src/pinocchio/pinocchio.ml:37-60, flattened as above.
 * `one` is G1.one (used by ZKCompute.f's "- one * dy", pinocchio.ml:485). */
typedef struct {
  size_t n, m, n_mid;
  const uint32_t *mid_index;
  const uint8_t *vv, *yy, *vav, *yay, *bvwy; /* G1, n_mid points each */
  const uint8_t *ww, *waw;                   /* G2, n_mid points each */
  const uint8_t *si;                         /* G1, n + 1 points (pinocchio.ml:133) */
  const uint8_t *v_all, *w_all;              /* G1, m points each */
  const uint8_t *one, *vt, *yt, *vavt, *yayt, *vbt, *wbt, *ybt; /* G1 single points */
  const uint8_t *wt, *wawt;                  /* G2 single points */
} zk_pinocchio_pkey;

#define ZK_PINOCCHIO_PROOF_OUT (6 * ZK_G1_OUT + 2 * ZK_G2_OUT)

/* Device time (ms, CUDA events on the primary device's stream: first upload to last download) of
 * the last zk_groth16_prove* call on this key; for bench.py. */
/* One process per GPU: adds the k shards' partial results — k proof-out buffers, as returned by
 * zk_groth16_prove* on keys loaded with shard_count = k — into the finished proof (one upload, one
 * launch, one download; replaces three zk_g*_sum calls). */
int zk_groth16_combine(const uint8_t *parts, size_t k, uint8_t proof_out[ZK_GROTH16_PROOF_OUT]);
int zk_groth16_last_device_ms(uint64_t pk, float *ms);
/* Stage split of that time, in ms (primary device): [0] witness upload, [1] QAP evaluation
 * (V | W | Y), [2] B: scalars, counting sort, accumulation, [3] quotient h(x) — B's tail runs beside
 * it on a second stream —, [4] A and C: scalars and counting sort (one entry list), [5] their
 * accumulation (one launch), [6] the G1 tail and what is left of the G2 tail, [7] wait for the
 * other devices + combine + download. */
int zk_groth16_last_stage_ms(uint64_t pk, float out[8]);

int zk_pinocchio_pk_load(const zk_pinocchio_pkey *pk, int shard_index, int shard_count, uint64_t *handle);
/* d = dv | dw | dy (3 * 32 B, the draws of pinocchio.ml:428-430) selects ZK.prove (:559-561);
 * d = NULL selects NonZK.prove (:536-538).  proof_out holds vv | ww | yy | h | vavv | waww |
 * yayy | bvwy point results (record order of pinocchio.ml:195-208). */
int zk_pinocchio_prove(uint64_t pk, uint64_t qap, const uint8_t *sol, const uint8_t *d,
                       uint8_t proof_out[ZK_PINOCCHIO_PROOF_OUT]);
int zk_key_free(uint64_t handle);

/* ---- verifier side ----------------------------------------------
 * out = prod_i e(+-g1[i], g2[i]) with ONE final exponentiation.  Replaces Pairing.pairing and the
 * GT sums / differences of groth16.ml:103,163-173 and pinocchio.ml:254-420 (GT is written
 * additively there, curve.ml:212-220): `e a b + e c d - e f g` is one call with n = 3 and
 * negate = {0, 0, 1}.  negate may be NULL.  Pairs containing the identity contribute 1.  n <= 4096.
 * Every point must be canonical, on its curve and in the prime-order subgroup (ZK_EPOINT).
 *
 * A GT value is ZK_GT_BYTES opaque bytes: the 12 Fp coefficients of the library's own
 * Fp2-Fp6-Fp12 tower, 48 B big-endian each.  It is NOT blst's GT encoding and the pairing is the
 * cube of the reduced ate pairing (see zukelang_b200/csrc/pairing.cuh); both are invisible to a
 * caller that only compares and multiplies GT values, which is all the reference does. */
#define ZK_GT_BYTES 576
int zk_pairing_product(const uint8_t *g1_96, const uint8_t *g2_192, const uint8_t *negate, size_t n,
                       uint8_t out[ZK_GT_BYTES]);
/* k independent products in one call: product g takes the next counts[g] pairs of g1 / g2 / negate
 * (sum of counts <= 4096) and writes out[g * ZK_GT_BYTES ..].  A verifier with several GT
 * equations (the five checks of pinocchio.ml:254-420) pays the latency of one. */
int zk_pairing_product_batch(const uint8_t *g1_96, const uint8_t *g2_192, const uint8_t *negate,
                             const uint32_t *counts, size_t k, uint8_t *out);
/* GT.( + ) of curve.ml:212-220 (the product in the multiplicative notation). */
int zk_gt_mul(const uint8_t a[ZK_GT_BYTES], const uint8_t b[ZK_GT_BYTES], uint8_t out[ZK_GT_BYTES]);

/* ---- wire format reader --------------------------------------------
 * G1/G2.of_compressed_bytes_exn (curve.ml:201,210): n compressed points (48 B / 96 B each) ->
 * n uncompressed points.  ZK_EPOINT when a flag byte is malformed, a coordinate is >= p, x is not
 * on the curve, or the point is outside the prime-order subgroup. */
int zk_g1_decompress(const uint8_t *comp48, size_t n, uint8_t *out96);
int zk_g2_decompress(const uint8_t *comp96, size_t n, uint8_t *out192);

/* ---- measurement helpers --------------------------------------------------------
 * Integer-pipe microbenchmarks.  kind: 0 = mad.lo.u32 chains,
 * 1 = mad.lo.cc / madc.hi.cc carry chains, 2 = mad.wide.u32, 3 = Fp Montgomery products
 * (register resident), 4 = Fr products, 5 = G1 XYZZ mixed adds.
 * Returns operations per second of that kind in *ops_per_s. */
int zk_bench_intpipe(int kind, int iters, double *ops_per_s, double *elapsed_ms);

/* Device unit-test hook: applies op to n operand pairs of field elements given as raw
 * little-endian 32-bit limbs (tests/test_device_field.py).  field: 0 = Fp (12 limbs),
 * 1 = Fr (8 limbs).  op: 0 mul, 1 add, 2 sub, 3 neg, 4 to_mont, 5 from_mont, 6 inverse, 7 dbl,
 * 8 sqr (the dedicated Montgomery square of the bucket accumulation), 9 / 10 = first / second
 * result of the interleaved pair mul2(a, b, a + b, a - b). */
int zk_test_field_op(int field, int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n);
/* Device unit-test hook for the mixed addition of the bucket accumulation: out[i] = 2 p[i] + q[i]
 * (n uncompressed G1 points each; out = n point results).  variant: 0 = madd, 1 = madd_paired
 * (interleaved products, what k_accumulate runs), 2 = general XYZZ add. */
int zk_test_g1_madd(const uint8_t *p, const uint8_t *q, int variant, size_t n, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
