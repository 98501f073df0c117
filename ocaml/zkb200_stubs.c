/* OCaml foreign stubs over libzkb200 (include/zkb200.h).
 *
 * Shipped as SOURCE: no OCaml toolchain exists in the build image (SURVEY.md headline fact 3),
 * so this file is compiled by the zukelang maintainer with the dune stanza in ocaml/dune.
 * Every stub copies its OCaml Bytes arguments before releasing the runtime lock and copies the
 * result back afterwards; no OCaml heap pointer is held across the blocking section.
 */
#include <string.h>
#include <stdlib.h>
#include <caml/alloc.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>
#include "zkb200.h"

static void zk_raise(int rc) {
  if (rc == ZK_EARG) caml_invalid_argument(zk_last_error());   /* curve.ml:116 Invalid_argument */
  caml_failwith(zk_last_error());                              /* assert false / assert (is_zero rem) */
}
/* an argument error found by the stub itself: zk_last_error() would be a stale library message */
static void zk_raise_arg(const char *msg) { caml_invalid_argument(msg); }

static uint8_t *dup_bytes(value v, size_t *len) {
  size_t n = caml_string_length(v);
  uint8_t *p = (uint8_t *)malloc(n ? n : 1);
  if (!p) caml_failwith("zkb200: out of memory copying an argument");
  memcpy(p, Bytes_val(v), n);
  if (len) *len = n;
  return p;
}

CAMLprim value zkb200_init(value dev) {
  CAMLparam1(dev);
  int rc = zk_init(Int_val(dev));
  if (rc) zk_raise(rc);
  CAMLreturn(Val_unit);
}

/* init_devices : int array -> unit.  One process drives the listed GPUs of the box; keys and tables
 * loaded afterwards spread over them and prove returns the finished proof (protocol.mli:23). */
CAMLprim value zkb200_init_devices(value devs) {
  CAMLparam1(devs);
  int d[8];
  size_t n = Wosize_val(devs);
  if (n < 1 || n > 8) zk_raise_arg("zkb200_init_devices: 1 to 8 devices expected");
  for (size_t i = 0; i < n; i++) d[i] = Int_val(Field(devs, i));
  int rc = zk_init_devices(d, (int)n);
  if (rc) zk_raise(rc);
  CAMLreturn(Val_unit);
}

/* ---- resident base tables: a key field uploaded once, many MSMs against it -------------------
 * table_load : g2:bool -> bases:bytes -> precompute:bool -> int64;  table_msm : g2:bool -> int64 ->
 * scalars:bytes -> bytes;  table_free : int64 -> unit.  Behind Curve.G.dot / apply_powers
 * (curve.ml:94-118) when the same point list is used again (ocaml/zkb200.ml: Resident). */
CAMLprim value zkb200_table_load(value g2, value bases, value precompute) {
  CAMLparam3(g2, bases, precompute);
  size_t nb;
  const int is2 = Int_val(g2) != 0, pre = Int_val(precompute) != 0;
  const size_t raw = is2 ? ZK_G2_RAW : ZK_G1_RAW;
  uint8_t *b = dup_bytes(bases, &nb);
  uint64_t h = 0;
  int rc = 0;
  if (nb == 0 || nb % raw) { free(b); zk_raise_arg("zkb200_table_load: bases must be a non-empty multiple of the point size"); }
  caml_release_runtime_system();
  rc = is2 ? zk_g2_table_load(b, NULL, nb / raw, pre, 0, &h) : zk_g1_table_load(b, NULL, nb / raw, pre, 0, &h);
  caml_acquire_runtime_system();
  free(b);
  if (rc) zk_raise(rc);
  CAMLreturn(caml_copy_int64((int64_t)h));
}

CAMLprim value zkb200_table_msm(value g2, value handle, value scalars) {
  CAMLparam3(g2, handle, scalars);
  CAMLlocal1(out);
  size_t ns;
  const int is2 = Int_val(g2) != 0;
  const size_t outn = is2 ? ZK_G2_OUT : ZK_G1_OUT;
  uint8_t *s = dup_bytes(scalars, &ns);
  uint8_t res[ZK_G2_OUT];
  uint64_t h = (uint64_t)Int64_val(handle);
  int rc = 0;
  if (ns == 0 || ns % ZK_FR_BYTES) { free(s); zk_raise_arg("zkb200_table_msm: scalars must be a non-empty multiple of 32 bytes"); }
  caml_release_runtime_system();
  rc = is2 ? zk_g2_table_msm(h, s, ns / ZK_FR_BYTES, res) : zk_g1_table_msm(h, s, ns / ZK_FR_BYTES, res);
  caml_acquire_runtime_system();
  free(s);
  if (rc) zk_raise(rc);
  out = caml_alloc_string(outn);
  memcpy(Bytes_val(out), res, outn);
  CAMLreturn(out);
}

CAMLprim value zkb200_table_free(value h) {
  CAMLparam1(h);
  int rc = zk_table_free((uint64_t)Int64_val(h));
  if (rc) zk_raise(rc);
  CAMLreturn(Val_unit);
}

/* g1_msm : bases:bytes (96 n) -> scalars:bytes (32 n) -> bytes (144)   — Curve.G.dot / apply_powers */
CAMLprim value zkb200_g1_msm(value bases, value scalars) {
  CAMLparam2(bases, scalars);
  CAMLlocal1(out);
  size_t nb, ns;
  uint8_t *b = dup_bytes(bases, &nb), *s = dup_bytes(scalars, &ns);
  uint8_t res[ZK_G1_OUT];
  size_t n = ns / ZK_FR_BYTES;
  int rc = 0;
  if (nb != n * ZK_G1_RAW || ns % ZK_FR_BYTES) { free(b); free(s); zk_raise_arg("zkb200_g1_msm: 96 bytes per base and 32 per scalar expected"); }
  caml_release_runtime_system();
  rc = zk_g1_msm(b, NULL, s, n, res);
  caml_acquire_runtime_system();
  free(b); free(s);
  if (rc) zk_raise(rc);
  out = caml_alloc_string(ZK_G1_OUT);
  memcpy(Bytes_val(out), res, ZK_G1_OUT);
  CAMLreturn(out);
}

CAMLprim value zkb200_g2_msm(value bases, value scalars) {
  CAMLparam2(bases, scalars);
  CAMLlocal1(out);
  size_t nb, ns;
  uint8_t *b = dup_bytes(bases, &nb), *s = dup_bytes(scalars, &ns);
  uint8_t res[ZK_G2_OUT];
  size_t n = ns / ZK_FR_BYTES;
  int rc = 0;
  if (nb != n * ZK_G2_RAW || ns % ZK_FR_BYTES) { free(b); free(s); zk_raise_arg("zkb200_g2_msm: 192 bytes per base and 32 per scalar expected"); }
  caml_release_runtime_system();
  rc = zk_g2_msm(b, NULL, s, n, res);
  caml_acquire_runtime_system();
  free(b); free(s);
  if (rc) zk_raise(rc);
  out = caml_alloc_string(ZK_G2_OUT);
  memcpy(Bytes_val(out), res, ZK_G2_OUT);
  CAMLreturn(out);
}

/* qap_load : v:bytes -> w:bytes -> y:bytes -> target:bytes -> m:int -> n:int -> int64 handle */
CAMLprim value zkb200_qap_load_native(value v, value w, value y, value t, value m, value n) {
  CAMLparam5(v, w, y, t, m);
  CAMLxparam1(n);
  uint8_t *pv = dup_bytes(v, NULL), *pw = dup_bytes(w, NULL), *py = dup_bytes(y, NULL), *pt = dup_bytes(t, NULL);
  uint64_t h = 0;
  caml_release_runtime_system();
  int rc = zk_qap_load(pv, pw, py, pt, (size_t)Long_val(m), (size_t)Long_val(n), &h);
  caml_acquire_runtime_system();
  free(pv); free(pw); free(py); free(pt);
  if (rc) zk_raise(rc);
  CAMLreturn(caml_copy_int64((int64_t)h));
}
CAMLprim value zkb200_qap_load_bytecode(value *argv, int argn) {
  (void)argn;
  return zkb200_qap_load_native(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5]);
}

/* groth16_prove : pk:int64 -> qap:int64 -> sol:bytes -> r:bytes -> s:bytes -> bytes (576) */
CAMLprim value zkb200_groth16_prove(value pk, value qap, value sol, value r, value s) {
  CAMLparam5(pk, qap, sol, r, s);
  CAMLlocal1(out);
  uint8_t *ps = dup_bytes(sol, NULL), *pr = dup_bytes(r, NULL), *pss = dup_bytes(s, NULL);
  uint8_t res[ZK_GROTH16_PROOF_OUT];
  uint64_t hpk = (uint64_t)Int64_val(pk), hq = (uint64_t)Int64_val(qap);
  caml_release_runtime_system();
  int rc = zk_groth16_prove(hpk, hq, ps, pr, pss, res);
  caml_acquire_runtime_system();
  free(ps); free(pr); free(pss);
  if (rc) zk_raise(rc);
  out = caml_alloc_string(ZK_GROTH16_PROOF_OUT);
  memcpy(Bytes_val(out), res, ZK_GROTH16_PROOF_OUT);
  CAMLreturn(out);
}

/* pinocchio_prove : pk:int64 -> qap:int64 -> sol:bytes -> d:bytes (96, or empty for NonZK) -> bytes (1440) */
CAMLprim value zkb200_pinocchio_prove(value pk, value qap, value sol, value d) {
  CAMLparam4(pk, qap, sol, d);
  CAMLlocal1(out);
  size_t dl;
  uint8_t *ps = dup_bytes(sol, NULL), *pd = dup_bytes(d, &dl);
  uint8_t res[ZK_PINOCCHIO_PROOF_OUT];
  uint64_t hpk = (uint64_t)Int64_val(pk), hq = (uint64_t)Int64_val(qap);
  caml_release_runtime_system();
  int rc = zk_pinocchio_prove(hpk, hq, ps, dl == 96 ? pd : NULL, res);
  caml_acquire_runtime_system();
  free(ps); free(pd);
  if (rc) zk_raise(rc);
  out = caml_alloc_string(ZK_PINOCCHIO_PROOF_OUT);
  memcpy(Bytes_val(out), res, ZK_PINOCCHIO_PROOF_OUT);
  CAMLreturn(out);
}
/* ---- proving-key upload ---------------------------------------------------------------------
 * The OCaml side passes the key as an array of Bytes in the field order of the C struct (see
 * ocaml/zkb200.ml: groth16_key_fields / pinocchio_key_fields) plus the dimensions and mid_index as
 * an int array; every field is copied before the runtime lock is released. */
static uint32_t *dup_index(value idx, size_t *len) {
  size_t n = Wosize_val(idx);
  uint32_t *p = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
  for (size_t i = 0; i < n; i++) p[i] = (uint32_t)Long_val(Field(idx, i));
  if (len) *len = n;
  return p;
}

/* groth16_pk_load : n:int -> m:int -> mid_index:int array -> fields:bytes array (9) -> shard:(int*int) -> int64
 * fields = [| a; b1; d1; b2; d2; ti1; ti2; tiztd; ltd_mid |] */
CAMLprim value zkb200_groth16_pk_load(value n, value m, value mid_index, value fields, value shard) {
  CAMLparam5(n, m, mid_index, fields, shard);
  if (Wosize_val(fields) != 9) caml_invalid_argument("zkb200_groth16_pk_load: 9 key fields expected");
  uint8_t *f[9];
  for (int i = 0; i < 9; i++) f[i] = dup_bytes(Field(fields, i), NULL);
  size_t n_mid;
  uint32_t *idx = dup_index(mid_index, &n_mid);
  zk_groth16_pkey pk;
  memset(&pk, 0, sizeof pk);
  pk.n = (size_t)Long_val(n); pk.m = (size_t)Long_val(m); pk.n_mid = n_mid; pk.n_h = 0; pk.mid_index = idx;
  pk.a = f[0]; pk.b1 = f[1]; pk.d1 = f[2]; pk.b2 = f[3]; pk.d2 = f[4];
  pk.ti1 = f[5]; pk.ti2 = f[6]; pk.tiztd = f[7]; pk.ltd_mid = f[8];
  int si = Int_val(Field(shard, 0)), sc = Int_val(Field(shard, 1));
  uint64_t h = 0;
  caml_release_runtime_system();
  int rc = zk_groth16_pk_load(&pk, si, sc, &h);
  caml_acquire_runtime_system();
  for (int i = 0; i < 9; i++) free(f[i]);
  free(idx);
  if (rc) zk_raise(rc);
  CAMLreturn(caml_copy_int64((int64_t)h));
}

/* pinocchio_pk_load : n -> m -> mid_index -> fields:bytes array (20) -> shard -> int64
 * fields = [| vv; yy; vav; yay; bvwy; ww; waw; si; v_all; w_all;
 *             one; vt; yt; vavt; yayt; vbt; wbt; ybt; wt; wawt |] */
CAMLprim value zkb200_pinocchio_pk_load(value n, value m, value mid_index, value fields, value shard) {
  CAMLparam5(n, m, mid_index, fields, shard);
  if (Wosize_val(fields) != 20) caml_invalid_argument("zkb200_pinocchio_pk_load: 20 key fields expected");
  uint8_t *f[20];
  for (int i = 0; i < 20; i++) f[i] = dup_bytes(Field(fields, i), NULL);
  size_t n_mid;
  uint32_t *idx = dup_index(mid_index, &n_mid);
  zk_pinocchio_pkey pk;
  memset(&pk, 0, sizeof pk);
  pk.n = (size_t)Long_val(n); pk.m = (size_t)Long_val(m); pk.n_mid = n_mid; pk.mid_index = idx;
  pk.vv = f[0]; pk.yy = f[1]; pk.vav = f[2]; pk.yay = f[3]; pk.bvwy = f[4]; pk.ww = f[5]; pk.waw = f[6];
  pk.si = f[7]; pk.v_all = f[8]; pk.w_all = f[9]; pk.one = f[10]; pk.vt = f[11]; pk.yt = f[12]; pk.vavt = f[13];
  pk.yayt = f[14]; pk.vbt = f[15]; pk.wbt = f[16]; pk.ybt = f[17]; pk.wt = f[18]; pk.wawt = f[19];
  int si = Int_val(Field(shard, 0)), sc = Int_val(Field(shard, 1));
  uint64_t h = 0;
  caml_release_runtime_system();
  int rc = zk_pinocchio_pk_load(&pk, si, sc, &h);
  caml_acquire_runtime_system();
  for (int i = 0; i < 20; i++) free(f[i]);
  free(idx);
  if (rc) zk_raise(rc);
  CAMLreturn(caml_copy_int64((int64_t)h));
}

CAMLprim value zkb200_qap_free(value h) {
  CAMLparam1(h);
  int rc = zk_qap_free((uint64_t)Int64_val(h));
  if (rc) zk_raise(rc);
  CAMLreturn(Val_unit);
}

CAMLprim value zkb200_key_free(value h) {
  CAMLparam1(h);
  int rc = zk_key_free((uint64_t)Int64_val(h));
  if (rc) zk_raise(rc);
  CAMLreturn(Val_unit);
}

/* g1_sum / g2_sum : bytes (k points, uncompressed) -> bytes (point result) — combining shard partials */
CAMLprim value zkb200_g1_sum(value pts) {
  CAMLparam1(pts);
  CAMLlocal1(out);
  size_t nb;
  uint8_t *b = dup_bytes(pts, &nb);
  uint8_t res[ZK_G1_OUT];
  caml_release_runtime_system();
  int rc = zk_g1_sum(b, nb / ZK_G1_RAW, res);
  caml_acquire_runtime_system();
  free(b);
  if (rc) zk_raise(rc);
  out = caml_alloc_string(ZK_G1_OUT);
  memcpy(Bytes_val(out), res, ZK_G1_OUT);
  CAMLreturn(out);
}

CAMLprim value zkb200_g2_sum(value pts) {
  CAMLparam1(pts);
  CAMLlocal1(out);
  size_t nb;
  uint8_t *b = dup_bytes(pts, &nb);
  uint8_t res[ZK_G2_OUT];
  caml_release_runtime_system();
  int rc = zk_g2_sum(b, nb / ZK_G2_RAW, res);
  caml_acquire_runtime_system();
  free(b);
  if (rc) zk_raise(rc);
  out = caml_alloc_string(ZK_G2_OUT);
  memcpy(Bytes_val(out), res, ZK_G2_OUT);
  CAMLreturn(out);
}

/* ---- verifier side -------------------------------------------------------------------------- */
/* pairing_product : bytes (n G1, uncompressed) -> bytes (n G2) -> bytes (n negate flags, or empty)
 *                   -> bytes (ZK_GT_BYTES).  Replaces Pairing.pairing and the GT sums of
 *                   groth16.ml:163-173 / pinocchio.ml:254-420. */
CAMLprim value zkb200_pairing_product(value g1, value g2, value neg) {
  CAMLparam3(g1, g2, neg);
  CAMLlocal1(out);
  size_t n1, n2, nn;
  uint8_t *a = dup_bytes(g1, &n1);
  uint8_t *b = dup_bytes(g2, &n2);
  uint8_t *f = dup_bytes(neg, &nn);
  size_t n = n1 / ZK_G1_RAW;
  uint8_t res[ZK_GT_BYTES];
  int rc = 0;
  if (n1 % ZK_G1_RAW || n2 != n * ZK_G2_RAW || (nn != 0 && nn != n)) {
    free(a); free(b); free(f);
    zk_raise_arg("zkb200_pairing_product: n G1 points, n G2 points and 0 or n flags expected");
  }
  caml_release_runtime_system();
  rc = zk_pairing_product(a, b, nn ? f : NULL, n, res);
  caml_acquire_runtime_system();
  free(a); free(b); free(f);
  if (rc) zk_raise(rc);
  out = caml_alloc_string(ZK_GT_BYTES);
  memcpy(Bytes_val(out), res, ZK_GT_BYTES);
  CAMLreturn(out);
}

/* gt_mul : bytes -> bytes -> bytes   (GT.( + ) of curve.ml:212-220) */
CAMLprim value zkb200_gt_mul(value x, value y) {
  CAMLparam2(x, y);
  CAMLlocal1(out);
  size_t nx, ny;
  uint8_t *a = dup_bytes(x, &nx);
  uint8_t *b = dup_bytes(y, &ny);
  uint8_t res[ZK_GT_BYTES];
  int rc = 0;
  if (nx != ZK_GT_BYTES || ny != ZK_GT_BYTES) { free(a); free(b); zk_raise_arg("zkb200_gt_mul: two 576-byte GT values expected"); }
  caml_release_runtime_system();
  rc = zk_gt_mul(a, b, res);
  caml_acquire_runtime_system();
  free(a); free(b);
  if (rc) zk_raise(rc);
  out = caml_alloc_string(ZK_GT_BYTES);
  memcpy(Bytes_val(out), res, ZK_GT_BYTES);
  CAMLreturn(out);
}

/* g1_decompress / g2_decompress : bytes (n compressed points) -> bytes (n uncompressed points)
 * (G1/G2.of_compressed_bytes_exn, curve.ml:201,210, over a batch) */
static value decompress_stub(value comp, int g2) {
  CAMLparam1(comp);
  CAMLlocal1(out);
  const size_t cb = g2 ? ZK_G2_COMP : ZK_G1_COMP, rb = g2 ? ZK_G2_RAW : ZK_G1_RAW;
  size_t nb;
  uint8_t *in = dup_bytes(comp, &nb);
  size_t n = nb / cb;
  uint8_t *res = (uint8_t *)malloc(n * rb + 1);
  int rc = 0;
  if (!res) { free(in); caml_failwith("zkb200: out of memory"); }
  if (nb % cb || n == 0) { free(in); free(res); zk_raise_arg("zkb200_decompress: a non-empty multiple of the compressed point size expected"); }
  caml_release_runtime_system();
  rc = g2 ? zk_g2_decompress(in, n, res) : zk_g1_decompress(in, n, res);
  caml_acquire_runtime_system();
  free(in);
  if (rc) { free(res); zk_raise(rc); }
  out = caml_alloc_string(n * rb);
  memcpy(Bytes_val(out), res, n * rb);
  free(res);
  CAMLreturn(out);
}
CAMLprim value zkb200_g1_decompress(value comp) { return decompress_stub(comp, 0); }
CAMLprim value zkb200_g2_decompress(value comp) { return decompress_stub(comp, 1); }
