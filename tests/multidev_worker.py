"""Worker of tests/test_gpu_multidev.py (its own process: the device list of libzkb200 is fixed
between zk_init_devices and zk_shutdown).  Drives every visible GPU (up to 8; one is fine) from ONE
process through the C ABI and checks that the finished results equal the single-device ones and the
oracle's closed forms.  Prints one JSON line."""
import ctypes
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import bls12_381 as O            # noqa: E402  (test infrastructure)
from oracle import zk as Z                   # noqa: E402
from zukelang_b200 import _lib, sparse as S  # noqa: E402
from zukelang_b200.curve import Fr, fr_vector  # noqa: E402

R = O.R


def visible_devices():
    import torch
    return torch.cuda.device_count()


def table_check(zk, group, n, rng):
    """zk_g*_table_msm and _batch on a table with known discrete logs: exact expected point."""
    raw, outn = (96, 144) if group == "g1" else (192, 288)
    dl = [rng.randrange(1, R) for _ in range(n)]
    bases = (ctypes.c_uint8 * (raw * n))()
    _lib.check(getattr(zk, "zk_%s_fixed_base_mul" % group)(fr_vector(dl), n, bases))
    h = ctypes.c_uint64()
    _lib.check(getattr(zk, "zk_%s_table_load" % group)(bases, None, n, 1, 0, ctypes.byref(h)))
    info = (ctypes.c_uint64 * 8)()
    _lib.check(zk.zk_table_info(h.value, info))
    results = []
    vecs = []
    for _ in range(3):
        ks = [rng.randrange(R) for _ in range(n)]
        ks[0], ks[1], ks[2] = 0, 1, R - 1
        tot = sum(a * b for a, b in zip(ks, dl)) % R
        exp = (ctypes.c_uint8 * raw)()
        _lib.check(getattr(zk, "zk_%s_fixed_base_mul" % group)(Fr.to_bytes(tot), 1, exp))
        vecs.append((fr_vector(ks), bytes(exp)))
    out = (ctypes.c_uint8 * outn)()
    _lib.check(getattr(zk, "zk_%s_table_msm" % group)(h.value, vecs[0][0], n, out))
    results.append(bytes(out)[:raw] == vecs[0][1])
    # a prefix of the table (n_used < n): only some parts are active
    nu = n // 3
    tot = sum(a * b for a, b in zip([int.from_bytes(vecs[1][0][32 * i:32 * i + 32], "little") for i in range(nu)], dl)) % R
    exp = (ctypes.c_uint8 * raw)()
    _lib.check(getattr(zk, "zk_%s_fixed_base_mul" % group)(Fr.to_bytes(tot), 1, exp))
    _lib.check(getattr(zk, "zk_%s_table_msm" % group)(h.value, vecs[1][0], nu, out))
    results.append(bytes(out)[:raw] == bytes(exp))
    bufs = [ctypes.create_string_buffer(v, len(v)) for v, _ in vecs]
    ptrs = (ctypes.c_void_p * len(bufs))(*[ctypes.addressof(b) for b in bufs])
    outs = (ctypes.c_uint8 * (outn * len(bufs)))()
    _lib.check(getattr(zk, "zk_%s_table_msm_batch" % group)(h.value, ptrs, n, len(bufs), outs))
    for i, (_, e) in enumerate(vecs):
        results.append(bytes(outs)[i * outn:i * outn + raw] == e)
    # a non-canonical scalar must be reported, not reduced silently
    bad = bytearray(vecs[0][0])
    bad[32 * (n - 1):32 * n] = (R + 5).to_bytes(32, "little")
    rc = getattr(zk, "zk_%s_table_msm" % group)(h.value, bytes(bad), n, out)
    results.append(rc == _lib.ZK_EPOINT)
    # device-pointer entry points need a single-device table
    if zk.zk_device_count() > 1 and n >= 4096 * zk.zk_device_count():
        results.append(zk.zk_table_join(h.value, None) == _lib.ZK_EARG)
    _lib.check(zk.zk_table_free(h.value))
    return all(results), int(info[5])


def prove_bytes(zk, circ, sc, td, sol, r, s):
    dom = S.EvalDomain(sc)
    P = S.Groth16Sparse()
    pk, vk = P.keygen_from_trapdoor(td, sc, dom.w)
    proof = P.prove_with(r, s, dom, pk, sol)
    bad = dict(sol)
    bad[circ.mids[3]] = (bad[circ.mids[3]] + 1) % R
    try:
        P.prove_with(r, s, dom, pk, bad)
        rejects = False
    except AssertionError:
        rejects = True
    proof2 = P.prove_with(r, s, dom, pk, sol)          # the handle still works after a failed prove
    ok_verify = P.verify({k: sol[k] for k in vk.ltgm_io}, vk, proof)
    P.free(pk)
    dom.free()
    return proof.to_compressed_bytes(), rejects and proof2.to_compressed_bytes() == proof.to_compressed_bytes(), ok_verify


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    ndev = min(visible_devices(), 8)
    rng = random.Random(0x4D44 + ndev)
    circ, wit = Z.circuit_random_r1cs(1 << logn)
    sc = S.SparseCircuit([(dict(g.lhs), dict(g.l), dict(g.r)) for g in circ.gates], circ.inputs_public, circ.outputs, circ.mids)
    td = tuple(rng.randrange(1, R) for _ in range(5))
    sol = wit(rng.randrange(R))
    r, s = rng.randrange(R), rng.randrange(R)
    A, B, C = Z.groth16_closed_form_scalars(Z.Groth16Trapdoor(*td), r, s, circ, sol)
    expect = O.g1_compress(O.G1.of_Fr(A)) + O.g2_compress(O.G2.of_Fr(B)) + O.g1_compress(O.G1.of_Fr(C))

    zk = _lib.init_devices(list(range(ndev)))
    rec = {"devices": ndev, "device_count": zk.zk_device_count(), "log_n": logn}
    multi, rec["multi_rejects_bad_witness_then_recovers"], rec["multi_verifies"] = prove_bytes(zk, circ, sc, td, sol, r, s)
    rec["multi_equals_oracle"] = multi == expect
    rec["g1_table_ok"], rec["g1_table_bytes"] = table_check(zk, "g1", 1 << 16, rng)
    rec["g2_table_ok"], _ = table_check(zk, "g2", 1 << 15, rng)
    rec["small_table_ok"], _ = table_check(zk, "g1", 300, rng)          # stays on the primary device
    # Pinocchio keys stay on the primary device in a multi-device process: the prover still works
    _lib.check(zk.zk_shutdown())
    _lib.check(zk.zk_init(0))
    single, _, _ = prove_bytes(zk, circ, sc, td, sol, r, s)
    rec["multi_equals_single_device"] = multi == single
    rec["ok"] = all(v for k, v in rec.items() if isinstance(v, bool))
    print(json.dumps(rec), flush=True)
    sys.exit(0 if rec["ok"] else 1)


if __name__ == "__main__":
    main()
