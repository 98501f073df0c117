"""QAP.eval, Groth16 and Pinocchio prove through the C ABI against the oracle's restatement of
the reference (QAP.ml:120-135, groth16.ml:123-161, pinocchio.ml:210-248,427-514).

Bit-exact: compressed proof bytes must equal the oracle's, the committed golden vectors and
the closed-form trapdoor identity; every proof is also replayed through the oracle's copy of
the reference verifier (the reference's own acceptance test, test.ml:178)."""
import ctypes
import json
import os
import random

import pytest

from oracle import bls12_381 as O
from oracle import zk as Z
from tests import helpers as H

pytestmark = pytest.mark.gpu
R = O.R
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))


def _circuits():
    return {"cubic": (Z.circuit_cubic(), 3), "mulchain8": (Z.circuit_mulchain(8), 5),
            "pair_case7": (Z.circuit_pair_case(7), 11), "mulchain33": (Z.circuit_mulchain(33), 9)}


@pytest.mark.parametrize("name", ["cubic", "mulchain8", "pair_case7", "mulchain33"])
def test_qap_eval_matches_reference(zk, name):
    from zukelang_b200 import qap as Q
    (circ, wit), x = _circuits()[name]
    sol = wit(x)
    oq = Z.qap_build(circ.gates)
    _p, h_ref = Z.qap_eval(sol, oq)
    q = H.mirror_qap(oq)
    _none, h, (V, W, Y) = Q.eval_full(sol, q)
    assert h == h_ref
    assert V == Z.qap_combine(sol, oq.v) and W == Z.qap_combine(sol, oq.w) and Y == Z.qap_combine(sol, oq.y)
    # a witness that violates a gate trips the reference's assert (QAP.ml:134)
    bad = dict(sol)
    k = circ.mids[0]
    bad[k] = (bad[k] + 1) % R
    with pytest.raises(AssertionError):
        Q.eval(bad, q)
    # a missing variable is `assert false` in var.ml:72-78
    with pytest.raises(AssertionError):
        Q.eval({kk: v for kk, v in sol.items() if kk != k}, q)
    q.free()


def test_fr_quotient_golden(zk):
    from zukelang_b200 import _lib
    from zukelang_b200.curve import fr_vector
    for vec in GOLD["quotient"]:
        n = vec["n"]
        ints = lambda xs: [int(x) for x in xs]
        out = (ctypes.c_uint8 * (32 * (n - 1)))()
        _lib.check(zk.zk_fr_quotient(fr_vector(ints(vec["V"])), fr_vector(ints(vec["W"])), fr_vector(ints(vec["Y"])),
                                     fr_vector(ints(vec["T"])), n, out))
        got = [int.from_bytes(bytes(out)[i:i + 32], "little") for i in range(0, 32 * (n - 1), 32)]
        assert got == ints(vec["h"]), vec["name"]
        # not divisible -> ZK_EREMAINDER
        Ybad = ints(vec["Y"])
        Ybad[0] = (Ybad[0] + 1) % R
        rc = zk.zk_fr_quotient(fr_vector(ints(vec["V"])), fr_vector(ints(vec["W"])), fr_vector(Ybad),
                               fr_vector(ints(vec["T"])), n, out)
        assert rc == _lib.ZK_EREMAINDER


def test_groth16_config1_golden_and_verifier(zk):
    """Config 1: README x*x*x + x + 3 (test.ml:194-197), Groth16."""
    from zukelang_b200 import groth16 as G16
    g = GOLD["groth16_config1"]
    circ, wit = Z.circuit_cubic()
    td = Z.Groth16Trapdoor(*[int(v) for v in g["trapdoor"]])
    x, r, s = int(g["x"]), int(g["r"]), int(g["s"])
    sol = wit(x)
    oq = Z.qap_build(circ.gates, literal=True)
    opk, ovk = Z.groth16_keygen(td, circ, oq)
    P = G16.Make()
    q = H.mirror_qap(oq)
    pk = H.mirror_groth16_pkey(opk)
    proof = P.prove_with(r, s, q, pk, sol)
    assert proof.to_compressed_bytes().hex() == g["proof_compressed"]
    oproof = Z.groth16_prove(r, s, oq, opk, sol)                  # literal reference algorithm
    assert proof.to_compressed_bytes() == H.groth16_proof_compressed(oproof)
    assert H.decode_groth16_proof(proof) == Z.groth16_closed_form(td, r, s, oq, circ, sol)
    # the reference's acceptance test: verify accepts (test.ml:178) — and rejects a wrong public input
    pub = {k: sol[k] for k in ovk.ltgm_io}
    assert Z.groth16_verify(pub, ovk, H.decode_groth16_proof(proof))
    # prove rng ... draws r then s (groth16.ml:124-125)
    rng = random.Random(5)
    r2, s2 = random.Random(5).randrange(R), None
    rr = random.Random(5); r2 = rr.randrange(R); s2 = rr.randrange(R)
    assert P.prove(rng, q, pk, sol) == P.prove_with(r2, s2, q, pk, sol)
    P.free(pk)
    q.free()


def test_groth16_keygen_matches_oracle(zk):
    """Mirror keygen (trapdoor scalars on the host, every point from the fixed-base kernel)
    against the oracle's restatement of groth16.ml:45-108."""
    from zukelang_b200 import groth16 as G16
    circ, wit = Z.circuit_mulchain(8)
    oq = Z.qap_build(circ.gates)
    rng = random.Random(77)
    td = Z.Groth16Trapdoor(*[random.Random(77).randrange(R)] * 5)
    rr = random.Random(77)
    td = Z.Groth16Trapdoor(*[rr.randrange(R) for _ in range(5)])
    opk, ovk = Z.groth16_keygen(td, circ, oq, with_ab=False)
    pk, vk = G16.Make().keygen(rng, H.mirror_circuit(circ), H.mirror_qap(oq))
    exp = H.mirror_groth16_pkey(opk)
    for f in ("a", "d1", "ti1", "ltd_mid", "tiztd", "b1", "b2", "d2", "ti2"):
        assert getattr(pk, f) == getattr(exp, f), f
    assert vk.ltgm_io == {k: H.mirror_point(v) for k, v in ovk.ltgm_io.items()}
    assert vk.gm == H.mirror_point(ovk.gm, True) and vk.d == H.mirror_point(ovk.d, True)


@pytest.mark.parametrize("n", [16, 100])
def test_groth16_mulchain_closed_form(zk, n):
    """Larger circuits: GPU keygen + prove, checked against the single-scalar trapdoor identity."""
    from zukelang_b200 import groth16 as G16
    circ, wit = Z.circuit_mulchain(n)
    oq = Z.qap_build(circ.gates)
    rr = random.Random(n)
    td = Z.Groth16Trapdoor(*[rr.randrange(R) for _ in range(5)])
    P = G16.Make()
    q = H.mirror_qap(oq)
    pk, _vk = P.keygen(random.Random(n), H.mirror_circuit(circ), q)
    for trial in range(2):
        sol = wit(rr.randrange(R))
        r, s = rr.randrange(R), rr.randrange(R)
        proof = P.prove_with(r, s, q, pk, sol)
        assert H.decode_groth16_proof(proof) == Z.groth16_closed_form(td, r, s, oq, circ, sol)
    P.free(pk)
    q.free()


def test_pinocchio_small_golden_and_verifier(zk):
    from zukelang_b200 import pinocchio as PN
    g = GOLD["pinocchio_small"]
    circ, wit = Z.circuit_pair_case(7)
    td = Z.PinocchioTrapdoor(*[int(v) for v in g["trapdoor"]])
    d = tuple(int(v) for v in g["d"])
    sol = wit(g["witness_seed"])
    oq = Z.qap_build(circ.gates, literal=True)
    opk, ovk = Z.pinocchio_keygen(td, circ, oq)
    M = PN.Make()
    q = H.mirror_qap(oq)
    pk = H.mirror_pinocchio_pkey(opk)
    ios = {k: sol[k] for k in ovk["vv_io"]}
    p_nonzk = M.NonZK.prove(None, q, pk, sol)
    assert p_nonzk.to_compressed_bytes().hex() == g["nonzk"]
    p_zk = M.ZK.prove_with(d, q, pk, sol)
    assert p_zk.to_compressed_bytes().hex() == g["zk"]
    assert p_zk.to_compressed_bytes() == H.pinocchio_proof_compressed(Z.pinocchio_prove(oq, opk, sol, d))
    assert Z.pinocchio_verify(ios, ovk, H.decode_pinocchio_proof(p_zk))
    # ZK.prove draws dv, dw, dy in that order (pinocchio.ml:428-430)
    rr = random.Random(9)
    dd = (rr.randrange(R), rr.randrange(R), rr.randrange(R))
    assert M.ZK.prove(random.Random(9), q, pk, sol) == M.ZK.prove_with(dd, q, pk, sol)
    # mirror keygen against the oracle's restatement of pinocchio.ml:77-189
    rk = random.Random(31)
    td2 = Z.PinocchioTrapdoor(*[random.Random(31).randrange(R)] * 8)
    r3 = random.Random(31)
    td2 = Z.PinocchioTrapdoor(*[r3.randrange(R) for _ in range(8)])
    opk2, ovk2 = Z.pinocchio_keygen(td2, circ, oq)
    pk2, vk2 = M.ZK.keygen(rk, H.mirror_circuit(circ), q)
    exp = H.mirror_pinocchio_pkey(opk2)
    for f in opk2:
        assert getattr(pk2, f) == getattr(exp, f), f
    for f in ("av", "aw", "ay", "gm2", "bgm", "bgm2", "yt"):
        assert vk2[f] == H.mirror_point(ovk2[f], f in ("av", "ay", "gm2", "bgm2", "yt")), f
    M.ZK.free(pk)
    q.free()


def test_pinocchio_config2_zk_1024(zk):
    """Config 2: Pinocchio ZK on a synthetic 2^10-gate pair/case-shaped circuit.  All eight
    elements against the closed-form trapdoor identity (exact at this size)."""
    from zukelang_b200 import pinocchio as PN
    n = 1024
    circ, wit = Z.circuit_pair_case(n)
    oq = Z.qap_build(circ.gates)
    assert Z.poly_degree(oq.target) == n
    rr = random.Random(0x434F4E32)
    td = Z.PinocchioTrapdoor(*[rr.randrange(R) for _ in range(8)])
    M = PN.Make()
    q = H.mirror_qap(oq)
    pk, _vk = M.ZK.keygen(random.Random(0x434F4E32), H.mirror_circuit(circ), q)
    sol = wit(2024)
    d = (rr.randrange(R), rr.randrange(R), rr.randrange(R))
    proof = M.ZK.prove_with(d, q, pk, sol)
    assert H.decode_pinocchio_proof(proof) == Z.pinocchio_closed_form(td, oq, circ, sol, d)
    p0 = M.NonZK.prove(None, q, pk, sol)
    assert H.decode_pinocchio_proof(p0) == Z.pinocchio_closed_form(td, oq, circ, sol, None)
    M.ZK.free(pk)
    q.free()


def test_curve_mirror_semantics(zk):
    """Curve.G surface: dot's domain check, apply_powers' Invalid_argument, sum_map, powers."""
    from zukelang_b200 import _lib
    from zukelang_b200.curve import Bls12_381 as C
    G1 = C.G1
    p3 = G1.of_Fr(3)
    assert p3 == H.mirror_point(O.G1.mul(O.G1.one, 3))
    assert G1.add(p3, G1.neg(p3)) == G1.zero
    assert G1.mul(G1.one, 12345) == G1.of_Fr(12345)                    # curve.ml:232
    a, b, c, d = 17, 23, 101, 977
    assert G1.of_Fr(a * b + c * d) == G1.add(G1.of_Fr(a * b), G1.of_Fr(c * d))   # curve.ml:236-239
    pw = G1.powers(4, 5)
    assert pw == [H.mirror_point(O.G1.mul(O.G1.one, 5 ** i)) for i in range(5)]
    assert G1.apply_powers([1, 2, 3], pw) == G1.of_Fr(1 + 2 * 5 + 3 * 25)
    assert G1.apply_powers([], pw) == G1.zero
    with pytest.raises(_lib.InvalidArgument):
        G1.apply_powers([1] * 6, pw)                                    # curve.ml:116
    m = {("a", 1): pw[1], ("b", 2): pw[2]}
    assert G1.dot(m, {("a", 1): 2, ("b", 2): 3}) == G1.of_Fr(2 * 5 + 3 * 25)
    with pytest.raises(AssertionError):
        G1.dot(m, {("a", 1): 2})                                        # curve.ml:96-100
    assert G1.sum_map(m, lambda k, v: v) == G1.of_Fr(30)
    assert G1.to_compressed_bytes(G1.one) == O.G1_GEN_COMPRESSED
    assert C.G2.to_compressed_bytes(C.G2.one) == O.G2_GEN_COMPRESSED
    assert C.G2.add(C.G2.of_Fr(2), C.G2.of_Fr(3)) == C.G2.of_Fr(5)


@pytest.mark.parametrize("shards", [2, 3])
def test_sharded_prove_equals_unsharded(zk, shards):
    """key handles loaded as shard i of N return partial sums whose sum is the proof
    (N emulated in one process; the N > 1 transport is covered by tests/test_cpu_dist.py)."""
    import ctypes
    from zukelang_b200 import _lib, dist as D, groth16 as G16, pinocchio as PN
    from zukelang_b200.curve import Fr, fr_vector
    circ, wit = Z.circuit_mulchain(21)
    oq = Z.qap_build(circ.gates)
    q = H.mirror_qap(oq)
    sol = wit(12345)
    keys = q.variables()
    # Groth16
    whole = G16.Make()
    pk, _ = whole.keygen(random.Random(3), H.mirror_circuit(circ), q)
    r, s = 1111, 2222
    ref = whole.prove_with(r, s, q, pk, sol)
    parts = []
    for i in range(shards):
        P = G16.Make(shard=(i, shards))
        out = (ctypes.c_uint8 * _lib.GROTH16_PROOF_OUT)()
        _lib.check(zk.zk_groth16_prove(P._key_handle(pk, q), q.handle(), fr_vector(sol[k] for k in keys),
                                       Fr.to_bytes(r), Fr.to_bytes(s), out))
        parts.append(bytes(out))
    comb = D.combine_groth16(parts)
    assert comb[96:144] + comb[336:432] + comb[528:576] == ref.to_compressed_bytes()
    G16.Make.free(pk)
    # Pinocchio ZK
    circ2, wit2 = Z.circuit_pair_case(10)
    oq2 = Z.qap_build(circ2.gates)
    q2 = H.mirror_qap(oq2)
    sol2 = wit2(5)
    keys2 = q2.variables()
    M = PN.Make()
    pk2, _ = M.ZK.keygen(random.Random(8), H.mirror_circuit(circ2), q2)
    d = (31, 41, 59)
    ref2 = M.ZK.prove_with(d, q2, pk2, sol2)
    parts2 = []
    for i in range(shards):
        S = PN.ZK(shard=(i, shards))
        out = (ctypes.c_uint8 * _lib.PINOCCHIO_PROOF_OUT)()
        _lib.check(zk.zk_pinocchio_prove(S._key_handle(pk2, q2), q2.handle(), fr_vector(sol2[k] for k in keys2),
                                         fr_vector(d), out))
        parts2.append(bytes(out))
    comb2 = D.combine_pinocchio(parts2)
    got, o = b"", 0
    for is2 in PN.PROOF_IS_G2:
        raw, comp = (192, 96) if is2 else (96, 48)
        got += comb2[o + raw:o + raw + comp]
        o += raw + comp
    assert got == ref2.to_compressed_bytes()
    M.ZK.free(pk2)
    q.free()
    q2.free()
