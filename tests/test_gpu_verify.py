"""Verifier side and wire formats through the C ABI: zk_pairing_product,
zk_gt_mul, zk_g*_decompress, Groth16 / Pinocchio ``verify`` (groth16.ml:163-173,
pinocchio.ml:254-420) and the yojson encodings of keys and proofs.

Oracle: oracle/pairing.py (ate pairing in Python integers).  The device pairing is the CUBE of the
oracle's reduced pairing expressed in another basis of Fp12 (zukelang_b200/csrc/pairing.cuh);
helpers.gt_bytes_to_oracle converts, and the comparison is exact."""
import ctypes
import random

import pytest

from oracle import bls12_381 as O
from oracle import pairing as OP
from oracle import zk as Z
from tests import helpers as H
from zukelang_b200 import _lib

pytestmark = pytest.mark.gpu
R, P = O.R, O.P


def _pairing_product(zk, pairs, negate=None):
    out = (ctypes.c_uint8 * 576)()
    g1 = b"".join(O.g1_to_uncompressed(p) for p, _ in pairs)
    g2 = b"".join(O.g2_to_uncompressed(q) for _, q in pairs)
    neg = bytes(negate) if negate is not None else None
    rc = zk.zk_pairing_product(g1, g2, neg, len(pairs), out)
    return rc, bytes(out)


def test_pairing_matches_the_oracle(zk):
    rng = random.Random(41)
    a, b = rng.randrange(1, R), rng.randrange(1, R)
    p, q = O.G1.mul(O.G1.one, a), O.G2.mul(O.G2.one, b)
    rc, gt = _pairing_product(zk, [(p, q)])
    assert rc == 0
    e = H.gt_bytes_to_oracle(gt)
    assert e == OP.f12_pow(OP.pairing(p, q), 3)                       # exact, up to the documented cube
    # bilinearity against the generator pairing, computed by the device and raised by the oracle
    rc, gt_gen = _pairing_product(zk, [(O.G1.one, O.G2.one)])
    assert rc == 0
    assert e == OP.f12_pow(H.gt_bytes_to_oracle(gt_gen), a * b % R)
    assert H.oracle_to_gt_bytes(e) == gt


def test_pairing_products_negation_identity_and_gt_mul(zk):
    rng = random.Random(42)
    a, b, c = (rng.randrange(1, R) for _ in range(3))
    G1, G2 = O.G1, O.G2
    one = (1).to_bytes(48, "big") + bytes(528)
    # e(aG, bH) * e(-abG, H) = 1, with the negation done by the flag and by the caller
    rc, gt = _pairing_product(zk, [(G1.mul(G1.one, a), G2.mul(G2.one, b)), (G1.mul(G1.one, a * b % R), G2.one)], [0, 1])
    assert rc == 0 and gt == one
    rc, gt = _pairing_product(zk, [(G1.mul(G1.one, a), G2.mul(G2.one, b)), (G1.neg(G1.mul(G1.one, a * b % R)), G2.one)])
    assert rc == 0 and gt == one
    # pairs with the identity contribute nothing
    rc, e1 = _pairing_product(zk, [(G1.mul(G1.one, c), G2.one)])
    rc2, e2 = _pairing_product(zk, [(None, G2.one), (G1.mul(G1.one, c), G2.one), (G1.one, None)])
    assert rc == 0 and rc2 == 0 and e1 == e2 and e1 != one
    rc, e0 = _pairing_product(zk, [(None, None)])
    assert rc == 0 and e0 == one
    # GT product: e(cG, H) * e(aG, H) = e((a + c) G, H)
    rc, ea = _pairing_product(zk, [(G1.mul(G1.one, a), G2.one)])
    rc, eac = _pairing_product(zk, [(G1.mul(G1.one, (a + c) % R), G2.one)])
    out = (ctypes.c_uint8 * 576)()
    assert zk.zk_gt_mul(e1, ea, out) == 0 and bytes(out) == eac
    assert zk.zk_gt_mul(e1, one, out) == 0 and bytes(out) == e1
    bad = bytes([0xff]) * 48 + bytes(528)                              # coefficient >= p
    assert zk.zk_gt_mul(bad, one, out) == _lib.ZK_EPOINT


def _off_subgroup_g1():
    x = 1
    while True:
        y2 = (x ** 3 + 4) % P
        y = pow(y2, (P + 1) // 4, P)
        if y * y % P == y2 and O.G1.add(O.G1.mul((x, y), R - 1), (x, y)) is not None:
            return (x, y)
        x += 1


def test_pairing_rejects_bad_points(zk):
    out = (ctypes.c_uint8 * 576)()
    g1, g2 = O.g1_to_uncompressed(O.G1.one), O.g2_to_uncompressed(O.G2.one)
    off_curve = g1[:95] + bytes([g1[95] ^ 1])
    assert zk.zk_pairing_product(off_curve, g2, None, 1, out) == _lib.ZK_EPOINT
    assert zk.zk_pairing_product(O.g1_to_uncompressed(_off_subgroup_g1()), g2, None, 1, out) == _lib.ZK_EPOINT
    assert b"subgroup" in zk.zk_last_error()
    assert zk.zk_pairing_product(g1, g2, None, 0, out) == _lib.ZK_EARG


@pytest.mark.parametrize("G,comp,unc,raw", [(O.G1, O.g1_compress, O.g1_to_uncompressed, 96),
                                            (O.G2, O.g2_compress, O.g2_to_uncompressed, 192)], ids=["G1", "G2"])
def test_decompress_matches_the_oracle(zk, G, comp, unc, raw):
    fn = zk.zk_g1_decompress if G is O.G1 else zk.zk_g2_decompress
    rng = random.Random(43)
    pts = [None, G.one, G.neg(G.one)] + [G.mul(G.one, rng.randrange(1, R)) for _ in range(29)]
    pts += [G.neg(p) for p in pts[3:8]]                                # both values of the sign flag
    blob = b"".join(comp(p) for p in pts)
    out = (ctypes.c_uint8 * (raw * len(pts)))()
    assert fn(blob, len(pts), out) == 0
    assert bytes(out) == b"".join(unc(p) for p in pts)
    c = raw // 2
    one = comp(G.one)
    o1 = (ctypes.c_uint8 * raw)()
    cases = {
        "uncompressed flag": bytes([one[0] & 0x7f]) + one[1:],
        "identity with payload": bytes([0xc0]) + bytes(c - 2) + b"\x01",
        "identity with sign": bytes([0xe0]) + bytes(c - 1),
        "x >= p": bytes([0x9f]) + b"\xff" * (c - 1),
    }
    for name, b in cases.items():
        assert fn(b, 1, o1) == _lib.ZK_EPOINT, name
    # an x that is not on the curve: walk x upwards from the generator's until decompression fails
    tail = int.from_bytes(one[-8:], "big")
    seen = set()
    for d in range(1, 40):
        b = one[:-8] + (tail + d).to_bytes(8, "big")
        rc = fn(b, 1, o1)
        assert rc == _lib.ZK_EPOINT                                    # no curve point, or outside the subgroup
        seen.add(zk.zk_last_error())
    assert any(b"not the abscissa" in m for m in seen) and any(b"subgroup" in m for m in seen)
    assert fn(one, 0, o1) == _lib.ZK_EARG


def test_groth16_keygen_prove_verify_on_the_device(zk):
    """Protocol.S round trip with no oracle in the loop (test.ml:170-178), then cross-checks of the
    device verifier against oracle-made keys and proofs."""
    from zukelang_b200 import groth16 as G16
    from zukelang_b200.curve import Bls12_381 as C
    circ, wit = Z.circuit_mulchain(8)
    oq = Z.qap_build(circ.gates)
    Pr = G16.Make()
    q = H.mirror_qap(oq)
    rng = random.Random(51)
    pk, vk = Pr.keygen(rng, H.mirror_circuit(circ), q)
    sol = wit(rng.randrange(R))
    proof = Pr.prove(rng, q, pk, sol)
    pub = {k: sol[k] for k in vk.ltgm_io}
    assert Pr.verify(pub, vk, proof) is True
    k0 = sorted(pub)[-1]
    assert Pr.verify({**pub, k0: (pub[k0] + 1) % R}, vk, proof) is False
    assert Pr.verify(pub, vk, G16.Proof(proof.a, proof.b, C.G1.add(proof.c, C.G1.one))) is False
    with pytest.raises(AssertionError):
        Pr.verify({k: v for k, v in pub.items() if k != k0}, vk, proof)     # Domain mismatch (curve.ml:96-100)
    # vkey.ab is e(alpha, beta): same GT value as the oracle's (cubed, other basis)
    td = None
    rr = random.Random(51)
    td = Z.Groth16Trapdoor(*[rr.randrange(R) for _ in range(5)])
    opk, ovk = Z.groth16_keygen(td, circ, oq)
    assert H.gt_bytes_to_oracle(vk.ab.raw) == OP.f12_pow(ovk.ab, 3)
    # an oracle-made proof (literal reference prover) is accepted by the device verifier
    r, s = rr.randrange(R), rr.randrange(R)
    oproof = Z.groth16_prove(r, s, oq, opk, sol)
    mp = G16.Proof(H.mirror_point(oproof[0]), H.mirror_point(oproof[1], True), H.mirror_point(oproof[2]))
    assert Pr.verify(pub, vk, mp) is True
    # and the oracle's verifier agrees with the device's on both outcomes
    assert Z.groth16_verify(pub, ovk, H.decode_groth16_proof(proof)) is True
    Pr.free(pk)
    q.free()


def test_pinocchio_verify_on_the_device(zk):
    from zukelang_b200 import pinocchio as PN
    from zukelang_b200.curve import Bls12_381 as C
    circ, wit = Z.circuit_pair_case(7)
    oq = Z.qap_build(circ.gates)
    M = PN.Make()
    q = H.mirror_qap(oq)
    rng = random.Random(61)
    pk, vk = M.ZK.keygen(rng, H.mirror_circuit(circ), q)
    sol = wit(5)
    ios = {k: sol[k] for k in vk["vv_io"]}
    for prover in (M.NonZK, M.ZK):
        proof = prover.prove(rng, q, pk, sol)
        assert prover.verify(ios, vk, proof) is True
        k0 = sorted(ios)[-1]
        assert prover.verify({**ios, k0: (ios[k0] + 1) % R}, vk, proof) is False   # divisibility check fails
        bad = PN.Proof(**{f: getattr(proof, f) for f in PN.PROOF_FIELDS})
        bad.vavv = C.G1.add(bad.vavv, C.G1.one)
        with pytest.raises(AssertionError):                                       # KC check is an assert (:296)
            prover.verify(ios, vk, bad)
        bad = PN.Proof(**{f: getattr(proof, f) for f in PN.PROOF_FIELDS})
        bad.h = C.G1.add(bad.h, C.G1.one)
        assert prover.verify(ios, vk, bad) is False
    M.ZK.free(pk)
    q.free()


@pytest.mark.parametrize("which", ["groth16", "pinocchio_nonzk", "pinocchio_zk"])
def test_protocol_harness_round_trip(zk, which):
    """``Test.Make(F)(Protocol)`` behind the front-end (test.ml:119-178): keygen, prove, verify on the
    public part of the witness — the reference's own acceptance test, no oracle in the loop."""
    from zukelang_b200 import groth16 as G16, pinocchio as PN
    from zukelang_b200.protocol import Test
    proto = {"groth16": G16.Make(), "pinocchio_nonzk": PN.Make().NonZK, "pinocchio_zk": PN.Make().ZK}[which]
    rng = random.Random(81)
    for (circ, wit), x in ((Z.circuit_cubic(), 10), (Z.circuit_pair_case(5), 6), (Z.circuit_mulchain(20), 3)):
        oq = Z.qap_build(circ.gates)
        q = H.mirror_qap(oq)
        Test(proto).run(rng, H.mirror_circuit(circ), q, wit(x))
        q.free()
    # a witness that violates a gate never reaches the verifier: QAP.eval asserts (QAP.ml:134)
    circ, wit = Z.circuit_cubic()
    q = H.mirror_qap(Z.qap_build(circ.gates))
    sol = wit(10)
    k = sorted(set(circ.mids))[0]
    with pytest.raises(AssertionError):
        Test(proto).run(rng, H.mirror_circuit(circ), q, {**sol, k: (sol[k] + 1) % R})
    q.free()


def test_batched_pairing_products(zk):
    from zukelang_b200.curve import Bls12_381 as C
    g1 = C.G1.fixed_base([3, 5, 15, 7])
    g2 = C.G2.fixed_base([5, 3, 1, 11])
    e = C.Pairing.product
    groups = [([(g1[0], g2[0]), (g1[2], g2[2])], [False, True]),        # e(3G,5H) / e(15G,H) = 1
              ([], None),
              ([(g1[3], g2[3])], None),
              ([(g1[0], g2[0]), (g1[1], g2[1])], [False, True])]         # e(3G,5H) / e(5G,3H) = 1
    got = C.Pairing.products(groups)
    assert got[0] == C.GT.zero and got[1] == C.GT.zero and got[3] == C.GT.zero
    assert got[2] == e([(g1[3], g2[3])]) and got[2] != C.GT.zero
    assert [x for x in got] == [e(p, n) for p, n in groups]


def test_wire_formats_round_trip(zk):
    from zukelang_b200 import groth16 as G16, pinocchio as PN, wire
    circ, wit = Z.circuit_cubic()
    oq = Z.qap_build(circ.gates, literal=True)
    q = H.mirror_qap(oq)
    rng = random.Random(71)
    Pr = G16.Make()
    pk, vk = Pr.keygen(rng, H.mirror_circuit(circ), q)
    sol = wit(9)
    proof = Pr.prove(rng, q, pk, sol)
    jp = wire.Groth16Wire.yojson_of_proof(proof)
    # record -> object in declaration order, points as raw compressed bytes (escaped like Yojson)
    assert wire.loads(jp) == {b"a": proof.a._comp, b"b": proof.b._comp, b"c": proof.c._comp}
    assert jp.startswith(b'{"a":"') and list(wire.loads(jp)) == [b"a", b"b", b"c"]
    assert wire.Groth16Wire.proof_of_yojson(jp) == proof
    pk2 = wire.Groth16Wire.pkey_of_yojson(wire.Groth16Wire.yojson_of_pkey(pk))
    for f in ("a", "d1", "ti1", "ltd_mid", "tiztd", "b1", "b2", "d2", "ti2"):
        assert getattr(pk2, f) == getattr(pk, f), f
    jv = wire.Groth16Wire.yojson_of_vkey(vk)
    vk2 = wire.Groth16Wire.vkey_of_yojson(jv)
    assert vk2 == vk
    vk3 = wire.Groth16Wire.vkey_of_yojson(jv, recompute_ab_from=pk2)
    assert vk3.ab == vk.ab
    pub = {k: sol[k] for k in vk.ltgm_io}
    pub2 = wire.solution_of_yojson(wire.yojson_of_solution(pub))
    # a proof made with the reloaded key verifies under the reloaded vkey
    proof2 = Pr.prove(rng, q, pk2, sol)
    assert Pr.verify(pub2, vk2, proof2) is True
    # the oracle's own compression of the same points gives the same document
    op = H.decode_groth16_proof(proof)
    assert jp == wire.dumps({"a": O.g1_compress(op[0]), "b": O.g2_compress(op[1]), "c": O.g1_compress(op[2])})
    # a tampered point string is rejected when it is read
    bad = wire.loads(jp)
    bad[b"a"] = bytes([bad[b"a"][0] & 0x7f]) + bad[b"a"][1:]
    with pytest.raises(_lib.ZkError):
        wire.Groth16Wire.proof_of_yojson(wire.dumps({k.decode(): v for k, v in bad.items()}))
    Pr.free(pk); Pr.free(pk2)
    # Pinocchio
    circ, wit = Z.circuit_pair_case(3)
    oq2 = Z.qap_build(circ.gates)
    q2 = H.mirror_qap(oq2)
    M = PN.Make()
    ppk, pvk = M.ZK.keygen(rng, H.mirror_circuit(circ), q2)
    sol = wit(4)
    pproof = M.ZK.prove(rng, q2, ppk, sol)
    W = wire.PinocchioWire
    assert W.proof_of_yojson(W.yojson_of_proof(pproof)) == pproof
    assert list(wire.loads(W.yojson_of_proof(pproof))) == [f.encode() for f in PN.PROOF_FIELDS]
    pvk2 = W.vkey_of_yojson(W.yojson_of_vkey(pvk))
    assert pvk2 == pvk
    ppk2 = W.pkey_of_yojson(W.yojson_of_pkey(ppk))
    for f, _t in W.PKEY[1]:
        assert getattr(ppk2, f) == getattr(ppk, f), f
    ios = {k: sol[k] for k in pvk["vv_io"]}
    assert M.ZK.verify(ios, pvk2, M.ZK.prove(rng, q2, ppk2, sol)) is True
    M.ZK.free(ppk); M.ZK.free(ppk2)
    q.free(); q2.free()
