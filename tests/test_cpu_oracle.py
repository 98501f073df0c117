"""CPU suite: the oracle against public constants, the reference's own KATs and the committed
golden vectors; the C ABI's symbol table; host-side mirror logic.  No GPU needed."""
import ctypes
import json
import os
import re

import pytest

from oracle import bls12_381 as O
from oracle import pairing as OP
from oracle import zk as Z

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
R = O.R


def test_oracle_public_constants():
    O.self_check()


def test_oracle_pairing_bilinear():
    OP.self_check()


def test_reference_polynomial_kats():
    Z.polynomial_self_test()


def test_golden_msm_vectors():
    for vec in GOLD["msm"]:
        G = O.G1 if vec["group"] == "G1" else O.G2
        comp = O.g1_compress if vec["group"] == "G1" else O.g2_compress
        dl = [int(x) for x in vec["dlogs"]]
        ks = [int(x) for x in vec["scalars"]]
        total = sum(a * b for a, b in zip(dl, ks)) % R
        assert comp(G.mul(G.one, total)).hex() == vec["result_compressed"]
    # the literal fold on the small ones
    for vec in GOLD["msm"]:
        if vec["n"] > 3:
            continue
        G = O.G1 if vec["group"] == "G1" else O.G2
        comp = O.g1_compress if vec["group"] == "G1" else O.g2_compress
        pts = [G.mul(G.one, int(d)) for d in vec["dlogs"]]
        acc = None
        for p, k in zip(pts, vec["scalars"]):
            acc = G.add(G.mul(p, int(k)), acc)
        assert comp(acc).hex() == vec["result_compressed"]


def test_golden_quotients():
    for vec in GOLD["quotient"]:
        ints = lambda xs: [int(x) for x in xs]
        p = Z.poly_sub(Z.poly_mul(ints(vec["V"]), ints(vec["W"])), ints(vec["Y"]))
        h, rem = Z.poly_div_rem(p, ints(vec["T"]))
        assert rem == []
        n = vec["n"]
        assert (h + [0] * n)[:n - 1] == ints(vec["h"])


def test_golden_groth16_config1():
    g = GOLD["groth16_config1"]
    circ, wit = Z.circuit_cubic()
    td = Z.Groth16Trapdoor(*[int(v) for v in g["trapdoor"]])
    sol = wit(int(g["x"]))
    qap = Z.qap_build(circ.gates, literal=True)
    assert qap == Z.qap_build(circ.gates)
    proof = Z.groth16_closed_form(td, int(g["r"]), int(g["s"]), qap, circ, sol)
    got = O.g1_compress(proof[0]) + O.g2_compress(proof[1]) + O.g1_compress(proof[2])
    assert got.hex() == g["proof_compressed"]
    # round trip of the wire format
    assert O.g1_decompress(got[:48]) == proof[0] and O.g2_decompress(got[48:144]) == proof[1]


def test_golden_pinocchio_small():
    g = GOLD["pinocchio_small"]
    circ, wit = Z.circuit_pair_case(7)
    td = Z.PinocchioTrapdoor(*[int(v) for v in g["trapdoor"]])
    sol = wit(g["witness_seed"])
    qap = Z.qap_build(circ.gates)
    for name, zk in (("nonzk", None), ("zk", tuple(int(v) for v in g["d"]))):
        pr = Z.pinocchio_closed_form(td, qap, circ, sol, zk)
        got = b"".join((O.g2_compress if grp == "G2" else O.g1_compress)(pr[f])
                       for f, grp in zip(Z.PINOCCHIO_PROOF_FIELDS, Z.PINOCCHIO_PROOF_GROUPS))
        assert got.hex() == g[name]


def test_serialisation_round_trip_and_rejects():
    import random
    rng = random.Random(3)
    for _ in range(4):
        p = O.G1.mul(O.G1.one, rng.randrange(R))
        q = O.G2.mul(O.G2.one, rng.randrange(R))
        assert O.g1_decompress(O.g1_compress(p)) == p and O.g1_from_uncompressed(O.g1_to_uncompressed(p)) == p
        assert O.g2_decompress(O.g2_compress(q)) == q and O.g2_from_uncompressed(O.g2_to_uncompressed(q)) == q
    assert O.g1_decompress(O.g1_compress(None)) is None and O.g2_from_uncompressed(O.g2_to_uncompressed(None)) is None
    with pytest.raises(ValueError):
        O.g1_from_uncompressed(bytes(95) + b"\x01")


def test_qap_fast_build_equals_literal_build():
    for circ, _ in (Z.circuit_cubic(), Z.circuit_mulchain(6), Z.circuit_pair_case(8)):
        assert Z.qap_build(circ.gates) == Z.qap_build(circ.gates, literal=True)


def test_abi_exports_every_declared_symbol():
    """include/zkb200.h <-> libzkb200.so <-> the ctypes table: same set of entry points."""
    from zukelang_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(zk_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no prototypes parsed"
    lib = _lib.load()
    missing = sorted(s for s in declared if not hasattr(lib, s))
    assert not missing, "declared in zkb200.h but not exported: %s" % missing
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))


def test_library_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from zukelang_b200 import _lib
    lib = _lib.load()
    assert lib.zk_init(0) == _lib.ZK_ECUDA
    assert b"no CPU fallback" in lib.zk_last_error()
    out = (ctypes.c_uint8 * 144)()
    assert lib.zk_g1_msm(None, None, None, 0, out) == _lib.ZK_ECUDA     # even the trivial call refuses


def test_product_tree_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zukelang_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
