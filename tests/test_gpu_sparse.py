"""Evaluation-form Groth16 (zukelang_b200/sparse.py, zk_groth16_prove_r1cs): same proof bytes as the
dense-QAP path and as the oracle's literal restatement of groth16.ml:123-161, plus the exact
closed-form check at sizes where no dense QAP can exist."""
import random

import pytest

from oracle import bls12_381 as O
from oracle import zk as Z
from tests import helpers as H

pytestmark = pytest.mark.gpu
R = O.R


def _sparse(circ):
    from zukelang_b200.sparse import SparseCircuit
    gates = [(dict(g.lhs), dict(g.l), dict(g.r)) for g in circ.gates]
    return SparseCircuit(gates, circ.inputs_public, circ.outputs, circ.mids)


@pytest.mark.parametrize("name,n", [("mulchain", 8), ("mulchain", 13), ("pair_case", 32), ("cubic", 3)])
def test_sparse_path_equals_dense_path_and_oracle(zk, name, n):
    from zukelang_b200 import groth16 as G16, sparse as S
    circ, wit = {"mulchain": lambda: Z.circuit_mulchain(n), "pair_case": lambda: Z.circuit_pair_case(n),
                 "cubic": Z.circuit_cubic}[name]()
    oq = Z.qap_build(circ.gates)
    sc = _sparse(circ)
    dom = S.EvalDomain(sc)
    P = S.Groth16Sparse()
    pk, vk, td = P.keygen_with_trapdoor(random.Random(n), sc, dom.w)
    otd = Z.Groth16Trapdoor(*td)
    opk, ovk = Z.groth16_keygen(otd, circ, oq, with_ab=(n <= 8))
    # the reference-shaped fields of the derived key equal the oracle's
    assert pk.a == H.mirror_point(opk.a) and pk.b2 == H.mirror_point(opk.b2, True)
    assert pk.ltd_mid == {k: H.mirror_point(v) for k, v in opk.ltd_mid.items()}
    rng = random.Random(99 + n)
    for trial in range(2):
        sol = wit(rng.randrange(R)) if name != "pair_case" else wit(trial)
        r, s = rng.randrange(R), rng.randrange(R)
        proof = P.prove_with(r, s, dom, pk, sol)
        # dense path through the monomial key, same trapdoor
        dense = G16.Make()
        q = H.mirror_qap(oq)
        dproof = dense.prove_with(r, s, q, H.mirror_groth16_pkey(opk), sol)
        assert proof.to_compressed_bytes() == dproof.to_compressed_bytes()
        assert H.decode_groth16_proof(proof) == Z.groth16_closed_form(otd, r, s, oq, circ, sol)
        if n <= 8:
            assert proof.to_compressed_bytes() == H.groth16_proof_compressed(Z.groth16_prove(r, s, oq, opk, sol))
            assert Z.groth16_verify({k: sol[k] for k in ovk.ltgm_io}, ovk, H.decode_groth16_proof(proof))
        q.free()
    bad = dict(sol)
    bad[circ.mids[0]] = (bad[circ.mids[0]] + 1) % R
    with pytest.raises(AssertionError):
        P.prove_with(1, 2, dom, pk, bad)
    P.free(pk)
    dom.free()


def test_sparse_groth16_4096_closed_form(zk):
    """A size no dense QAP.t could reach in the reference (m n = 2^24 coefficients per map)."""
    from zukelang_b200 import sparse as S
    n = 4096
    circ, wit = Z.circuit_mulchain(n)
    sc = _sparse(circ)
    dom = S.EvalDomain(sc)
    P = S.Groth16Sparse()
    pk, vk, td = P.keygen_with_trapdoor(random.Random(7), sc, dom.w)
    sol = wit(0xC0FFEE)
    r, s = 123456789, 987654321
    proof = P.prove_with(r, s, dom, pk, sol)
    A, B, C = S.closed_form_scalars(td, r, s, sc, dom.w, sol)
    G1, G2 = P.C.G1, P.C.G2
    assert proof.a == G1.of_Fr(A) and proof.b == G2.of_Fr(B) and proof.c == G1.of_Fr(C)
    P.free(pk)
    dom.free()
