"""BASELINE.json's named configurations at their NAMED sizes ("next round" 1b/1c):
config 3 = Groth16 prove, 2^16-gate multiply chain; config 5 = Groth16 prove at 2^20 gates
including the G2 B-query, sharded over 8; plus the seeded random R1CS SURVEY §8d offers for config 5.

No dense QAP.t can exist at these sizes, so the proofs go through the evaluation-form
prover (zk_groth16_prove_r1cs) and are compared, byte for byte, with the ORACLE's gate-list trapdoor
identity (oracle/zk.py:groth16_closed_form_scalars — not the product's own closed form), then
replayed through the device verifier (groth16.ml:163-173) with a right and a wrong public input.
The shard test emulates base-range shards (i, N) in one process: the N partial results, added,
must be the one-shard proof — G2 included (src/groth16/groth16.ml:123-161).
"""
import ctypes
import random

import pytest

from oracle import bls12_381 as O
from oracle import zk as Z

pytestmark = pytest.mark.gpu
R = O.R


def _sparse(circ):
    from zukelang_b200.sparse import SparseCircuit
    gates = [(dict(g.lhs), dict(g.l), dict(g.r)) for g in circ.gates]
    return SparseCircuit(gates, circ.inputs_public, circ.outputs, circ.mids)


def _oracle_proof_bytes(td, r, s, circ, sol):
    """The proof the reference computes (groth16.ml:123-161), from the oracle's trapdoor identity:
    a | b | c compressed (curve.ml:199,208), 48 + 96 + 48 bytes."""
    A, B, C = Z.groth16_closed_form_scalars(Z.Groth16Trapdoor(*td), r, s, circ, sol)
    return O.g1_compress(O.G1.of_Fr(A)) + O.g2_compress(O.G2.of_Fr(B)) + O.g1_compress(O.G1.of_Fr(C))


def _prove_shard(zk, S, pk, dom, sol_b, r, s, shard):
    from zukelang_b200 import _lib
    from zukelang_b200.curve import Fr
    P = S.Groth16Sparse(shard=shard)
    out = (ctypes.c_uint8 * _lib.GROTH16_PROOF_OUT)()
    _lib.check(zk.zk_groth16_prove_r1cs(P._key_handle(pk, dom.circuit), dom.handle, sol_b, Fr.to_bytes(r), Fr.to_bytes(s), out))
    return bytes(out)


def _run_config(zk, circ, wit, seed, shards=0, verify=True):
    from zukelang_b200 import dist as D, sparse as S
    from zukelang_b200.curve import fr_vector
    sc = _sparse(circ)
    dom = S.EvalDomain(sc)
    P = S.Groth16Sparse()
    rng = random.Random(seed)
    pk, vk, td = P.keygen_with_trapdoor(rng, sc, dom.w)
    sol = wit(rng.randrange(R))
    r, s = rng.randrange(R), rng.randrange(R)
    proof = P.prove_with(r, s, dom, pk, sol)
    assert proof.to_compressed_bytes() == _oracle_proof_bytes(td, r, s, circ, sol)
    # the stage split of that proof adds up to its device time (zk_groth16_last_stage_ms)
    from zukelang_b200 import _lib
    ms, stages = ctypes.c_float(), (ctypes.c_float * 8)()
    hk = P._key_handle(pk, dom.circuit)
    _lib.check(zk.zk_groth16_last_device_ms(hk, ctypes.byref(ms)))
    _lib.check(zk.zk_groth16_last_stage_ms(hk, stages))
    assert all(x >= 0 for x in stages) and abs(sum(stages) - ms.value) <= 0.02 * ms.value + 0.05
    if verify:
        pub = {k: sol[k] for k in vk.ltgm_io}
        assert P.verify(pub, vk, proof)
        k0 = sorted(pub)[-1]
        assert not P.verify({**pub, k0: (pub[k0] + 1) % R}, vk, proof)
    if shards:
        sol_b = fr_vector(sol[k] for k in sc.variables)
        parts = [_prove_shard(zk, S, pk, dom, sol_b, r, s, (i, shards)) for i in range(shards)]
        comb = D.combine_groth16(parts)
        assert comb[96:144] + comb[336:432] + comb[528:576] == proof.to_compressed_bytes()
        assert comb[0:96] == proof.a.raw and comb[144:336] == proof.b.raw and comb[432:528] == proof.c.raw
    # a violated gate must trip the reference's `assert (is_zero rem)` (QAP.ml:134)
    bad = dict(sol)
    bad[circ.mids[len(circ.mids) // 2]] = (bad[circ.mids[len(circ.mids) // 2]] + 1) % R
    with pytest.raises(AssertionError):
        P.prove_with(r, s, dom, pk, bad)
    P.free(pk)
    dom.free()


def test_config3_groth16_2e16_multiply_chain_and_8_shards(zk):
    """BASELINE configs[2]: Groth16 prove, synthetic 2^16-constraint multiply chain; plus configs[4]'s
    sharding rule at this size: shards (i, 8) of every key query, partials added = the proof."""
    circ, wit = Z.circuit_mulchain(1 << 16)
    _run_config(zk, circ, wit, seed=0x47524F54 + 16, shards=8)


def test_config3_size_random_r1cs_2e16(zk):
    """Same size on the seeded random R1CS: every gate has its own W(j), so the G2
    B-query's scalars are spread over all buckets (the multiply chain's are all equal)."""
    circ, wit = Z.circuit_random_r1cs(1 << 16)
    _run_config(zk, circ, wit, seed=0x47524F54 + 116, shards=3)


@pytest.mark.slow
def test_config5_groth16_2e20_multiply_chain(zk):
    """BASELINE configs[4] on one GPU: 2^20 constraints, G1 queries of 2^20 / 3 * 2^20 points and the
    2^20-point G2 B-query, bytes equal to the oracle's closed form; 2 shards emulated."""
    circ, wit = Z.circuit_mulchain(1 << 20)
    _run_config(zk, circ, wit, seed=0x47524F54 + 20, shards=2, verify=True)


@pytest.mark.slow
def test_config5_size_random_r1cs_2e18(zk):
    circ, wit = Z.circuit_random_r1cs(1 << 18)
    _run_config(zk, circ, wit, seed=0x47524F54 + 118, shards=0, verify=False)
