#!/usr/bin/env python3
"""Generates tests/golden/*.json from the CPU oracle (oracle/).

The reference ships no byte-level vectors and cannot run here (no OCaml toolchain), so
these vectors are produced by the oracle restatement and are pinned only as far as the
oracle is (see oracle/bls12_381.py: PARITY UNPINNED).  Every proof in the file has been
accepted by the oracle's replay of the reference verifier at generation time.

Run:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import bls12_381 as O          # noqa: E402
from oracle import zk as Z                 # noqa: E402

R = O.R


def msm_vectors():
    out = []
    rng = random.Random(0x474F4C44)
    for gname, G, comp, sizes in (("G1", O.G1, O.g1_compress, (1, 2, 3, 17, 64)), ("G2", O.G2, O.g2_compress, (1, 2, 3, 9))):
        for n in sizes:
            dl = [rng.randrange(1, R) for _ in range(n)]
            ks = [rng.randrange(R) for _ in range(n)]
            if n >= 3:
                dl[1] = 0                          # identity base
                ks[2] = 0                          # zero scalar
                ks[0] = R - 1
            if n >= 17:
                ks[5], ks[6], ks[7] = 1, 1 << 255 & (R - 1), 1 << 128
                dl[9] = dl[8]                      # repeated base
                dl[11] = R - dl[10]                # P and -P
                ks[11] = ks[10]
            pts = [G.mul(G.one, d) for d in dl]
            acc = None
            for p, k in zip(pts, ks):              # curve.ml:91-103 fold
                acc = G.add(G.mul(p, k), acc)
            assert acc == G.mul(G.one, sum(a * b for a, b in zip(dl, ks)) % R)
            out.append(dict(group=gname, n=n, dlogs=[str(d) for d in dl], scalars=[str(k) for k in ks],
                            result_compressed=comp(acc).hex()))
    return out


def quotient_vectors():
    out = []
    for name, (circ, wit), x in (("cubic", Z.circuit_cubic(), 3), ("mulchain8", Z.circuit_mulchain(8), 5),
                                 ("pair_case7", Z.circuit_pair_case(7), 11)):
        sol = wit(x)
        qap = Z.qap_build(circ.gates, literal=True)
        _p, h = Z.qap_eval(sol, qap)
        n = Z.poly_degree(qap.target)
        pad = lambda p: [str(c) for c in (list(p) + [0] * n)[:n]]
        out.append(dict(name=name, n=n, V=pad(Z.qap_combine(sol, qap.v)), W=pad(Z.qap_combine(sol, qap.w)),
                        Y=pad(Z.qap_combine(sol, qap.y)), T=[str(c) for c in qap.target],
                        h=[str(c) for c in (list(h) + [0] * n)[:n - 1]]))
    return out


def groth16_vector():
    circ, wit = Z.circuit_cubic()
    rng = random.Random(0x47313631)
    td = Z.Groth16Trapdoor(*[rng.randrange(R) for _ in range(5)])
    x, r, s = rng.randrange(R), rng.randrange(R), rng.randrange(R)
    sol = wit(x)
    qap = Z.qap_build(circ.gates, literal=True)
    pk, vk = Z.groth16_keygen(td, circ, qap)
    proof = Z.groth16_prove(r, s, qap, pk, sol)
    assert proof == Z.groth16_closed_form(td, r, s, qap, circ, sol)
    assert Z.groth16_verify({k: sol[k] for k in vk.ltgm_io}, vk, proof)
    return dict(config="README x*x*x + x + 3 (test.ml:194-197)", trapdoor=[str(v) for v in (td.a, td.b, td.gm, td.d, td.t)],
                x=str(x), r=str(r), s=str(s),
                proof_compressed=(O.g1_compress(proof[0]) + O.g2_compress(proof[1]) + O.g1_compress(proof[2])).hex())


def pinocchio_vector():
    circ, wit = Z.circuit_pair_case(7)
    rng = random.Random(0x50494E4F)
    td = Z.PinocchioTrapdoor(*[rng.randrange(R) for _ in range(8)])
    d = tuple(rng.randrange(R) for _ in range(3))
    sol = wit(42)
    qap = Z.qap_build(circ.gates, literal=True)
    pk, vk = Z.pinocchio_keygen(td, circ, qap)
    res = {}
    for name, zk in (("nonzk", None), ("zk", d)):
        pr = Z.pinocchio_prove(qap, pk, sol, zk)
        assert pr == Z.pinocchio_closed_form(td, qap, circ, sol, zk)
        assert Z.pinocchio_verify({k: sol[k] for k in vk["vv_io"]}, vk, pr)
        res[name] = b"".join((O.g2_compress if g == "G2" else O.g1_compress)(pr[f])
                             for f, g in zip(Z.PINOCCHIO_PROOF_FIELDS, Z.PINOCCHIO_PROOF_GROUPS)).hex()
    return dict(config="pair/case shaped circuit, 7 gates", trapdoor=[str(getattr(td, f)) for f in
                ("rv", "rw", "s", "av", "aw", "ay", "b", "gm")], witness_seed=42, d=[str(v) for v in d], **res)


def main():
    data = dict(msm=msm_vectors(), quotient=quotient_vectors(), groth16_config1=groth16_vector(),
                pinocchio_small=pinocchio_vector())
    with open(os.path.join(HERE, "vectors.json"), "w") as f:
        json.dump(data, f, indent=1)
    print("wrote vectors.json:", {k: (len(v) if isinstance(v, list) else 1) for k, v in data.items()})


if __name__ == "__main__":
    main()
