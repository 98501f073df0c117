#!/usr/bin/env python3
"""Pinocchio prove timing on BASELINE configs[1]: ZK and NonZK on the synthetic 2^10-gate pair/case
circuit (dense QAP path), each proof checked against the closed-form trapdoor identity.
Development tool: builds the circuit and the expected values with the oracle."""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import bls12_381 as O
from oracle import zk as Z
from tests import helpers as H
from zukelang_b200 import pinocchio as PN

R = O.R


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    circ, wit = Z.circuit_pair_case(n)
    oq = Z.qap_build(circ.gates)
    rr = random.Random(0x434F4E32)
    td = Z.PinocchioTrapdoor(*[rr.randrange(R) for _ in range(8)])
    M = PN.Make()
    q = H.mirror_qap(oq)
    t0 = time.time()
    pk, _ = M.ZK.keygen(random.Random(0x434F4E32), H.mirror_circuit(circ), q)
    sol = wit(2024)
    d = (rr.randrange(R), rr.randrange(R), rr.randrange(R))
    M.ZK.prove_with(d, q, pk, sol)                       # uploads key + QAP
    setup = time.time() - t0
    out = {}
    for name, fn in (("zk", lambda: M.ZK.prove_with(d, q, pk, sol)), ("nonzk", lambda: M.NonZK.prove(None, q, pk, sol))):
        ts = []
        for _ in range(5):
            t1 = time.perf_counter()
            proof = fn()
            ts.append(time.perf_counter() - t1)
        exp = Z.pinocchio_closed_form(td, oq, circ, sol, d if name == "zk" else None)
        out[name] = {"prove_ms_best": min(ts) * 1e3, "proofs_per_s": 1 / min(ts),
                     "exact_ok": H.decode_pinocchio_proof(proof) == exp}
    print(json.dumps({"probe": "pinocchio", "gates": n, "variables": len(circ.vars()), "mids": len(circ.mids),
                      "setup_s": setup, **out}))


if __name__ == "__main__":
    main()
