"""``Protocol.S`` (/root/reference/src/lib/zk/protocol.mli:3-28) as a Python protocol class,
plus the slice of ``Circuit.t`` (src/lib/zk/circuit.ml:108-134) that key generation reads."""

from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Sequence

from .curve import Var


@dataclass
class Circuit:
    """circuit.ml:108-113 without the gates (the QAP already encodes them)."""
    inputs_public: Sequence[Var]
    outputs: Sequence[Var]
    mids: Sequence[Var]
    vars: Sequence[Var]          # Circuit.vars circuit.gates (circuit.ml:125-130)

    def ios(self) -> List[Var]:
        """circuit.ml:132-134."""
        mids = set(self.mids)
        return [v for v in sorted(self.vars) if v not in mids]


class ProtocolS:
    """protocol.mli:3-28: keygen / prove / verify.  ``rng`` is a ``random.Random`` standing for
    ``Gen.rng = Random.State.t`` (misclib/gen.ml:1)."""

    def keygen(self, rng: random.Random, circuit: Circuit, qap):      # pragma: no cover - interface
        raise NotImplementedError

    def prove(self, rng: random.Random, qap, pkey, sol: Dict[Var, int]):  # pragma: no cover
        raise NotImplementedError

    def verify(self, input_output: Dict[Var, int], vkey, proof) -> bool:
        raise NotImplementedError                                      # pragma: no cover


class Test:
    """``Test.Make(F)(Protocol)`` (src/lib/test/test.mli:4-25), the part behind the DSL front-end:
    what ``test`` / ``random_test`` do once ``Comp.compile``, ``QAP.build`` and the witness
    evaluation have produced ``circuit``, ``qap`` and ``sol`` (test.ml:60-97, 119-178) —
    ``keygen``, ``prove``, ``public = sol`` minus the circuit's mids, ``assert (verify …)``."""

    def __init__(self, protocol: ProtocolS):
        self.protocol = protocol

    def run(self, rng: random.Random, circuit: Circuit, qap, sol: Dict[Var, int]):
        pkey, vkey = self.protocol.keygen(rng, circuit, qap)             # test.ml:121-122
        try:
            proof = self.protocol.prove(rng, qap, pkey, sol)             # :170
            mids = set(circuit.mids)
            public = {k: v for k, v in sol.items() if k not in mids}     # :174-176
            assert self.protocol.verify(public, vkey, proof), "Protocol.verify rejected the proof"   # :178
        finally:
            free = getattr(self.protocol, "free", None)
            if free is not None:
                free(pkey)
        return pkey, vkey, proof
