"""``Protocol.S`` (/root/reference/src/lib/zk/protocol.mli:3-28) as a Python protocol class,
plus the slice of ``Circuit.t`` (src/lib/zk/circuit.ml:108-134) that key generation reads."""

from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

from .curve import R as _R, Var


Affine = Dict[Var, int]      # circuit.ml:8-71: a linear combination  sum_k c_k * var_k


@dataclass(frozen=True)
class Gate:
    """circuit.ml:73-76: ``{ lhs; l; r }`` stands for the constraint ``lhs = l * r`` between three
    affine forms.  Stored as sorted binding tuples (the shape ``Var.Map.bindings`` gives)."""
    lhs: Tuple[Tuple[Var, int], ...]
    l: Tuple[Tuple[Var, int], ...]
    r: Tuple[Tuple[Var, int], ...]

    @staticmethod
    def make(lhs: Affine, l: Affine, r: Affine) -> "Gate":
        norm = lambda a: tuple(sorted((k, c % _R) for k, c in a.items()))
        return Gate(norm(lhs), norm(l), norm(r))

    def compare_key(self):
        """circuit.ml:85-91: lexicographic on (lhs, l, r), each compared as ``Var.Map.compare``
        does (binding lists in key order).  Coefficients are compared as integers; how the
        reference's ``F.compare`` orders Fr values cannot be checked here."""
        return (list(self.lhs), list(self.l), list(self.r))

    def vars(self):
        """circuit.ml:93-94."""
        return {k for part in (self.lhs, self.l, self.r) for k, _ in part}


def gate_set(gates: Sequence[Gate]) -> List[Gate]:
    """``Gate.Set`` (circuit.ml:96-105) as a list: duplicates removed, ``Gate.compare`` order —
    the order that numbers the gates r_g = 0..n-1 in ``QAP.build`` (QAP.ml:22)."""
    return sorted(set(gates), key=Gate.compare_key)


@dataclass
class Circuit:
    """circuit.ml:108-113.  ``vars`` is ``Circuit.vars circuit.gates`` (circuit.ml:125-130); the
    gates themselves are optional here because key generation and proving only read the QAP."""
    inputs_public: Sequence[Var]
    outputs: Sequence[Var]
    mids: Sequence[Var]
    vars: Sequence[Var]
    gates: Sequence[Gate] = ()

    @staticmethod
    def of_gates(gates: Sequence[Gate], inputs_public: Sequence[Var], outputs: Sequence[Var],
                 mids: Sequence[Var]) -> "Circuit":
        gs = gate_set(gates)
        vs = sorted(set().union(*[g.vars() for g in gs])) if gs else []
        return Circuit(inputs_public=list(inputs_public), outputs=list(outputs), mids=list(mids), vars=vs, gates=gs)

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
