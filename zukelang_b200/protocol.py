"""``Protocol.S`` (/root/reference/src/lib/zk/protocol.mli:3-28) as a Python protocol class,
plus the slice of ``Circuit.t`` (src/lib/zk/circuit.ml:108-134) that key generation reads."""

from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

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
