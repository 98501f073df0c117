"""Host mirror of ``QAP.Make(F)`` (/root/reference/src/lib/zk/QAP.ml) — the evaluation half.

``QAP.t`` keeps the reference's shape (QAP.ml:11-16): three ``Var.Map``s of polynomials and
the target.  ``eval`` (QAP.ml:120-135) runs on the GPU: the dense maps are flattened once
into m x n matrices resident in HBM (``zk_qap_load``), then every call is one
``zk_qap_eval``.  ``build`` (QAP.ml:18-94) is upstream of the hot path; it is here in the fast
form SURVEY.md §8f-2 asks for — one shared Lagrange basis on the points 0..n-1 instead of one
``Polynomial.interpolate`` per variable — so that the dense path is usable at 2^10 gates.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

from . import _lib
from .curve import R, Var, fr_vector

Poly = List[int]


def degree(p: Poly) -> int:
    """polynomial.ml:253-255."""
    n = len(p)
    while n and p[n - 1] % R == 0:
        n -= 1
    return 0 if n == 0 else n - 1


@dataclass
class QAP:
    v: Dict[Var, Poly]
    w: Dict[Var, Poly]
    y: Dict[Var, Poly]
    target: Poly
    _handle: int = field(default=0, repr=False, compare=False)

    # ---- device residency ------------------------------------------------------------
    def variables(self) -> List[Var]:
        return sorted(self.v)

    @property
    def n(self) -> int:
        return degree(self.target)

    def handle(self) -> int:
        if self._handle:
            return self._handle
        if not (set(self.v) == set(self.w) == set(self.y)):
            raise AssertionError("QAP domains differ")          # QAP.ml:104-105
        n, keys = self.n, self.variables()
        def flat(mp):
            rows = []
            for k in keys:
                p = mp[k]
                if len(p) > n and any(c % R for c in p[n:]):
                    raise _lib.InvalidArgument(_lib.ZK_EARG, "QAP polynomial of degree >= n")
                rows.append(fr_vector(list(p[:n]) + [0] * (n - min(len(p), n))))
            return b"".join(rows)
        h = ctypes.c_uint64()
        _lib.check(_lib.lib().zk_qap_load(flat(self.v), flat(self.w), flat(self.y),
                                          fr_vector(list(self.target[:n + 1])), len(keys), n, ctypes.byref(h)))
        self._handle = h.value
        return self._handle

    def free(self) -> None:
        if self._handle:
            _lib.check(_lib.lib().zk_qap_free(self._handle))
            self._handle = 0


def _normalize(p: Poly) -> Poly:
    n = len(p)
    while n and p[n - 1] == 0:
        n -= 1
    return p[:n]


def _unpack(buf: bytes) -> Poly:
    return [int.from_bytes(buf[i:i + 32], "little") for i in range(0, len(buf), 32)]


def eval(sol: Dict[Var, int], qap: QAP) -> Tuple[None, Poly]:
    """QAP.ml:120-135.  Returns ``(p, h)`` like the reference; ``p`` (unused by every caller —
    groth16.ml:236 and pinocchio.ml:537,560 bind it to ``_p``) is not materialised.
    A witness that does not satisfy the circuit trips the reference's ``assert`` (QAP.ml:134):
    here that is an ``AssertionError``."""
    _, h, _ = eval_full(sol, qap)
    return None, h


def eval_full(sol: Dict[Var, int], qap: QAP):
    """(None, h, (V, W, Y)) — additionally exposes the combinations of QAP.ml:129-131."""
    keys = qap.variables()
    missing = [k for k in keys if k not in sol]
    if missing:
        raise AssertionError("Variable %s__%d not found" % missing[0])   # var.ml:72-78
    n = qap.n
    h_out = (ctypes.c_uint8 * (32 * max(n - 1, 1)))()
    vwy = (ctypes.c_uint8 * (96 * n))()
    rc = _lib.lib().zk_qap_eval(qap.handle(), fr_vector(sol[k] for k in keys), h_out, vwy)
    if rc == _lib.ZK_EREMAINDER:
        raise AssertionError("QAP.eval: remainder is not zero")
    _lib.check(rc)
    h = _normalize(_unpack(bytes(h_out)[:32 * (n - 1)]))
    c = _unpack(bytes(vwy))
    return None, h, (_normalize(c[:n]), _normalize(c[n:2 * n]), _normalize(c[2 * n:]))


# ---- QAP.ml:18-94 ---------------------------------------------------------------------------
def _vanishing(n: int) -> Poly:
    """prod_{j<n} (x - j), lowest degree first (``Polynomial.z``, polynomial.ml:248-251)."""
    z = [1]
    for j in range(n):
        nxt = [0] * (len(z) + 1)
        for i, c in enumerate(z):
            nxt[i] = (nxt[i] - j * c) % R
            nxt[i + 1] = (nxt[i + 1] + c) % R
        z = nxt
    return z


def _basis(n: int, z: Poly) -> List[Poly]:
    """The n Lagrange basis polynomials of the points 0..n-1:  L_j = w_j * Z / (x - j) with the
    barycentric weight  w_j = 1 / prod_{i != j} (j - i) = (-1)^(n-1-j) / (j! (n-1-j)!)."""
    fact = [1] * max(n, 1)
    for i in range(1, n):
        fact[i] = fact[i - 1] * i % R
    out = []
    for j in range(n):
        w = pow(fact[j] * fact[n - 1 - j] % R, -1, R)
        if (n - 1 - j) & 1:
            w = R - w
        q = [0] * n                      # synthetic division of Z by (x - j); the remainder is 0
        carry = 0
        for i in range(n, 0, -1):
            carry = (z[i] + carry * j) % R
            q[i - 1] = carry
        out.append([c * w % R for c in q])
    return out


def build(gates) -> Tuple[QAP, List[Tuple[int, object]]]:
    """``QAP.build`` (QAP.ml:18-94).  Gates are numbered r_g = 0..n-1 in ``Gate.Set`` order (:22);
    for every variable k, v_k / w_k / y_k interpolate its coefficient in the left factor / right
    factor / left-hand side of gate r_g at the point ``F.of_int r_g`` (:26-49, :81-90); the target is
    ``prod (x - r_g)`` (:92).  Returns ``(qap, rgs)`` like the reference.

    Same polynomials as the reference's per-variable ``Polynomial.interpolate`` (they are unique),
    obtained from one shared basis: O(n^2 + nnz * n) field operations instead of O(m * n^3)."""
    from .protocol import gate_set
    gs = gate_set(gates)
    n = len(gs)
    variables = sorted(set().union(*[g.vars() for g in gs])) if gs else []
    z = _vanishing(n)
    basis = _basis(n, z)

    def interpolate(select):
        acc = {k: [0] * n for k in variables}
        for rg, g in enumerate(gs):
            bj = basis[rg]
            for k, c in select(g):
                if c:
                    row = acc[k]
                    for i in range(n):
                        row[i] = (row[i] + c * bj[i]) % R
        return {k: _normalize(p) for k, p in acc.items()}

    qap = QAP(v=interpolate(lambda g: g.l), w=interpolate(lambda g: g.r), y=interpolate(lambda g: g.lhs), target=z)
    return qap, list(enumerate(gs))
