"""Large-circuit front door: Groth16 prove from the sparse gate list.

The reference's ``prove rng qap pkey sol`` needs the dense ``QAP.t`` — m polynomials of n
coefficients — which cannot exist at 2^16+ constraints (137 GB at n = m = 2^16).  This module is
the sibling entry for such circuits.  It keeps the reference's QAP *definition* (domain 0..n-1,
``QAP.ml:84``; target prod (x - j), ``QAP.ml:92``) and its proof equations (``groth16.ml:123-161``)
and changes only the basis the work is done in:

* key generation (which knows tau) emits the Lagrange-basis points ``[L_j(tau)]1``, ``[L_j(tau)]2``
  and ``[L'_k(tau) Z(tau)/delta]1`` (L' on the shifted points n..2n-1) next to the reference's
  fields a, b1, b2, d1, d2, ltd_mid;
* the prover evaluates V, W, Y on the domain with a sparse mat-vec, extrapolates them to n..2n-1
  with one NTT convolution, and runs the same three MSMs.

A, B, C are the same group elements as the reference's — ``tests/test_gpu_sparse.py`` compares the
bytes of this path with the dense path on the same circuit, key trapdoor and (r, s).

Scalars that depend on the trapdoor are host-side Python integers (key generation is outside
the prover hot path); every group element comes from the CUDA fixed-base kernel.
"""

from __future__ import annotations

import ctypes
import random
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

from . import _lib
from .curve import Bls12_381, Fr, Point, R, Var, fr_vector


def batch_inverse(xs: Sequence[int]) -> List[int]:
    pre, acc = [], 1
    for x in xs:
        pre.append(acc)
        acc = acc * x % R
    inv = pow(acc, -1, R)
    out = [0] * len(xs)
    for i in range(len(xs) - 1, -1, -1):
        out[i] = inv * pre[i] % R
        inv = inv * xs[i] % R
    return out


def domain_constants(n: int) -> Tuple[List[int], List[int]]:
    """(w, t_shift): w_j = 1 / prod_{i != j} (j - i) on 0..n-1 and t(n + k) = (n+k)! / k!."""
    fact = [1] * (2 * n)
    for i in range(1, 2 * n):
        fact[i] = fact[i - 1] * i % R
    ifact = batch_inverse(fact)
    w = []
    for j in range(n):
        v = ifact[j] * ifact[n - 1 - j] % R
        w.append((R - v) % R if (n - 1 - j) & 1 else v)
    t_shift = [fact[n + k] * ifact[k] % R for k in range(n)]
    return w, t_shift


@dataclass
class SparseCircuit:
    """Gates in Gate.Set order (circuit.ml:73-106): each gate is (lhs, l, r) with sparse rows
    {Var: coeff}.  ``variables`` lists every variable in increasing Var order."""
    gates: List[Tuple[Dict[Var, int], Dict[Var, int], Dict[Var, int]]]
    inputs_public: Sequence[Var]
    outputs: Sequence[Var]
    mids: Sequence[Var]
    variables: List[Var] = field(default_factory=list)

    def __post_init__(self):
        if not self.variables:
            s = set()
            for g in self.gates:
                for row in g:
                    s.update(row)
            self.variables = sorted(s)

    @property
    def n(self) -> int:
        return len(self.gates)

    def csr(self, which: int):
        """CSR of matrix `which` (0 = l, 1 = r, 2 = lhs) as (row_ptr, col, val) lists."""
        pos = {v: i for i, v in enumerate(self.variables)}
        sel = (1, 2, 0)[which]          # gates are stored (lhs, l, r)
        row_ptr, col, val = [0], [], []
        for g in self.gates:
            for v in sorted(g[sel]):
                c = g[sel][v] % R
                if c:
                    col.append(pos[v])
                    val.append(c)
            row_ptr.append(len(col))
        return row_ptr, col, val


class EvalDomain:
    """Device-resident evaluation domain + the circuit's three sparse matrices."""

    def __init__(self, circuit: SparseCircuit):
        self.circuit = circuit
        n = circuit.n
        self.w, self.t_shift = domain_constants(n)
        h = ctypes.c_uint64()
        _lib.check(_lib.lib().zk_eval_domain_load(n, fr_vector(self.w), fr_vector(self.t_shift), ctypes.byref(h)))
        self.handle = h.value
        m = len(circuit.variables)
        for which in range(3):
            rp, col, val = circuit.csr(which)
            a_rp = (ctypes.c_uint32 * len(rp))(*rp)
            a_col = (ctypes.c_uint32 * max(len(col), 1))(*col)
            _lib.check(_lib.lib().zk_r1cs_load(self.handle, which, m, a_rp, a_col, fr_vector(val) or b"\0"))

    def free(self):
        if self.handle:
            _lib.check(_lib.lib().zk_qap_free(self.handle))
            self.handle = 0


@dataclass
class DerivedPKey:
    """The reference's pkey (groth16.ml:24-34) with the three monomial lists replaced by their
    Lagrange-basis counterparts."""
    a: Point
    d1: Point
    b1: Point
    b2: Point
    d2: Point
    lag1: List[Point]            # [L_j(tau)]1, j < n          (plays ti1)
    lag2: List[Point]            # [L_j(tau)]2                 (plays ti2)
    hk: List[Point]              # [L'_k(tau) Z(tau)/delta]1   (plays tiztd, n points)
    ltd_mid: Dict[Var, Point]
    _handles: Dict[Tuple[int, int], int] = field(default_factory=dict, repr=False, compare=False)


def lagrange_at(n: int, w: Sequence[int], tau: int, shift: int = 0) -> Tuple[List[int], int]:
    """([L_j(tau)] for the points shift..shift+n-1, prod (tau - point))."""
    diffs = [(tau - shift - j) % R for j in range(n)]
    z = 1
    for d in diffs:
        z = z * d % R
    inv = batch_inverse(diffs)
    return [z * w[j] % R * inv[j] % R for j in range(n)], z


class Groth16Sparse:
    """keygen / prove for a SparseCircuit; same RNG draw order as Groth16.Make (groth16.ml:51-55,
    124-125)."""

    def __init__(self, C=Bls12_381, shard: Tuple[int, int] = (0, 1)):
        self.C = C
        self.shard = shard

    def keygen_scalars(self, trapdoor, circuit: SparseCircuit, w: Sequence[int]):
        a, b, gm, d, t = trapdoor
        n = circuit.n
        lag, zt = lagrange_at(n, w, t)                       # L_j(tau), Z(tau)
        lagh, _ = lagrange_at(n, w, t, shift=n)              # L'_k(tau)
        dinv, gminv = pow(d, -1, R), pow(gm, -1, R)
        ztd = zt * dinv % R
        acc = {v: 0 for v in circuit.variables}              # L_k(tau) = b v_k + a w_k + y_k at tau
        for j, (lhs, l, r) in enumerate(circuit.gates):
            lj = lag[j]
            for v, c in l.items():
                acc[v] = (acc[v] + b * c % R * lj) % R
            for v, c in r.items():
                acc[v] = (acc[v] + a * c % R * lj) % R
            for v, c in lhs.items():
                acc[v] = (acc[v] + c * lj) % R
        mids = sorted(set(circuit.mids) & set(circuit.variables))
        ios = sorted((set(circuit.inputs_public) | set(circuit.outputs)) & set(circuit.variables))
        return dict(lag=lag, hk=[x * ztd % R for x in lagh], ltd={k: acc[k] * dinv % R for k in mids},
                    ltgm={k: acc[k] * gminv % R for k in ios}, zt=zt, mids=mids, ios=ios)

    def keygen(self, rng: random.Random, circuit: SparseCircuit, w: Sequence[int]):
        """``keygen rng circuit qap`` of Protocol.S (protocol.mli:21) for a sparse circuit: returns
        (pkey, vkey) only — the toxic waste is dropped, as groth16.ml:227-233 does."""
        pkey, vkey, _ = self.keygen_with_trapdoor(rng, circuit, w)
        return pkey, vkey

    def keygen_with_trapdoor(self, rng: random.Random, circuit: SparseCircuit, w: Sequence[int]):
        """TEST-ONLY variant that also returns (a, b, gm, d, t): whoever holds it can forge proofs.
        The parity tests need it for the closed-form identities (SURVEY.md §8c iv)."""
        trapdoor = tuple(Fr.gen(rng) for _ in range(5))     # a, b, gm, d, t  (groth16.ml:51-55)
        pkey, vkey = self.keygen_from_trapdoor(trapdoor, circuit, w)
        return pkey, vkey, trapdoor

    def keygen_from_trapdoor(self, trapdoor, circuit: SparseCircuit, w: Sequence[int]):
        from .groth16 import VKey
        G1, G2 = self.C.G1, self.C.G2
        a, b, gm, d, t = trapdoor
        sc = self.keygen_scalars(trapdoor, circuit, w)
        n = circuit.n
        s1 = [a, d, b] + sc["lag"] + sc["hk"] + [sc["ltd"][k] for k in sc["mids"]] + [sc["ltgm"][k] for k in sc["ios"]]
        p1 = G1.fixed_base(s1)
        p2 = G2.fixed_base([b, d, gm] + sc["lag"])
        o = 3
        lag1 = p1[o:o + n]; o += n
        hk = p1[o:o + n]; o += n
        ltd = p1[o:o + len(sc["mids"])]; o += len(sc["mids"])
        pkey = DerivedPKey(a=p1[0], d1=p1[1], b1=p1[2], b2=p2[0], d2=p2[1], lag1=lag1, lag2=p2[3:], hk=hk,
                           ltd_mid=dict(zip(sc["mids"], ltd)))
        vkey = VKey(one1=G1.one, ltgm_io=dict(zip(sc["ios"], p1[o:])), one2=G2.one, gm=p2[2], d=p2[1],
                    ab=self.C.Pairing.pairing(p1[0], p2[0]))               # groth16.ml:103
        return pkey, vkey

    def verify(self, w_io: Dict[Var, int], vkey, proof) -> bool:
        """groth16.ml:163-173 — the verification key and equation do not depend on the basis the
        prover worked in, so this is Groth16.Make(C).verify."""
        from .groth16 import Make
        return Make(self.C).verify(w_io, vkey, proof)

    def _key_handle(self, pkey: DerivedPKey, circuit: SparseCircuit) -> int:
        if self.shard in pkey._handles:
            return pkey._handles[self.shard]
        pos = {k: i for i, k in enumerate(circuit.variables)}
        mids = sorted(pkey.ltd_mid)
        n = circuit.n
        cat = lambda pts: b"".join(p.raw for p in pts)
        bufs = dict(a=pkey.a.raw, b1=pkey.b1.raw, d1=pkey.d1.raw, b2=pkey.b2.raw, d2=pkey.d2.raw,
                    ti1=cat(pkey.lag1), ti2=cat(pkey.lag2), tiztd=cat(pkey.hk), ltd_mid=cat(pkey.ltd_mid[k] for k in mids))
        cb = {k: ctypes.create_string_buffer(v, len(v)) for k, v in bufs.items()}
        idx = (ctypes.c_uint32 * max(len(mids), 1))(*[pos[k] for k in mids])
        st = _lib.Groth16PKeyStruct(n=n, m=len(circuit.variables), n_mid=len(mids), n_h=n,
                                    mid_index=ctypes.addressof(idx), **{k: ctypes.addressof(v) for k, v in cb.items()})
        h = ctypes.c_uint64()
        _lib.check(_lib.lib().zk_groth16_pk_load(ctypes.byref(st), self.shard[0], self.shard[1], ctypes.byref(h)))
        pkey._handles[self.shard] = h.value
        return h.value

    def prove_with(self, r: int, s: int, domain: EvalDomain, pkey: DerivedPKey, sol: Dict[Var, int]):
        from .groth16 import Proof
        circuit = domain.circuit
        out = (ctypes.c_uint8 * _lib.GROTH16_PROOF_OUT)()
        sol_b = sol if isinstance(sol, (bytes, bytearray)) else fr_vector(sol[k] for k in circuit.variables)
        rc = _lib.lib().zk_groth16_prove_r1cs(self._key_handle(pkey, circuit), domain.handle, sol_b,
                                              Fr.to_bytes(r), Fr.to_bytes(s), out)
        if rc == _lib.ZK_EREMAINDER:
            raise AssertionError("QAP.eval: remainder is not zero")        # QAP.ml:134
        _lib.check(rc)
        b = bytes(out)
        return Proof(Point(b[0:96], b[96:144]), Point(b[144:336], b[336:432]), Point(b[432:528], b[528:576]))

    def prove(self, rng: random.Random, domain: EvalDomain, pkey: DerivedPKey, sol):
        r = Fr.gen(rng)
        s = Fr.gen(rng)
        return self.prove_with(r, s, domain, pkey, sol)

    @staticmethod
    def free(pkey: DerivedPKey) -> None:
        for h in pkey._handles.values():
            _lib.check(_lib.lib().zk_key_free(h))
        pkey._handles.clear()


def closed_form_scalars(trapdoor, r: int, s: int, circuit: SparseCircuit, w: Sequence[int], sol: Dict[Var, int]):
    """(A, B, C) as single scalars of the generators — the trapdoor identity (SURVEY.md §8c iv),
    O(n + nnz) host integer work, exact at any size."""
    a, b, gm, d, t = trapdoor
    n = circuit.n
    lag, zt = lagrange_at(n, w, t)
    dinv = pow(d, -1, R)
    V = W = Y = L = 0
    mid = set(circuit.mids)
    for j, (lhs, l, rr) in enumerate(circuit.gates):
        lj = lag[j]
        vj = sum(c * sol[v] for v, c in l.items()) % R
        wj = sum(c * sol[v] for v, c in rr.items()) % R
        yj = sum(c * sol[v] for v, c in lhs.items()) % R
        V = (V + vj * lj) % R
        W = (W + wj * lj) % R
        Y = (Y + yj * lj) % R
        for v, c in l.items():
            if v in mid:
                L = (L + b * c % R * sol[v] % R * lj) % R
        for v, c in rr.items():
            if v in mid:
                L = (L + a * c % R * sol[v] % R * lj) % R
        for v, c in lhs.items():
            if v in mid:
                L = (L + c * sol[v] % R * lj) % R
    H = (V * W - Y) * pow(zt, -1, R) % R
    A = (a + V + r * d) % R
    B = (b + W + s * d) % R
    Cc = (L * dinv + H * zt % R * dinv + s * A + r * B - r * s % R * d) % R
    return A, B, Cc
