"""CPU oracle — restatement of zukelang's prover path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED at the bls12-381 boundary (see ``oracle/bls12_381.py``): the
reference's tests carry no byte vectors, so this restatement is pinned by
(1) the reference's own polynomial KATs (``polynomial.ml:94-97,135-139,
180-209,232-246``) re-run over Fr, (2) the verifier equations
(``groth16.ml:163-173``, ``pinocchio.ml:254-420``) accepting every proof this
file makes, and (3) the closed-form trapdoor identities (``*_closed_form``).

Every function cites the reference lines it follows (paths relative to
``/root/reference/src``).  Fr elements are ints mod R, polynomials are lists of
Fr coefficients, lowest degree first (``lib/zk/polynomial.ml``), ``Var.t`` is a
``(str, int)`` tuple (``lib/zk/var.ml:4``) and ``Var.Map`` is a dict iterated
in sorted key order.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

from .bls12_381 import R, G1, G2, Group, fr_inv
from . import pairing as _pairing

Var = Tuple[str, int]
Poly = List[int]

ONE: Var = ("ONE", 1)            # circuit.ml:3  let one = Var.make "ONE"


# ==========================================================================
# lib/zk/polynomial.ml
# ==========================================================================
def poly_apply(f: Poly, x: int) -> int:
    """polynomial.ml:87-92."""
    acc, xi = 0, 1
    for a in f:
        acc = (acc + a * xi) % R
        xi = xi * x % R
    return acc


def poly_normalize(p: Poly) -> Poly:
    """polynomial.ml:100-107 — strip trailing zeros."""
    n = len(p)
    while n and p[n - 1] % R == 0:
        n -= 1
    return [c % R for c in p[:n]]


def poly_add(p1: Poly, p2: Poly) -> Poly:
    """polynomial.ml:109-115."""
    n = max(len(p1), len(p2))
    out = [0] * n
    for i, c in enumerate(p1):
        out[i] = c
    for i, c in enumerate(p2):
        out[i] = (out[i] + c) % R
    return poly_normalize(out)


def poly_sum(ps: Sequence[Poly]) -> Poly:
    """polynomial.ml:117 — List.fold_left add zero."""
    acc: Poly = []
    for p in ps:
        acc = poly_add(acc, p)
    return acc


def poly_mul_scalar(n: int, p: Poly) -> Poly:
    """polynomial.ml:119-120 (note: does NOT normalise)."""
    if n % R == 0:
        return []
    return [n * m % R for m in p]


def poly_neg(p: Poly) -> Poly:
    """polynomial.ml:122."""
    return [(-c) % R for c in p]


def poly_mul(p1: Poly, p2: Poly) -> Poly:
    """polynomial.ml:124-131 — schoolbook, sum of shifted scalar multiples."""
    return poly_sum([[0] * i + poly_mul_scalar(a, p2) for i, a in enumerate(p1)])


def poly_sub(p1: Poly, p2: Poly) -> Poly:
    return poly_add(p1, poly_neg(p2))


def poly_div_rem(p1: Poly, p2: Poly) -> Tuple[Poly, Poly]:
    """polynomial.ml:142-169 — long division, returns (quotient, remainder)."""
    p1 = poly_normalize(p1)
    p2 = poly_normalize(p2)
    assert p2 != []
    rp1 = p1[::-1]
    rp2 = p2[::-1]
    hd_inv = fr_inv(rp2[0])
    tl = rp2[1:]
    ds: Poly = []
    while len(rp1) >= len(rp2):
        d = rp1[0] * hd_inv % R
        rest = rp1[1:]
        for i, a2 in enumerate(tl):
            rest[i] = (rest[i] - d * a2) % R
        rp1 = rest
        ds.append(d)
    return ds[::-1], poly_normalize(rp1[::-1])


def poly_lagrange_basis(xs: Sequence[int]) -> List[Poly]:
    """polynomial.ml:212-226."""
    out = []
    for j, xj in enumerate(xs):
        acc: Poly = [1]
        for i, xi in enumerate(xs):
            if i == j:
                continue
            d = (xj - xi) % R
            assert d != 0
            di = fr_inv(d)
            acc = poly_mul(acc, [(-xi) * di % R, di])
        out.append(acc)
    return out


def poly_interpolate(xys: Sequence[Tuple[int, int]]) -> Poly:
    """polynomial.ml:228-230."""
    ls = poly_lagrange_basis([x for x, _ in xys])
    return poly_sum([poly_mul_scalar(y, l) for (_, y), l in zip(xys, ls)])


def poly_z(fs: Sequence[int]) -> Poly:
    """polynomial.ml:248-251 — prod (x - f)."""
    acc: Poly = [1]
    for f in fs:
        acc = poly_mul(acc, [(-f) % R, 1])
    return acc


def poly_degree(t: Poly) -> int:
    """polynomial.ml:253-255."""
    n = len(poly_normalize(t))
    return 0 if n == 0 else n - 1


def poly_is_zero(t: Poly) -> bool:
    """polynomial.ml:257."""
    return poly_normalize(t) == []


def poly_equal(t1: Poly, t2: Poly) -> bool:
    """polynomial.ml:259-269."""
    return poly_normalize(t1) == poly_normalize(t2)


def polynomial_self_test() -> None:
    """The reference's own KATs (run over Q at polynomial.ml:290-292), over Fr."""
    assert poly_apply([1, 2, 3, 4], 2) == 49                                   # :94-97
    assert poly_mul([1, 1, 1], [1, 1, 1, 1]) == [1, 2, 3, 3, 2, 1]             # :135-139
    for xys in ([(0, 1), (1, 2)], [(0, 10), (3, 9)], [(1, 3), (2, 2), (3, 4)]):  # :232-243
        f = poly_interpolate(xys)
        assert all(poly_apply(f, x) == y for x, y in xys)
    assert poly_normalize([0]) == []                                           # :245-246
    d, r = poly_div_rem([1, 2, 1], [1, 1])                                     # :182
    assert d == [1, 1] and r == []
    d, r = poly_div_rem([1, 1], [1, 2, 1])                                     # :186
    assert d == [] and r == [1, 1]
    import random
    rng = random.Random(1)
    for _ in range(200):                                                       # :190-209
        a = poly_normalize([rng.randrange(R) for _ in range(rng.randrange(20))])
        b = poly_normalize([rng.randrange(R) for _ in range(rng.randrange(20))])
        if b:
            d, r = poly_div_rem(a, b)
            assert len(r) < len(b)
            assert poly_add(poly_mul(d, b), r) == a


# ==========================================================================
# lib/zk/circuit.ml — just the data model
# ==========================================================================
Affine = Dict[Var, int]


@dataclass(frozen=True)
class Gate:
    """circuit.ml:75  { lhs; l; r } meaning  lhs = l * r."""
    lhs: Tuple[Tuple[Var, int], ...]
    l: Tuple[Tuple[Var, int], ...]
    r: Tuple[Tuple[Var, int], ...]

    @staticmethod
    def make(lhs: Affine, l: Affine, r: Affine) -> "Gate":
        f = lambda a: tuple(sorted((k, v % R) for k, v in a.items()))
        return Gate(f(lhs), f(l), f(r))

    def key(self):
        """circuit.ml:85-91 Gate.compare = lexicographic on (lhs, l, r) with
        Var.Map.compare F.compare (sorted binding lists, shorter prefix first).
        F.compare is taken to be integer order (unpinned)."""
        return (list(self.lhs), list(self.l), list(self.r))


@dataclass
class Circuit:
    """circuit.ml:108-113.  ``gates`` is kept in Gate.Set order."""
    gates: List[Gate]
    inputs_public: List[Var]
    outputs: List[Var]
    mids: List[Var]

    def __post_init__(self):
        uniq = {g: None for g in self.gates}
        self.gates = sorted(uniq, key=Gate.key)

    def vars(self) -> List[Var]:
        """circuit.ml:125-130."""
        s = set()
        for g in self.gates:
            for part in (g.lhs, g.l, g.r):
                s.update(k for k, _ in part)
        return sorted(s)

    def ios(self) -> List[Var]:
        """circuit.ml:132-134."""
        mids = set(self.mids)
        return [v for v in self.vars() if v not in mids]


def affine_eval(env: Dict[Var, int], a) -> int:
    """circuit.ml:60-62."""
    return sum(env[v] * c for v, c in a) % R


def circuit_check(circ: Circuit, sol: Dict[Var, int]) -> bool:
    return all(affine_eval(sol, g.lhs) == affine_eval(sol, g.l) * affine_eval(sol, g.r) % R
               for g in circ.gates)


# ==========================================================================
# lib/zk/QAP.ml
# ==========================================================================
@dataclass
class QAP:
    """QAP.ml:11-16."""
    v: Dict[Var, Poly]
    w: Dict[Var, Poly]
    y: Dict[Var, Poly]
    target: Poly


def qap_build(gates: List[Gate], literal: bool = False) -> QAP:
    """QAP.ml:18-94.  Gate ids rg = 0..n-1 in Gate.Set order (:22), evaluation
    points F.of_int rg (:84), target = prod (x - rg) (:92).

    ``literal=True`` follows the reference to the letter (one Lagrange basis
    per variable, O(m n^3)); the default computes the same polynomials from a
    single shared basis — used for the 2^10 configuration.  Both are compared
    in tests.
    """
    n = len(gates)
    vars_ = Circuit(gates, [], [], []).vars()
    xs = list(range(n))

    def matrix(sel):
        return {k: [dict(sel(g)).get(k, 0) for g in gates] for k in vars_}

    mv = matrix(lambda g: g.l)
    mw = matrix(lambda g: g.r)
    my = matrix(lambda g: g.lhs)
    if literal:
        interp = lambda col: poly_interpolate(list(zip(xs, col)))
        target = poly_z(xs)
    else:
        target = _z_fast(n)
        basis = _lagrange_basis_fast(n, target)

        def interp(col):
            acc = [0] * n
            for j, c in enumerate(col):
                if c:
                    bj = basis[j]
                    for i in range(n):
                        acc[i] = (acc[i] + c * bj[i]) % R
            return poly_normalize(acc)

    return QAP({k: interp(c) for k, c in mv.items()},
               {k: interp(c) for k, c in mw.items()},
               {k: interp(c) for k, c in my.items()},
               target)


def _z_fast(n: int) -> Poly:
    acc = [1]
    for j in range(n):
        nxt = [0] * (len(acc) + 1)
        for i, c in enumerate(acc):
            nxt[i] = (nxt[i] - j * c) % R
            nxt[i + 1] = (nxt[i + 1] + c) % R
        acc = nxt
    return acc


def lagrange_weights(n: int) -> List[int]:
    """w_j = 1 / prod_{i != j} (j - i) on the points 0..n-1."""
    fact = [1] * n
    for i in range(1, n):
        fact[i] = fact[i - 1] * i % R
    out = []
    for j in range(n):
        d = fact[j] * fact[n - 1 - j] % R
        if (n - 1 - j) & 1:
            d = (-d) % R
        out.append(fr_inv(d))
    return out


def _lagrange_basis_fast(n: int, target: Poly) -> List[Poly]:
    ws = lagrange_weights(n)
    out = []
    for j in range(n):
        # synthetic division of target by (x - j)
        q = [0] * n
        carry = 0
        for i in range(n, 0, -1):
            carry = (target[i] + carry * j) % R
            q[i - 1] = carry
        out.append([c * ws[j] % R for c in q])
    return out


def qap_eval(sol: Dict[Var, int], qap: QAP) -> Tuple[Poly, Poly]:
    """QAP.ml:120-135 — returns (p, h) with h * target = p."""
    def ev(vps):
        return poly_sum([poly_mul_scalar(sol[k], vps[k]) for k in sorted(vps)])

    v, w, y = ev(qap.v), ev(qap.w), ev(qap.y)
    p = poly_sub(poly_mul(v, w), y)
    h, rem = poly_div_rem(p, qap.target)
    assert poly_is_zero(rem)
    return p, h


def qap_combine(sol: Dict[Var, int], vps: Dict[Var, Poly]) -> Poly:
    """The V/W/Y coefficient vector of QAP.ml:121-131 (eval')."""
    return poly_sum([poly_mul_scalar(sol[k], vps[k]) for k in sorted(vps)])


# ==========================================================================
# lib/zk/curve.ml ExtendMap
# ==========================================================================
def g_sum_map(G: Group, m: Dict[Var, object], f):
    """curve.ml:91."""
    acc = None
    for k in sorted(m):
        acc = G.add(f(k, m[k]), acc)
    return acc


def g_dot(G: Group, m: Dict[Var, object], c: Dict[Var, int]):
    """curve.ml:94-103 — domains must match or the reference asserts false."""
    if set(m) != set(c):
        raise AssertionError("Domain mismatch")
    return g_sum_map(G, m, lambda k, mk: G.mul(mk, c[k]))


def g_powers(G: Group, d: int, s: int):
    """curve.ml:106-109 — d+1 points."""
    return [G.of_Fr(pow(s, i, R)) for i in range(d + 1)]


def g_apply_powers(G: Group, cs: Poly, xis: Sequence[object]):
    """curve.ml:112-118."""
    if len(cs) > len(xis):
        raise ValueError("apply_powers")          # Invalid_argument "apply_powers"
    acc = None
    for c, x in zip(cs, xis):
        acc = G.add(G.mul(x, c), acc)
    return acc


# ==========================================================================
# groth16/groth16.ml
# ==========================================================================
@dataclass
class Groth16PKey:
    """groth16.ml:24-34."""
    a: object
    d1: object
    ti1: list
    ltd_mid: Dict[Var, object]
    tiztd: list
    b1: object
    b2: object
    d2: object
    ti2: list


@dataclass
class Groth16VKey:
    """groth16.ml:36-43."""
    one1: object
    ltgm_io: Dict[Var, object]
    one2: object
    gm: object
    d: object
    ab: object


@dataclass
class Groth16Trapdoor:
    """The five Fr.gen draws of groth16.ml:51-55, in that order."""
    a: int
    b: int
    gm: int
    d: int
    t: int


def groth16_setup(td: Groth16Trapdoor, v_io: Sequence[Var], v_mid: Sequence[Var], n: int,
                  qap: QAP, with_ab: bool = True) -> Tuple[Groth16PKey, Groth16VKey]:
    """groth16.ml:45-108."""
    a, b, gm, d, t = td.a, td.b, td.gm, td.d, td.t
    z = qap.target
    l = {i: poly_add(poly_add(poly_mul_scalar(b, qap.v[i]), poly_mul_scalar(a, qap.w[i])),
                     qap.y[i]) for i in qap.v}                                  # :59-68
    dinv, gminv = fr_inv(d), fr_inv(gm)
    mid, io = set(v_mid), set(v_io)
    ztd = poly_apply(z, t) * dinv % R
    pkey = Groth16PKey(
        a=G1.of_Fr(a), d1=G1.of_Fr(d),
        ti1=g_powers(G1, n + 1, t),                                             # :73
        ltd_mid={k: G1.of_Fr(poly_apply(lk, t) * dinv % R) for k, lk in l.items() if k in mid},
        tiztd=[G1.of_Fr(pow(t, i, R) * ztd % R) for i in range(n - 1)],         # :80-83
        b1=G1.of_Fr(b), b2=G2.of_Fr(b), d2=G2.of_Fr(d),
        ti2=g_powers(G2, n + 1, t))                                             # :87
    vkey = Groth16VKey(
        one1=G1.one,
        ltgm_io={k: G1.of_Fr(poly_apply(lk, t) * gminv % R) for k, lk in l.items() if k in io},
        one2=G2.one, gm=G2.of_Fr(gm), d=G2.of_Fr(d),
        ab=_pairing.pairing(G1.of_Fr(a), G2.of_Fr(b)) if with_ab else None)     # :103
    return pkey, vkey


def groth16_keygen(td: Groth16Trapdoor, circuit: Circuit, qap: QAP, with_ab: bool = True):
    """groth16.ml:227-233."""
    d = poly_degree(qap.target)
    io = sorted(set(circuit.inputs_public) | set(circuit.outputs))
    return groth16_setup(td, io, circuit.mids, d, qap, with_ab)


def _sum_apply_powers(G: Group, ti, ps: Dict[Var, Poly], w: Dict[Var, int]):
    """groth16.ml:116-121."""
    acc = None
    for k in sorted(w):
        acc = G.add(G.mul(g_apply_powers(G, ps[k], ti), w[k]), acc)
    return acc


def groth16_prove(r: int, s: int, qap: QAP, pkey: Groth16PKey, sol: Dict[Var, int]):
    """groth16.ml:235-237 then :123-161.  (r, s) are the two Fr.gen draws of
    :124-125, r first.  Returns (a, b, c) as oracle points."""
    _p, h = qap_eval(sol, qap)
    a = G1.add(G1.add(pkey.a, _sum_apply_powers(G1, pkey.ti1, qap.v, sol)), G1.mul(pkey.d1, r))
    b = G2.add(G2.add(pkey.b2, _sum_apply_powers(G2, pkey.ti2, qap.w, sol)), G2.mul(pkey.d2, s))
    b1 = G1.add(G1.add(pkey.b1, _sum_apply_powers(G1, pkey.ti1, qap.w, sol)), G1.mul(pkey.d1, s))
    htztd = g_apply_powers(G1, h, pkey.tiztd)
    w_mid = {k: sol[k] for k in pkey.ltd_mid}                                   # restrict, :154
    c = g_dot(G1, pkey.ltd_mid, w_mid)
    c = G1.add(c, htztd)
    c = G1.add(c, G1.mul(a, s))
    c = G1.add(c, G1.mul(b1, r))
    c = G1.sub(c, G1.mul(pkey.d1, r * s % R))
    return a, b, c


def groth16_verify(w_io: Dict[Var, int], vkey: Groth16VKey, proof) -> bool:
    """groth16.ml:163-173."""
    a, b, c = proof
    lhs = _pairing.pairing(a, b)
    rhs = _pairing.gt_add(vkey.ab, _pairing.multi_pairing(
        [(g_dot(G1, vkey.ltgm_io, w_io), vkey.gm), (c, vkey.d)]))
    return lhs == rhs


def groth16_closed_form(td: Groth16Trapdoor, r: int, s: int, qap: QAP, circuit: Circuit,
                        sol: Dict[Var, int]):
    """Trapdoor identity (SURVEY.md §8c iv): each proof element as ONE scalar."""
    t, dinv = td.t, fr_inv(td.d)
    V = sum(sol[k] * poly_apply(qap.v[k], t) for k in qap.v) % R
    W = sum(sol[k] * poly_apply(qap.w[k], t) for k in qap.w) % R
    Y = sum(sol[k] * poly_apply(qap.y[k], t) for k in qap.y) % R
    Zt = poly_apply(qap.target, t)
    H = (V * W - Y) * fr_inv(Zt) % R
    A = (td.a + V + r * td.d) % R
    B = (td.b + W + s * td.d) % R
    mid = set(circuit.mids)
    L = sum(sol[k] * (td.b * poly_apply(qap.v[k], t) + td.a * poly_apply(qap.w[k], t)
                      + poly_apply(qap.y[k], t)) for k in qap.v if k in mid) % R
    C = (L * dinv + H * Zt * dinv + s * A + r * B - r * s * td.d) % R
    return G1.of_Fr(A), G2.of_Fr(B), G1.of_Fr(C)


def _batch_inverse(xs: Sequence[int]) -> List[int]:
    pre, acc = [], 1
    for x in xs:
        pre.append(acc)
        acc = acc * x % R
    inv = fr_inv(acc)
    out = [0] * len(xs)
    for i in range(len(xs) - 1, -1, -1):
        out[i] = inv * pre[i] % R
        inv = inv * xs[i] % R
    return out


def lagrange_values_at(n: int, t: int) -> Tuple[List[int], int]:
    """([L_j(t)] for j < n, Z(t)) on the reference's domain 0..n-1 (QAP.ml:84,92) for a point t
    outside it: the polynomials poly_lagrange_basis (polynomial.ml:212-230) builds, evaluated by
    L_j(t) = Z(t) / ((t - j) prod_{i != j} (j - i)) — O(n), usable at 2^20 gates."""
    fact = [1] * n
    for i in range(1, n):
        fact[i] = fact[i - 1] * i % R
    dens, z = [], 1
    for j in range(n):
        d = fact[j] * fact[n - 1 - j] % R            # |prod_{i != j} (j - i)|, sign (-1)^(n-1-j)
        if (n - 1 - j) & 1:
            d = R - d
        dens.append(d * ((t - j) % R) % R)
        z = z * ((t - j) % R) % R
    assert z != 0, "t lies on the evaluation domain"
    return [z * i % R for i in _batch_inverse(dens)], z


def groth16_closed_form_scalars(td: Groth16Trapdoor, r: int, s: int, circuit: Circuit,
                                sol: Dict[Var, int]) -> Tuple[int, int, int]:
    """The trapdoor identity of groth16_closed_form computed from the GATE LIST alone (no dense
    QAP.t, which cannot exist at 2^16+ gates): v_k(t) = sum_j l_j[k] L_j(t) by
    QAP.ml:81-86, so V(t) = sum_j <l_j, sol> L_j(t) and likewise W, Y and the mid-variable sum L.
    Returns the scalars (A, B, C) of the generators; O(n + nnz)."""
    lag, zt = lagrange_values_at(len(circuit.gates), td.t)
    mid = set(circuit.mids)
    V = W = Y = L = 0
    for g, lj in zip(circuit.gates, lag):
        vj, wj, yj = affine_eval(sol, g.l), affine_eval(sol, g.r), affine_eval(sol, g.lhs)
        V += vj * lj
        W += wj * lj
        Y += yj * lj
        lm = sum(td.b * c * sol[k] for k, c in g.l if k in mid)
        lm += sum(td.a * c * sol[k] for k, c in g.r if k in mid)
        lm += sum(c * sol[k] for k, c in g.lhs if k in mid)
        L += lm % R * lj
    V, W, Y, L = V % R, W % R, Y % R, L % R
    dinv = fr_inv(td.d)
    H = (V * W - Y) * fr_inv(zt) % R
    A = (td.a + V + r * td.d) % R
    B = (td.b + W + s * td.d) % R
    C = (L * dinv + H * zt * dinv + s * A + r * B - r * s * td.d) % R
    return A, B, C


# ==========================================================================
# pinocchio/pinocchio.ml
# ==========================================================================
@dataclass
class PinocchioTrapdoor:
    """The eight Fr.gen draws of pinocchio.ml:83-91, in that order."""
    rv: int
    rw: int
    s: int
    av: int
    aw: int
    ay: int
    b: int
    gm: int


PINOCCHIO_PROOF_FIELDS = ("vv", "ww", "yy", "h", "vavv", "waww", "yayy", "bvwy")  # :195-208
PINOCCHIO_PROOF_GROUPS = ("G1", "G2", "G1", "G1", "G1", "G2", "G1", "G1")


def pinocchio_keygen(td: PinocchioTrapdoor, circuit: Circuit, qap: QAP):
    """pinocchio.ml:77-189 KeyGen.generate.  pkey / vkey are dicts keyed by the
    reference's record field names."""
    imid = list(circuit.mids)
    nio = circuit.ios()
    m = circuit.vars()
    d = poly_degree(qap.target)
    rv, rw, s, av, aw, ay, b, gm = td.rv, td.rw, td.s, td.av, td.aw, td.ay, td.b, td.gm
    ry = rv * rw % R
    gv, gw, gw2, gy, gy2 = G1.of_Fr(rv), G1.of_Fr(rw), G2.of_Fr(rw), G1.of_Fr(ry), G2.of_Fr(ry)
    t = poly_apply(qap.target, s)

    def map_apply_s(G, gu, u, keys):                                           # :100-105
        return {k: G.mul(gu, poly_apply(u[k], s)) for k in keys}

    vv = map_apply_s(G1, gv, qap.v, imid)
    ww1 = map_apply_s(G1, gw, qap.w, imid)
    ww = map_apply_s(G2, gw2, qap.w, imid)
    yy = map_apply_s(G1, gy, qap.y, imid)
    mul_map = lambda G, mp, a: {k: G.mul(g, a) for k, g in mp.items()}
    pkey = dict(
        vv=vv, ww=ww, yy=yy,
        vav=mul_map(G1, vv, av), waw=mul_map(G2, ww, aw), yay=mul_map(G1, yy, ay),
        si=g_powers(G1, d, s), si2=g_powers(G2, d, s),
        bvwy={k: G1.mul(G1.add(G1.add(vv[k], ww1[k]), yy[k]), b) for k in imid},
        vt=G1.mul(gv, t), wt=G2.mul(gw2, t), yt=G1.mul(gy, t),
        vavt=G1.mul(G1.mul(gv, av), t), wawt=G2.mul(G2.mul(gw2, aw), t),
        yayt=G1.mul(G1.mul(gy, ay), t),
        vbt=G1.mul(G1.mul(gv, b), t), wbt=G1.mul(G1.mul(gw, b), t), ybt=G1.mul(G1.mul(gy, b), t),
        v_all=map_apply_s(G1, G1.one, qap.v, m), w_all=map_apply_s(G1, G1.one, qap.w, m))
    gm1, gm2 = G1.of_Fr(gm), G2.of_Fr(gm)
    vkey = dict(
        one=G1.one, one2=G2.one, av=G2.of_Fr(av), aw=G1.of_Fr(aw), ay=G2.of_Fr(ay),
        gm2=gm2, bgm=G1.mul(gm1, b), bgm2=G2.mul(gm2, b), yt=G2.mul(gy2, t),
        vv_io=map_apply_s(G1, gv, qap.v, nio), ww_io=map_apply_s(G2, gw2, qap.w, nio),
        yy_io=map_apply_s(G1, gy, qap.y, nio))
    return pkey, vkey


def pinocchio_compute(pkey, sol: Dict[Var, int], h_poly: Poly):
    """pinocchio.ml:210-248 Compute.f (NonZK)."""
    c_mid = {k: sol[k] for k in pkey["vv"]}
    return dict(
        vv=g_dot(G1, pkey["vv"], c_mid), ww=g_dot(G2, pkey["ww"], c_mid),
        yy=g_dot(G1, pkey["yy"], c_mid), h=g_apply_powers(G1, h_poly, pkey["si"]),
        vavv=g_dot(G1, pkey["vav"], c_mid), waww=g_dot(G2, pkey["waw"], c_mid),
        yayy=g_dot(G1, pkey["yay"], c_mid), bvwy=g_dot(G1, pkey["bvwy"], c_mid))


def pinocchio_zk_compute(dv: int, dw: int, dy: int, target: Poly, pkey, sol: Dict[Var, int],
                         h_poly: Poly):
    """pinocchio.ml:427-514 ZKCompute.f; (dv, dw, dy) are the draws of :428-430."""
    t = g_apply_powers(G1, target, pkey["si"])
    c_mid = {k: sol[k] for k in pkey["vv"]}
    vv = G1.add(g_dot(G1, pkey["vv"], c_mid), G1.mul(pkey["vt"], dv))
    ww = G2.add(g_dot(G2, pkey["ww"], c_mid), G2.mul(pkey["wt"], dw))
    yy = G1.add(g_dot(G1, pkey["yy"], c_mid), G1.mul(pkey["yt"], dy))
    h = g_apply_powers(G1, h_poly, pkey["si"])
    v_all = g_dot(G1, pkey["v_all"], sol)
    w_all = g_dot(G1, pkey["w_all"], sol)
    hp = G1.add(h, G1.mul(v_all, dw))
    hp = G1.add(hp, G1.mul(w_all, dv))
    hp = G1.add(hp, G1.mul(G1.mul(t, dv), dw))
    hp = G1.sub(hp, G1.mul(G1.one, dy))
    vavv = G1.add(g_dot(G1, pkey["vav"], c_mid), G1.mul(pkey["vavt"], dv))
    waww = G2.add(g_dot(G2, pkey["waw"], c_mid), G2.mul(pkey["wawt"], dw))
    yayy = G1.add(g_dot(G1, pkey["yay"], c_mid), G1.mul(pkey["yayt"], dy))
    bvwy = g_dot(G1, pkey["bvwy"], c_mid)
    bvwy = G1.add(bvwy, G1.mul(pkey["vbt"], dv))
    bvwy = G1.add(bvwy, G1.mul(pkey["wbt"], dw))
    bvwy = G1.add(bvwy, G1.mul(pkey["ybt"], dy))
    return dict(vv=vv, ww=ww, yy=yy, h=hp, vavv=vavv, waww=waww, yayy=yayy, bvwy=bvwy)


def pinocchio_prove(qap: QAP, pkey, sol, zk: Optional[Tuple[int, int, int]] = None):
    """NonZK.prove pinocchio.ml:536-538 / ZK.prove :559-561."""
    _p, h = qap_eval(sol, qap)
    if zk is None:
        return pinocchio_compute(pkey, sol, h)
    return pinocchio_zk_compute(zk[0], zk[1], zk[2], qap.target, pkey, sol, h)


def pinocchio_verify(ios: Dict[Var, int], vkey, proof) -> bool:
    """pinocchio.ml:254-420 Verify.f (the asserts become a False return)."""
    e = _pairing.pairing
    mul = _pairing.gt_add
    if e(proof["vv"], vkey["av"]) != e(proof["vavv"], vkey["one2"]):
        return False
    if e(vkey["aw"], proof["ww"]) != e(vkey["one"], proof["waww"]):
        return False
    if e(proof["yy"], vkey["ay"]) != e(proof["yayy"], vkey["one2"]):
        return False
    if e(proof["bvwy"], vkey["gm2"]) != _pairing.multi_pairing(
            [(proof["vv"], vkey["bgm2"]), (vkey["bgm"], proof["ww"]), (proof["yy"], vkey["bgm2"])]):
        return False
    assert set(ios) == set(vkey["vv_io"]) == set(vkey["ww_io"]) == set(vkey["yy_io"])
    vio = g_sum_map(G1, ios, lambda k, ck: G1.mul(vkey["vv_io"][k], ck))
    wio = g_sum_map(G2, ios, lambda k, ck: G2.mul(vkey["ww_io"][k], ck))
    yio = g_sum_map(G1, ios, lambda k, ck: G1.mul(vkey["yy_io"][k], ck))
    # e(vio+vv, wio+ww) - e(yio+yy, one2) = e(h, yt), written multiplicatively
    lhs = e(G1.add(vio, proof["vv"]), G2.add(wio, proof["ww"]))
    rhs = mul(e(proof["h"], vkey["yt"]), e(G1.add(yio, proof["yy"]), vkey["one2"]))
    return lhs == rhs


def pinocchio_closed_form(td: PinocchioTrapdoor, qap: QAP, circuit: Circuit, sol,
                          zk: Optional[Tuple[int, int, int]] = None):
    """Trapdoor identity for the eight Pinocchio elements."""
    s = td.s
    ry = td.rv * td.rw % R
    mid = list(circuit.mids)
    ev = lambda u, keys: sum(sol[k] * poly_apply(u[k], s) for k in keys) % R
    vm, wm, ym = ev(qap.v, mid), ev(qap.w, mid), ev(qap.y, mid)
    va, wa, ya = ev(qap.v, qap.v), ev(qap.w, qap.w), ev(qap.y, qap.y)
    t = poly_apply(qap.target, s)
    h = (va * wa - ya) * fr_inv(t) % R
    dv, dw, dy = zk if zk is not None else (0, 0, 0)
    vm, wm, ym = (vm + dv * t) % R, (wm + dw * t) % R, (ym + dy * t) % R
    if zk is not None:
        h = (h + va * dw + wa * dv + t * dv * dw - dy) % R
    return dict(
        vv=G1.of_Fr(td.rv * vm), ww=G2.of_Fr(td.rw * wm), yy=G1.of_Fr(ry * ym), h=G1.of_Fr(h),
        vavv=G1.of_Fr(td.rv * td.av * vm), waww=G2.of_Fr(td.rw * td.aw * wm),
        yayy=G1.of_Fr(ry * td.ay * ym),
        bvwy=G1.of_Fr(td.b * (td.rv * vm + td.rw * wm + ry * ym)))


# ==========================================================================
# Synthetic circuits for BASELINE.json's configurations
# ==========================================================================
def circuit_cubic() -> Tuple[Circuit, callable]:
    """Config 1: README program x*x*x + x + 3 (README.md:44-46, test.ml:194-197).

    Three gates as Comp.compile emits them (comp.ml:233-244 for Mul,
    fix_output :448-473): c_a = x*x ; c_b = c_a*x ; v = (c_b + x + 3 ONE)*(1 ONE).
    Variable numbering follows Var.make's global counter with ONE made first
    (circuit.ml:3); exact numbers are unpinned.
    """
    x, ca, cb, v = ("input", 2), ("_tmp", 3), ("_tmp", 4), ("v", 5)
    gates = [Gate.make({ca: 1}, {x: 1}, {x: 1}),
             Gate.make({cb: 1}, {ca: 1}, {x: 1}),
             Gate.make({v: 1}, {cb: 1, x: 1, ONE: 3}, {ONE: 1})]
    circ = Circuit(gates, inputs_public=[ONE], outputs=[v], mids=[x, ca, cb])

    def witness(xv: int) -> Dict[Var, int]:
        xv %= R
        return {ONE: 1, x: xv, ca: xv * xv % R, cb: pow(xv, 3, R), v: (pow(xv, 3, R) + xv + 3) % R}

    return circ, witness


def circuit_mulchain(n: int) -> Tuple[Circuit, callable]:
    """Configs 3/5: multiply chain c_0 = x ; c_{i+1} = c_i * x, n gates, plus a
    final output gate as fix_output does.  n >= 2."""
    x = ("input", 2)
    cs = [x] + [("_tmp", 3 + i) for i in range(n - 1)]
    out = ("v", 3 + n)
    gates = [Gate.make({cs[i + 1]: 1}, {cs[i]: 1}, {x: 1}) for i in range(n - 1)]
    gates.append(Gate.make({out: 1}, {cs[-1]: 1, ONE: 3}, {ONE: 1}))
    circ = Circuit(gates, inputs_public=[ONE], outputs=[out], mids=cs)

    def witness(xv: int) -> Dict[Var, int]:
        xv %= R
        sol = {ONE: 1, x: xv}
        cur = xv
        for i in range(1, n):
            cur = cur * xv % R
            sol[cs[i]] = cur
        sol[out] = (cur + 3) % R
        return sol

    return circ, witness


def circuit_pair_case(n_gates: int, seed: int = 0x50494E4F) -> Tuple[Circuit, callable]:
    """Config 2: a synthetic DSL-shaped circuit of exactly ``n_gates`` gates built by
    repeating the gate shapes Comp.compile emits for the pair / case programs of
    test.ml:216-247: boolean tags (b*b = b, comp.ml:298-324), tag-selected
    branches (r = b*(l - r0) + r0 written as a Mul gate) and field products on
    the projected components, each block on fresh inputs.
    """
    import random
    rng = random.Random(seed)
    gates: List[Gate] = []
    mids: List[Var] = []
    plan = []
    ctr = [1]

    def fresh(name):
        ctr[0] += 1
        return (name, ctr[0])

    outs: List[Var] = []
    while len(gates) < n_gates:
        b, p, q = fresh("tag"), fresh("fst"), fresh("snd")
        sel, prod = fresh("case"), fresh("mul")
        k = rng.randrange(1, 1 << 16)
        blk = [Gate.make({b: 1}, {b: 1}, {b: 1}),                       # tag is boolean
               Gate.make({sel: 1, q: R - 1}, {b: 1}, {p: 1, q: R - 1}),  # sel = b ? p : q
               Gate.make({prod: 1}, {sel: 1, ONE: k}, {p: 1, q: 1})]     # (sel + k) * (p + q)
        blk = blk[:n_gates - len(gates)]
        gates.extend(blk)
        mids.extend([b, p, q, sel])
        used = len(blk)
        if used == 3:
            mids.append(prod)
        plan.append((b, p, q, sel, prod, k, used))
    live = set(Circuit(gates, [], [], []).vars())
    mids = [v for v in mids if v in live]
    out = mids.pop()                     # the last live variable is the program output
    circ = Circuit(gates, inputs_public=[ONE], outputs=[out], mids=mids)
    assert len(circ.gates) == n_gates

    def witness(seed2: int) -> Dict[Var, int]:
        r2 = random.Random(seed2)
        sol = {ONE: 1}
        for b, p, q, sel, prod, k, used in plan:
            bv, pv, qv = r2.randrange(2), r2.randrange(R), r2.randrange(R)
            sv = pv if bv else qv
            sol[b], sol[p], sol[q], sol[sel] = bv, pv, qv, sv
            if used == 3:
                sol[prod] = (sv + k) * (pv + qv) % R
        live = set(circ.vars())
        return {k: v for k, v in sol.items() if k in live}

    return circ, witness


def circuit_random_r1cs(n: int, seed: int = 0x47524F54) -> Tuple[Circuit, callable]:
    """Config 5's alternative workload: a seeded random sparse R1CS of n gates,
    gate i:  z_i = (a x_p + b x_q + k ONE) * (c x_s + d x_u)  with p, q, s, u drawn among the
    variables defined before gate i and a, b, c, d, k uniform in Fr (about 3 non-zeros per row of
    each of the l / r matrices' union).  Unlike the multiply chain, W(j) differs at every gate, so
    the G2 B-query sees uniformly spread scalars.  Satisfiable by construction: the witness
    evaluates the gates in order."""
    import random
    rng = random.Random(seed + n)
    xs = [("input", 2), ("input", 3)]
    zs = [("_tmp", 4 + i) for i in range(n - 1)] + [("v", 3 + n)]
    defined = list(xs)
    gates, plan = [], []
    for i in range(n):
        p, q, s_, u = (defined[rng.randrange(len(defined))] for _ in range(4))
        a, b, c, d, k = (rng.randrange(1, R) for _ in range(5))
        l, r = {}, {}
        for var, co in ((p, a), (q, b), (ONE, k)):
            l[var] = (l.get(var, 0) + co) % R
        for var, co in ((s_, c), (u, d)):
            r[var] = (r.get(var, 0) + co) % R
        gates.append(Gate.make({zs[i]: 1}, l, r))
        plan.append((zs[i], sorted(l.items()), sorted(r.items())))
        defined.append(zs[i])
    circ = Circuit(gates, inputs_public=[ONE], outputs=[zs[-1]], mids=xs + zs[:-1])
    assert len(circ.gates) == n

    def witness(seed2: int) -> Dict[Var, int]:
        r2 = random.Random(seed2)
        sol = {ONE: 1, xs[0]: r2.randrange(R), xs[1]: r2.randrange(R)}
        for z, l, r in plan:
            sol[z] = sum(c * sol[v] for v, c in l) % R * (sum(c * sol[v] for v, c in r) % R) % R
        live = set(circ.vars())
        return {k: v for k, v in sol.items() if k in live}

    return circ, witness


def self_check() -> None:
    polynomial_self_test()
    circ, wit = circuit_cubic()
    assert len(circ.gates) == 3 and len(circ.vars()) == 5
    sol = wit(7)
    assert circuit_check(circ, sol)
    qap_l = qap_build(circ.gates, literal=True)
    qap_f = qap_build(circ.gates)
    assert qap_l == qap_f
    td = Groth16Trapdoor(11, 22, 33, 44, 55)
    pk, vk = groth16_keygen(td, circ, qap_f)
    assert len(pk.ti1) == 5 and len(pk.tiztd) == 2 and len(pk.ltd_mid) == 3
    proof = groth16_prove(101, 202, qap_f, pk, sol)
    assert proof == groth16_closed_form(td, 101, 202, qap_f, circ, sol)
    pub = {k: sol[k] for k in vk.ltgm_io}
    assert groth16_verify(pub, vk, proof)
    bad = dict(pub)
    bad[ONE] = 2
    assert not groth16_verify(bad, vk, proof)
    # the gate-list closed form (used at sizes with no dense QAP) equals the dense one
    for c2, w2 in (circuit_cubic(), circuit_mulchain(9), circuit_random_r1cs(12)):
        s2 = w2(5)
        assert circuit_check(c2, s2)
        A, B, C = groth16_closed_form_scalars(td, 101, 202, c2, s2)
        assert (G1.of_Fr(A), G2.of_Fr(B), G1.of_Fr(C)) == groth16_closed_form(td, 101, 202, qap_build(c2.gates), c2, s2)


if __name__ == "__main__":
    import time
    t0 = time.time()
    self_check()
    print("oracle/zk.py self-check OK in %.1fs" % (time.time() - t0))
