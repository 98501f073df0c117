"""Host mirror of ``Groth16.Make(C)`` (/root/reference/src/groth16/groth16.ml).

``prove`` keeps the reference signature ``prove rng qap pkey sol`` (groth16.ml:235-237): it
draws ``r`` then ``s`` from ``rng`` (groth16.ml:124-125), and makes ONE C-ABI call
(``zk_groth16_prove``) that runs QAP.eval and the three proof MSMs on the GPU.  The proving
key is uploaded once (``zk_groth16_pk_load``) and cached next to the ``pkey`` object.
"""

from __future__ import annotations

import ctypes
import random
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

from . import _lib, qap as Q
from .curve import Bls12_381, Fr, Point, R, Var, fr_vector
from .protocol import Circuit, ProtocolS


@dataclass
class PKey:
    """groth16.ml:24-34."""
    a: Point
    d1: Point
    ti1: List[Point]
    ltd_mid: Dict[Var, Point]
    tiztd: List[Point]
    b1: Point
    b2: Point
    d2: Point
    ti2: List[Point]
    _handles: Dict[Tuple[int, int], int] = field(default_factory=dict, repr=False, compare=False)
    _keep: list = field(default_factory=list, repr=False, compare=False)


@dataclass
class VKey:
    """groth16.ml:36-43."""
    one1: Point
    ltgm_io: Dict[Var, Point]
    one2: Point
    gm: Point
    d: Point
    ab: object = None


@dataclass
class Proof:
    """groth16.ml:110-114."""
    a: Point
    b: Point
    c: Point

    def to_compressed_bytes(self) -> bytes:
        """48 + 96 + 48 bytes: the form bit-identity is judged in (curve.ml:199,208)."""
        return self.a._comp + self.b._comp + self.c._comp


def _poly_apply(f, x):
    acc, xi = 0, 1
    for c in f:
        acc = (acc + c * xi) % R
        xi = xi * x % R
    return acc


class Make(ProtocolS):
    """``Groth16.Make(C : Curve.S)``; ``C`` defaults to (and must be) ``Curve.Bls12_381``."""

    def __init__(self, C=Bls12_381, shard: Tuple[int, int] = (0, 1)):
        self.C = C
        self.shard = shard

    # ---- groth16.ml:45-108 / 227-233 ---------------------------------------------------
    def keygen(self, rng: random.Random, circuit: Circuit, qap: Q.QAP) -> Tuple[PKey, VKey]:
        G1, G2 = self.C.G1, self.C.G2
        a, b, gm, d, t = (Fr.gen(rng) for _ in range(5))          # :51-55, in this order
        n = Q.degree(qap.target)
        v_io = set(circuit.inputs_public) | set(circuit.outputs)
        v_mid = set(circuit.mids)
        dinv, gminv = pow(d, -1, R), pow(gm, -1, R)
        lk = {k: (b * _poly_apply(qap.v[k], t) + a * _poly_apply(qap.w[k], t) + _poly_apply(qap.y[k], t)) % R
              for k in qap.v}                                        # :59-68 evaluated at tau
        ztd = _poly_apply(qap.target, t) * dinv % R
        mids = sorted(k for k in lk if k in v_mid)
        ios = sorted(k for k in lk if k in v_io)
        tpow = [pow(t, i, R) for i in range(n + 2)]
        # one fixed-base batch per group: singles first, then the lists
        s1 = [a, d, b] + tpow + [lk[k] * dinv % R for k in mids] + [tpow[i] * ztd % R for i in range(n - 1)] \
            + [lk[k] * gminv % R for k in ios]
        p1 = G1.fixed_base(s1)
        p2 = G2.fixed_base([b, d, gm] + tpow)
        o = 3
        ti1 = p1[o:o + n + 2]; o += n + 2
        ltd = p1[o:o + len(mids)]; o += len(mids)
        tiztd = p1[o:o + n - 1]; o += n - 1
        ltgm = p1[o:]
        pkey = PKey(a=p1[0], d1=p1[1], ti1=ti1, ltd_mid=dict(zip(mids, ltd)), tiztd=tiztd, b1=p1[2],
                    b2=p2[0], d2=p2[1], ti2=p2[3:])
        vkey = VKey(one1=G1.one, ltgm_io=dict(zip(ios, ltgm)), one2=G2.one, gm=p2[2], d=p2[1],
                    ab=self.C.Pairing.pairing(p1[0], p2[0]))               # :103
        return pkey, vkey

    # ---- device key --------------------------------------------------------------------
    def _key_handle(self, pkey: PKey, qap: Q.QAP) -> int:
        if self.shard in pkey._handles:
            return pkey._handles[self.shard]
        keys = qap.variables()
        pos = {k: i for i, k in enumerate(keys)}
        mids = sorted(pkey.ltd_mid)
        if any(k not in pos for k in mids):
            raise AssertionError("Variable of ltd_mid not found in the QAP")
        n = qap.n
        if len(pkey.ti1) < n or len(pkey.ti2) < n or len(pkey.tiztd) < n - 1:
            raise _lib.InvalidArgument(_lib.ZK_EARG, "apply_powers")     # curve.ml:116
        cat = lambda pts: b"".join(p.raw for p in pts)
        bufs = dict(a=pkey.a.raw, b1=pkey.b1.raw, d1=pkey.d1.raw, b2=pkey.b2.raw, d2=pkey.d2.raw,
                    ti1=cat(pkey.ti1[:n]), ti2=cat(pkey.ti2[:n]), tiztd=cat(pkey.tiztd[:n - 1]),
                    ltd_mid=cat(pkey.ltd_mid[k] for k in mids))
        cb = {k: ctypes.create_string_buffer(v, len(v)) for k, v in bufs.items()}
        idx = (ctypes.c_uint32 * max(len(mids), 1))(*[pos[k] for k in mids])
        st = _lib.Groth16PKeyStruct(n=n, m=len(keys), n_mid=len(mids), mid_index=ctypes.addressof(idx),
                                    **{k: ctypes.addressof(v) for k, v in cb.items()})
        h = ctypes.c_uint64()
        _lib.check(_lib.lib().zk_groth16_pk_load(ctypes.byref(st), self.shard[0], self.shard[1], ctypes.byref(h)))
        pkey._handles[self.shard] = h.value
        return h.value

    # ---- groth16.ml:235-237 ---------------------------------------------------------------
    def prove(self, rng: random.Random, qap: Q.QAP, pkey: PKey, sol: Dict[Var, int]) -> Proof:
        r = Fr.gen(rng)                                            # groth16.ml:124
        s = Fr.gen(rng)                                            # groth16.ml:125
        return self.prove_with(r, s, qap, pkey, sol)

    def prove_with(self, r: int, s: int, qap: Q.QAP, pkey: PKey, sol: Dict[Var, int]) -> Proof:
        keys = qap.variables()
        missing = [k for k in keys if k not in sol]
        if missing:
            raise AssertionError("Variable %s__%d not found" % missing[0])      # var.ml:72-78
        out = (ctypes.c_uint8 * _lib.GROTH16_PROOF_OUT)()
        rc = _lib.lib().zk_groth16_prove(self._key_handle(pkey, qap), qap.handle(),
                                         fr_vector(sol[k] for k in keys), Fr.to_bytes(r), Fr.to_bytes(s), out)
        if rc == _lib.ZK_EREMAINDER:
            raise AssertionError("QAP.eval: remainder is not zero")               # QAP.ml:134
        _lib.check(rc)
        b = bytes(out)
        a = Point(b[0:96], b[96:144])
        bb = Point(b[144:336], b[336:432])
        c = Point(b[432:528], b[528:576])
        return Proof(a, bb, c)

    # ---- groth16.ml:163-173 ---------------------------------------------------------------
    def verify(self, w_io: Dict[Var, int], vkey: VKey, proof: Proof) -> bool:
        """``e a b = ab + e (dot ltgm_io w_io) gm + e c d`` in GT.  The two pairings of the right-hand
        side move to the left (negated) so that the whole check is one pairing product — one final
        exponentiation — compared with the stored ``ab``."""
        G1, Pairing = self.C.G1, self.C.Pairing
        io = G1.dot(vkey.ltgm_io, w_io)                                  # Domain mismatch -> assert
        lhs = Pairing.product([(proof.a, proof.b), (io, vkey.gm), (proof.c, vkey.d)], [False, True, True])
        return self.C.GT.eq(lhs, vkey.ab)

    @staticmethod
    def free(pkey: PKey) -> None:
        for h in pkey._handles.values():
            _lib.check(_lib.lib().zk_key_free(h))
        pkey._handles.clear()
