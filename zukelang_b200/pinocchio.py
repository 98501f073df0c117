"""Host mirror of ``Pinocchio.Make(C).{NonZK, ZK}`` (/root/reference/src/pinocchio/pinocchio.ml).

``NonZK.prove _rng qap pkey sol`` (pinocchio.ml:536-538) and ``ZK.prove rng qap pkey sol``
(:559-561, drawing dv, dw, dy at :428-430) each make one ``zk_pinocchio_prove`` call; the eight
proof elements come back in the record order of pinocchio.ml:195-208.
"""

from __future__ import annotations

import ctypes
import random
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from . import _lib, qap as Q
from .curve import Bls12_381, Fr, Point, R, Var, fr_vector
from .groth16 import _poly_apply
from .protocol import Circuit, ProtocolS

G1_LISTS = ("vv", "yy", "vav", "yay", "bvwy")
G2_LISTS = ("ww", "waw")
G1_SINGLES = ("vt", "yt", "vavt", "yayt", "vbt", "wbt", "ybt")
G2_SINGLES = ("wt", "wawt")
PROOF_FIELDS = ("vv", "ww", "yy", "h", "vavv", "waww", "yayy", "bvwy")
PROOF_IS_G2 = (False, True, False, False, False, True, False, False)


@dataclass
class PKey:
    """pinocchio.ml:37-60."""
    vv: Dict[Var, Point]
    ww: Dict[Var, Point]
    yy: Dict[Var, Point]
    vav: Dict[Var, Point]
    waw: Dict[Var, Point]
    yay: Dict[Var, Point]
    si: List[Point]
    bvwy: Dict[Var, Point]
    si2: List[Point]
    vt: Point
    wt: Point
    yt: Point
    vavt: Point
    wawt: Point
    yayt: Point
    vbt: Point
    wbt: Point
    ybt: Point
    v_all: Dict[Var, Point]
    w_all: Dict[Var, Point]
    _handles: Dict[Tuple[int, int], int] = field(default_factory=dict, repr=False, compare=False)


@dataclass
class Proof:
    """pinocchio.ml:195-208."""
    vv: Point
    ww: Point
    yy: Point
    h: Point
    vavv: Point
    waww: Point
    yayy: Point
    bvwy: Point

    def to_compressed_bytes(self) -> bytes:
        return b"".join(getattr(self, f)._comp for f in PROOF_FIELDS)


class _Base(ProtocolS):
    zk = False

    def __init__(self, C=Bls12_381, shard: Tuple[int, int] = (0, 1)):
        self.C = C
        self.shard = shard

    # ---- pinocchio.ml:77-189 KeyGen.generate ----------------------------------------------
    def keygen(self, rng: random.Random, circuit: Circuit, qap: Q.QAP):
        G1, G2 = self.C.G1, self.C.G2
        rv, rw, s, av, aw, ay, b, gm = (Fr.gen(rng) for _ in range(8))      # :83-91, in this order
        ry = rv * rw % R
        imid = sorted(circuit.mids)
        nio = circuit.ios()
        m = sorted(circuit.vars)
        d = Q.degree(qap.target)
        t = _poly_apply(qap.target, s)
        ev = {name: {k: _poly_apply(p[k], s) for k in m} for name, p in (("v", qap.v), ("w", qap.w), ("y", qap.y))}
        spow = [pow(s, i, R) for i in range(d + 1)]
        # G1 scalars, grouped:  lists over imid, lists over all, io lists, singles, powers
        L = lambda f: [f(k) % R for k in imid]
        g1 = (L(lambda k: rv * ev["v"][k]) + L(lambda k: ry * ev["y"][k]) + L(lambda k: rv * av * ev["v"][k])
              + L(lambda k: ry * ay * ev["y"][k])
              + L(lambda k: b * (rv * ev["v"][k] + rw * ev["w"][k] + ry * ev["y"][k]))
              + [ev["v"][k] for k in m] + [ev["w"][k] for k in m]
              + [rv * ev["v"][k] % R for k in nio] + [ry * ev["y"][k] % R for k in nio]
              + [rv * t, ry * t, rv * av * t, ry * ay * t, rv * b * t, rw * b * t, ry * b * t, aw, gm * b]
              + spow)
        g2 = (L(lambda k: rw * ev["w"][k]) + L(lambda k: rw * aw * ev["w"][k]) + [rw * ev["w"][k] % R for k in nio]
              + [rw * t, rw * aw * t, av, ay, gm, gm * b, ry * t] + spow)
        p1 = G1.fixed_base([x % R for x in g1])
        p2 = G2.fixed_base([x % R for x in g2])
        nm, na, ni = len(imid), len(m), len(nio)
        cut = lambda arr, o, n: (arr[o:o + n], o + n)
        o = 0
        vv, o = cut(p1, o, nm); yy, o = cut(p1, o, nm); vav, o = cut(p1, o, nm); yay, o = cut(p1, o, nm)
        bvwy, o = cut(p1, o, nm); v_all, o = cut(p1, o, na); w_all, o = cut(p1, o, na)
        vv_io, o = cut(p1, o, ni); yy_io, o = cut(p1, o, ni)
        (vt, yt, vavt, yayt, vbt, wbt, ybt, aw1, bgm), o = cut(p1, o, 9)
        si = p1[o:]
        o = 0
        ww, o = cut(p2, o, nm); waw, o = cut(p2, o, nm); ww_io, o = cut(p2, o, ni)
        (wt, wawt, av2, ay2, gm2, bgm2, yt2), o = cut(p2, o, 7)
        si2 = p2[o:]
        z = lambda keys, pts: dict(zip(keys, pts))
        pkey = PKey(vv=z(imid, vv), ww=z(imid, ww), yy=z(imid, yy), vav=z(imid, vav), waw=z(imid, waw),
                    yay=z(imid, yay), si=si, bvwy=z(imid, bvwy), si2=si2, vt=vt, wt=wt, yt=yt, vavt=vavt,
                    wawt=wawt, yayt=yayt, vbt=vbt, wbt=wbt, ybt=ybt, v_all=z(m, v_all), w_all=z(m, w_all))
        vkey = dict(one=G1.one, one2=G2.one, av=av2, aw=aw1, ay=ay2, gm2=gm2, bgm=bgm, bgm2=bgm2, yt=yt2,
                    vv_io=z(nio, vv_io), ww_io=z(nio, ww_io), yy_io=z(nio, yy_io))
        return pkey, vkey

    # ---- device key ------------------------------------------------------------------------
    def _key_handle(self, pkey: PKey, qap: Q.QAP) -> int:
        if self.shard in pkey._handles:
            return pkey._handles[self.shard]
        keys = qap.variables()
        pos = {k: i for i, k in enumerate(keys)}
        mids = sorted(pkey.vv)
        for name in G1_LISTS + G2_LISTS:
            if set(getattr(pkey, name)) != set(mids):
                raise AssertionError("Domain mismatch")                    # curve.ml:96-100
        if set(pkey.v_all) != set(keys) or set(pkey.w_all) != set(keys):
            raise AssertionError("Domain mismatch")
        n = qap.n
        if len(pkey.si) < n + 1:
            raise _lib.InvalidArgument(_lib.ZK_EARG, "apply_powers")     # curve.ml:116
        cat = lambda pts: b"".join(p.raw for p in pts)
        bufs = {name: cat(getattr(pkey, name)[k] for k in mids) for name in G1_LISTS + G2_LISTS}
        bufs["si"] = cat(pkey.si[:n + 1])
        bufs["v_all"] = cat(pkey.v_all[k] for k in keys)
        bufs["w_all"] = cat(pkey.w_all[k] for k in keys)
        bufs["one"] = self.C.G1.one.raw
        for name in G1_SINGLES + G2_SINGLES:
            bufs[name] = getattr(pkey, name).raw
        cb = {k: ctypes.create_string_buffer(v, len(v)) for k, v in bufs.items()}
        idx = (ctypes.c_uint32 * max(len(mids), 1))(*[pos[k] for k in mids])
        st = _lib.PinocchioPKeyStruct(n=n, m=len(keys), n_mid=len(mids), mid_index=ctypes.addressof(idx),
                                      **{k: ctypes.addressof(v) for k, v in cb.items()})
        h = ctypes.c_uint64()
        _lib.check(_lib.lib().zk_pinocchio_pk_load(ctypes.byref(st), self.shard[0], self.shard[1], ctypes.byref(h)))
        pkey._handles[self.shard] = h.value
        return h.value

    def _prove(self, d: Optional[Tuple[int, int, int]], qap: Q.QAP, pkey: PKey, sol: Dict[Var, int]) -> Proof:
        keys = qap.variables()
        missing = [k for k in keys if k not in sol]
        if missing:
            raise AssertionError("Variable %s__%d not found" % missing[0])
        out = (ctypes.c_uint8 * _lib.PINOCCHIO_PROOF_OUT)()
        dbuf = fr_vector(d) if d is not None else None
        rc = _lib.lib().zk_pinocchio_prove(self._key_handle(pkey, qap), qap.handle(),
                                           fr_vector(sol[k] for k in keys), dbuf, out)
        if rc == _lib.ZK_EREMAINDER:
            raise AssertionError("QAP.eval: remainder is not zero")
        _lib.check(rc)
        b, o, pts = bytes(out), 0, []
        for is2 in PROOF_IS_G2:
            raw, comp = (192, 96) if is2 else (96, 48)
            pts.append(Point(b[o:o + raw], b[o + raw:o + raw + comp]))
            o += raw + comp
        return Proof(*pts)

    # ---- pinocchio.ml:254-420 Verify.f -------------------------------------------------------
    def verify(self, ios: Dict[Var, int], vkey: dict, proof: Proof) -> bool:
        """The four knowledge-of-coefficient checks are ``assert``s in the reference (:296, :311,
        :326, :380): they raise here too; the divisibility check (:416-419) is the boolean result.
        Every GT equation ``lhs = rhs`` is evaluated as one pairing product ``lhs - rhs = 0``."""
        G1, G2, GT = self.C.G1, self.C.G2, self.C.GT
        vio = G1.dot(vkey["vv_io"], ios)                                  # Domain mismatch -> assert (:383,:392,:401)
        wio = G2.dot(vkey["ww_io"], ios)
        yio = G1.dot(vkey["yy_io"], ios)
        # all five equations in one device call: Miller loops and final exponentiations side by side
        checks = self.C.Pairing.products([
            ([(proof.vv, vkey["av"]), (proof.vavv, vkey["one2"])], [False, True]),                       # :296
            ([(vkey["aw"], proof.ww), (vkey["one"], proof.waww)], [False, True]),                        # :311
            ([(proof.yy, vkey["ay"]), (proof.yayy, vkey["one2"])], [False, True]),                       # :326
            ([(proof.bvwy, vkey["gm2"]), (proof.vv, vkey["bgm2"]), (vkey["bgm"], proof.ww),
              (proof.yy, vkey["bgm2"])], [False, True, True, True]),                                     # :380
            ([(G1.add(vio, proof.vv), G2.add(wio, proof.ww)), (G1.add(yio, proof.yy), vkey["one2"]),
              (proof.h, vkey["yt"])], [False, True, True]),                                              # :416-419
        ])
        for ok, what in zip(checks[:4], ("KC check of vv", "KC check of ww", "KC check of yy", "same-coefficient check")):
            if not GT.eq(ok, GT.zero):
                raise AssertionError("Pinocchio.verify: %s failed" % what)
        return GT.eq(checks[4], GT.zero)

    @staticmethod
    def free(pkey: PKey) -> None:
        for h in pkey._handles.values():
            _lib.check(_lib.lib().zk_key_free(h))
        pkey._handles.clear()


class NonZK(_Base):
    """pinocchio.ml:517-542."""

    def prove(self, _rng, qap, pkey, sol) -> Proof:
        return self._prove(None, qap, pkey, sol)


class ZK(_Base):
    """pinocchio.ml:544-564."""
    zk = True

    def prove(self, rng: random.Random, qap, pkey, sol) -> Proof:
        dv = Fr.gen(rng)                                            # pinocchio.ml:428
        dw = Fr.gen(rng)                                            # :429
        dy = Fr.gen(rng)                                            # :430
        return self._prove((dv, dw, dy), qap, pkey, sol)

    def prove_with(self, d, qap, pkey, sol) -> Proof:
        return self._prove(tuple(d), qap, pkey, sol)


class Make:
    """``Pinocchio.Make(C)`` exposing ``NonZK`` and ``ZK`` (pinocchio.mli:3-15)."""

    def __init__(self, C=Bls12_381, shard: Tuple[int, int] = (0, 1)):
        self.NonZK = NonZK(C, shard)
        self.ZK = ZK(C, shard)
