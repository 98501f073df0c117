"""Host mirror of ``Curve.Bls12_381`` (/root/reference/src/lib/zk/curve.ml:74-221).

``Fr`` values are Python ints mod r (host-side scalar bookkeeping, as in the OCaml
host).  ``G1`` / ``G2`` values are opaque byte-backed points (the counterpart of the
reference's custom blocks) and EVERY group operation — ``+``, ``*``, ``~-``, ``sum``,
``dot``, ``apply_powers``, ``powers``, ``of_Fr`` — is executed by the CUDA library
through the C ABI (``zk_g*_msm`` / ``zk_g*_fixed_base_mul``).  ``GT`` / ``Pairing`` (the verifier
side, curve.ml:212-220) go through ``zk_pairing_product`` / ``zk_gt_mul`` and
``of_compressed_bytes_exn`` through ``zk_g*_decompress``.  ``README.md:36-40`` of
the reference spells the module ``Ecp``; ``Ecp = Curve`` aliases are provided in
``zukelang_b200/__init__``-level imports for both spellings.
"""

from __future__ import annotations

import ctypes
import random
from typing import Callable, Dict, Iterable, List, Sequence, Tuple

from . import _lib

Var = Tuple[str, int]

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_P_HALF = (0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB - 1) // 2


class Fr:
    """curve.ml:121-155.  Elements are ints in [0, r)."""
    order = R
    zero = 0
    one = 1

    @staticmethod
    def of_int(i: int) -> int:
        return i % R

    of_z = of_int

    @staticmethod
    def add(a, b): return (a + b) % R
    @staticmethod
    def sub(a, b): return (a - b) % R
    @staticmethod
    def mul(a, b): return (a * b) % R
    @staticmethod
    def neg(a): return (-a) % R

    @staticmethod
    def div(a, b):
        if b % R == 0:
            raise ZeroDivisionError("Fr division by zero")
        return a * pow(b, -1, R) % R

    @staticmethod
    def pow(a, e: int): return pow(a, e, R)

    @staticmethod
    def gen(rng: random.Random) -> int:
        """curve.ml:136 ``Fr.random ~state:rng ()``."""
        return rng.randrange(R)

    @staticmethod
    def to_bytes(a: int) -> bytes:
        return (a % R).to_bytes(32, "little")

    @staticmethod
    def of_bytes(b: bytes) -> int:
        v = int.from_bytes(b, "little")
        if v >= R:
            raise ValueError("Fr.of_bytes: not canonical")
        return v


def fr_vector(ks: Iterable[int]) -> bytes:
    return b"".join((k % R).to_bytes(32, "little") for k in ks)


class Point:
    """An opaque group element: uncompressed wire bytes (+ lazily the compressed form)."""
    __slots__ = ("raw", "_comp")

    def __init__(self, raw: bytes, comp: bytes | None = None):
        self.raw = bytes(raw)
        self._comp = comp

    def __eq__(self, other):
        return isinstance(other, Point) and self.raw == other.raw

    def __hash__(self):
        return hash(self.raw)

    def __repr__(self):
        return "Point(%s…)" % self.raw[:8].hex()

    def is_zero(self) -> bool:
        return self.raw[0] == 0x40


class _Group:
    """The ``G`` signature of curve.ml:22-50 for one group, backed by the CUDA library."""

    def __init__(self, name: str, raw: int, comp: int, gen_raw_hex: str):
        self.name, self.RAW, self.COMP, self.OUT = name, raw, comp, raw + comp
        self._msm = "zk_%s_msm" % name.lower()
        self._fixed = "zk_%s_fixed_base_mul" % name.lower()
        self.zero = Point(bytes([0x40]) + bytes(raw - 1))
        self.one = Point(bytes.fromhex(gen_raw_hex))

    # ---- the MSM primitive everything else reduces to ------------------------------
    def msm(self, points: Sequence[Point], scalars: Sequence[int]) -> Point:
        """sum_i scalars[i] * points[i]   (curve.ml:91-103 as one call)."""
        if len(points) != len(scalars):
            raise _lib.InvalidArgument(_lib.ZK_EARG, "msm: length mismatch")
        out = (ctypes.c_uint8 * self.OUT)()
        n = len(points)
        if n == 0:
            return self.zero
        bases = b"".join(p.raw for p in points)
        _lib.check(getattr(_lib.lib(), self._msm)(bases, None, fr_vector(scalars), n, out))
        b = bytes(out)
        return Point(b[:self.RAW], b[self.RAW:])

    # ---- curve.ml:159-191 ExtendG -----------------------------------------------------
    def add(self, a: Point, b: Point) -> Point:
        return self.msm([a, b], [1, 1])

    def mul(self, a: Point, k: int) -> Point:
        return self.msm([a], [k % R])

    def neg(self, a: Point) -> Point:
        return self.msm([a], [R - 1])

    def sub(self, a: Point, b: Point) -> Point:
        return self.msm([a, b], [1, R - 1])

    def eq(self, a: Point, b: Point) -> bool:
        return a.raw == b.raw

    def sum(self, pts: Sequence[Point]) -> Point:
        return self.msm(list(pts), [1] * len(pts))

    def of_Fr(self, k: int) -> Point:                      # curve.ml:180
        return self.fixed_base([k])[0]

    def fixed_base(self, ks: Sequence[int]) -> List[Point]:
        n = len(ks)
        if n == 0:
            return []
        out = (ctypes.c_uint8 * (self.RAW * n))()
        _lib.check(getattr(_lib.lib(), self._fixed)(fr_vector(ks), n, out))
        b = bytes(out)
        return [Point(b[i * self.RAW:(i + 1) * self.RAW]) for i in range(n)]

    # ---- curve.ml:79-119 ExtendMap ------------------------------------------------------
    def sum_map(self, m: Dict[Var, object], f: Callable[[Var, object], Point]) -> Point:
        """curve.ml:91."""
        return self.sum([f(k, m[k]) for k in sorted(m)])

    def dot(self, m: Dict[Var, Point], c: Dict[Var, int]) -> Point:
        """curve.ml:94-103; a domain mismatch is ``assert false`` in the reference."""
        if set(m) != set(c):
            raise AssertionError("Domain mismatch")
        keys = sorted(m)
        return self.msm([m[k] for k in keys], [c[k] for k in keys])

    def powers(self, d: int, s: int) -> List[Point]:
        """curve.ml:106-109 — d + 1 points g^(s^i)."""
        ks, cur = [], 1
        for _ in range(d + 1):
            ks.append(cur)
            cur = cur * s % R
        return self.fixed_base(ks)

    def apply_powers(self, cs: Sequence[int], xis: Sequence[Point]) -> Point:
        """curve.ml:112-118 — Invalid_argument "apply_powers" when cs is longer than xis."""
        if len(cs) > len(xis):
            raise _lib.InvalidArgument(_lib.ZK_EARG, "apply_powers")
        return self.msm(list(xis[:len(cs)]), list(cs))

    def to_compressed_bytes(self, a: Point) -> bytes:
        """curve.ml:199 / :208.  Results of device calls already carry the compressed form; for
        the others it is a re-flagging of the uncompressed bytes (x with the three flag bits, the
        sign bit by comparing the canonical y with (p-1)/2 — c1 first, then c0, in G2): a byte
        format conversion, no field arithmetic."""
        if a._comp is None:
            half = self.RAW // 2
            if a.raw[0] & 0x40:
                a._comp = bytes([0xC0]) + bytes(half - 1)
            else:
                y = [int.from_bytes(a.raw[half + o:half + o + 48], "big") for o in range(0, half, 48)]
                lead = next((c for c in y if c), 0)                       # G2 wire order is c1 | c0
                a._comp = bytes([a.raw[0] | 0x80 | (0x20 if lead > _P_HALF else 0)]) + a.raw[1:half]
        return a._comp

    def to_bytes(self, a: Point) -> bytes:
        return a.raw

    def of_compressed_bytes_many(self, blobs: Sequence[bytes]) -> List[Point]:
        """``of_compressed_bytes_exn`` (curve.ml:201 / :210) over a batch: one device call."""
        n = len(blobs)
        if n == 0:
            return []
        if any(len(b) != self.COMP for b in blobs):
            raise _lib.InvalidArgument(_lib.ZK_EARG, "of_compressed_bytes_exn: wrong length")
        out = (ctypes.c_uint8 * (self.RAW * n))()
        _lib.check(getattr(_lib.lib(), "zk_%s_decompress" % self.name.lower())(b"".join(blobs), n, out))
        raw = bytes(out)
        return [Point(raw[i * self.RAW:(i + 1) * self.RAW], bytes(blobs[i])) for i in range(n)]

    def of_compressed_bytes_exn(self, b: bytes) -> Point:
        return self.of_compressed_bytes_many([b])[0]


_G1_GEN = ("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
           "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1")
_G2_GEN = ("13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
           "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8"
           "0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be"
           "0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801")


class GTElem:
    """An opaque GT element (576 bytes in the library's own encoding, include/zkb200.h)."""
    __slots__ = ("raw",)

    def __init__(self, raw: bytes):
        if len(raw) != _lib.GT_BYTES:
            raise _lib.InvalidArgument(_lib.ZK_EARG, "GT.of_bytes_exn: wrong length")
        self.raw = bytes(raw)

    def __eq__(self, other):
        return isinstance(other, GTElem) and self.raw == other.raw

    def __hash__(self):
        return hash(self.raw)

    def __repr__(self):
        return "GT(%s…)" % self.raw[40:48].hex()


class _GT:
    """``GT`` with ExtendG's additive notation (curve.ml:212-220): ``+`` is the GT product."""
    zero = GTElem((1).to_bytes(48, "big") + bytes(_lib.GT_BYTES - 48))     # the neutral element

    @staticmethod
    def add(a: GTElem, b: GTElem) -> GTElem:
        out = (ctypes.c_uint8 * _lib.GT_BYTES)()
        _lib.check(_lib.lib().zk_gt_mul(a.raw, b.raw, out))
        return GTElem(bytes(out))

    @staticmethod
    def eq(a: GTElem, b: GTElem) -> bool:
        return a.raw == b.raw

    @staticmethod
    def to_bytes(a: GTElem) -> bytes:
        return a.raw

    @staticmethod
    def of_bytes_exn(b: bytes) -> GTElem:
        return GTElem(b)


class _Pairing:
    """``Bls12_381.Pairing`` as used at groth16.ml:103,168 and pinocchio.ml:269."""

    @staticmethod
    def product(pairs: Sequence[Tuple[Point, Point]], negate: Sequence[bool] | None = None) -> GTElem:
        """prod_i e(+-p_i, q_i) with one final exponentiation: a GT sum / difference of pairings
        (``e a b + e c d - e f g`` in the reference's additive notation) as ONE device call."""
        n = len(pairs)
        if n == 0:
            return _GT.zero
        neg = bytes(1 if x else 0 for x in negate) if negate is not None else None
        if neg is not None and len(neg) != n:
            raise _lib.InvalidArgument(_lib.ZK_EARG, "pairing: length mismatch")
        out = (ctypes.c_uint8 * _lib.GT_BYTES)()
        _lib.check(_lib.lib().zk_pairing_product(b"".join(p.raw for p, _ in pairs), b"".join(q.raw for _, q in pairs),
                                                 neg, n, out))
        return GTElem(bytes(out))

    @staticmethod
    def products(groups: Sequence[Tuple[Sequence[Tuple[Point, Point]], Sequence[bool] | None]]) -> List[GTElem]:
        """Several independent pairing products in ONE device call (their Miller loops and final
        exponentiations run side by side): ``groups`` is a list of ``(pairs, negate)``."""
        k = len(groups)
        if k == 0:
            return []
        g1, g2, neg, counts = [], [], bytearray(), []
        for pairs, negate in groups:
            if negate is not None and len(negate) != len(pairs):
                raise _lib.InvalidArgument(_lib.ZK_EARG, "pairing: length mismatch")
            counts.append(len(pairs))
            g1 += [p.raw for p, _ in pairs]
            g2 += [q.raw for _, q in pairs]
            neg += bytes(1 if x else 0 for x in (negate if negate is not None else [False] * len(pairs)))
        if not g1:
            return [_GT.zero] * k
        out = (ctypes.c_uint8 * (_lib.GT_BYTES * k))()
        _lib.check(_lib.lib().zk_pairing_product_batch(b"".join(g1), b"".join(g2), bytes(neg),
                                                       (ctypes.c_uint32 * k)(*counts), k, out))
        b = bytes(out)
        return [GTElem(b[i * _lib.GT_BYTES:(i + 1) * _lib.GT_BYTES]) for i in range(k)]

    @staticmethod
    def pairing(p: Point, q: Point) -> GTElem:
        return _Pairing.product([(p, q)])


class Bls12_381:
    """``Curve.Bls12_381`` (curve.mli:46-60): Fr, G1, G2, GT and Pairing."""
    Fr = Fr
    G1 = _Group("G1", 96, 48, _G1_GEN)
    G2 = _Group("G2", 192, 96, _G2_GEN)
    GT = _GT
    Pairing = _Pairing
