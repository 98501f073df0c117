"""CPU oracle — BLS12-381 field / curve arithmetic in Python big integers.

TEST INFRASTRUCTURE ONLY.  Nothing under ``zukelang_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` leg use it, and only as the checker.

PARITY UNPINNED.  The reference (camlspotter/zukelang) performs all of this
arithmetic inside the un-vendored opam package ``bls12-381 = 6.1.0``
(``/root/reference/dune-project:23``; pulled in by ``include Bls12_381`` at
``src/lib/zk/curve.ml:77``), an OCaml binding over blst.  Neither it nor an
OCaml toolchain exists in this container, and the reference's tests hold no
byte-level vectors (``src/lib/test/test.ml:121`` uses a self-seeded RNG).  This
file therefore restates the *published* BLS12-381 definition and is pinned by
public constants instead (see ``self_check``):

* p, r primality and the BLS parametrisation r = z^4 - z^2 + 1,
  p = (z-1)^2 r / 3 + z with z = -0xd201000000010000;
* the standard generators lie on y^2 = x^3 + 4 and y^2 = x^3 + 4(1+u) and are
  killed by r;
* the zcash compressed encodings of both generators (the byte form used by the
  reference's ``to_compressed_bytes``, ``curve.ml:199,208``);
* the 2^32 root of unity 5^((r-1)/2^32) recorded by the reference at
  ``src/lib/zk/FFT.ml:179-219``;
* pairing bilinearity (``oracle/pairing.py``) and the verifier equations of
  ``groth16.ml:163-173`` / ``pinocchio.ml:254-420`` (``oracle/zk.py``).

Reference call sites restated here:
  G.add / G.mul / G.negate / G.eq / zero / one   curve.ml:159-171 (ExtendG)
  of_Fr = ( * ) one                              curve.ml:180
  to_compressed_bytes / of_compressed_bytes_exn  curve.ml:199-201, 208-210
"""

from __future__ import annotations

# --------------------------------------------------------------------------
# constants
# --------------------------------------------------------------------------
Z_PARAM = -0xD201000000010000
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

G1_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G2_X0 = 0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8
G2_X1 = 0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E
G2_Y0 = 0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801
G2_Y1 = 0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE

FR_TWO_ADICITY = 32
# FFT.ml:208 picks g = 5, FFT.ml:219 sets omega = g^((r-1)/2^32)
FR_ROOT_2_32 = pow(5, (R - 1) >> 32, R)

B1 = 4            # G1: y^2 = x^3 + 4
B2 = (4, 4)       # G2: y^2 = x^3 + 4(1+u)


# --------------------------------------------------------------------------
# Fr helpers (scalars are plain ints mod R)
# --------------------------------------------------------------------------
def fr(x: int) -> int:
    return x % R


def fr_inv(x: int) -> int:
    x %= R
    if x == 0:
        raise ZeroDivisionError("Fr inverse of zero")
    return pow(x, -1, R)


def fr_to_bytes(x: int) -> bytes:
    """32-byte little-endian canonical scalar (the C-ABI scalar format)."""
    return (x % R).to_bytes(32, "little")


def fr_from_bytes(b: bytes) -> int:
    v = int.from_bytes(b, "little")
    if v >= R:
        raise ValueError("non-canonical Fr")
    return v


# --------------------------------------------------------------------------
# Fp2 = Fp[u]/(u^2+1), elements are (c0, c1)
# --------------------------------------------------------------------------
def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_sqr(a):
    return f2_mul(a, a)


def f2_inv(a):
    n = (a[0] * a[0] + a[1] * a[1]) % P
    if n == 0:
        raise ZeroDivisionError("Fp2 inverse of zero")
    ni = pow(n, -1, P)
    return (a[0] * ni % P, (-a[1]) * ni % P)


def f2_pow(a, e):
    out = (1, 0)
    base = a
    while e:
        if e & 1:
            out = f2_mul(out, base)
        base = f2_sqr(base)
        e >>= 1
    return out


def f2_sqrt(a):
    """Square root in Fp2 or None.  p = 3 mod 4 (Adj–Rodríguez-Henríquez alg. 9)."""
    if a == (0, 0):
        return (0, 0)
    a1 = f2_pow(a, (P - 3) // 4)
    alpha = f2_mul(f2_sqr(a1), a)
    x0 = f2_mul(a1, a)
    if alpha == (P - 1, 0):
        cand = f2_mul((0, 1), x0)
    else:
        b = f2_pow(f2_add((1, 0), alpha), (P - 1) // 2)
        cand = f2_mul(b, x0)
    return cand if f2_sqr(cand) == (a[0] % P, a[1] % P) else None


# --------------------------------------------------------------------------
# Generic short-Weierstrass group (a = 0) over a field given by an ops table.
# Points are None (identity) or affine tuples (x, y).
# --------------------------------------------------------------------------
class _Field:
    def __init__(self, add, sub, mul, inv, neg, zero, one, b):
        self.add, self.sub, self.mul, self.inv, self.neg = add, sub, mul, inv, neg
        self.zero, self.one, self.b = zero, one, b


_FP = _Field(
    add=lambda a, b: (a + b) % P,
    sub=lambda a, b: (a - b) % P,
    mul=lambda a, b: (a * b) % P,
    inv=lambda a: pow(a, -1, P),
    neg=lambda a: (-a) % P,
    zero=0,
    one=1,
    b=B1,
)
_FP2 = _Field(f2_add, f2_sub, f2_mul, f2_inv, f2_neg, (0, 0), (1, 0), B2)


class Group:
    """Restates the G signature of curve.ml:22-50 for one group.

    ``add``/``mul``/``neg``/``eq`` stand for the bls12-381 package functions
    wrapped by ExtendG (curve.ml:159-191).
    """

    def __init__(self, name, field, gen):
        self.name, self.F, self.one = name, field, gen
        self.zero = None

    # -- curve.ml:160-171 ----------------------------------------------------
    def is_on_curve(self, pt):
        if pt is None:
            return True
        F = self.F
        x, y = pt
        return F.mul(y, y) == F.add(F.mul(F.mul(x, x), x), F.b)

    def neg(self, pt):
        if pt is None:
            return None
        return (pt[0], self.F.neg(pt[1]))

    def add(self, p1, p2):
        F = self.F
        if p1 is None:
            return p2
        if p2 is None:
            return p1
        x1, y1 = p1
        x2, y2 = p2
        if x1 == x2:
            if y1 == y2:
                if y1 == F.zero:
                    return None
                three_x2 = F.mul(F.add(F.add(x1, x1), x1), x1)
                lam = F.mul(three_x2, F.inv(F.add(y1, y1)))
            else:
                return None
        else:
            lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    def sub(self, p1, p2):
        return self.add(p1, self.neg(p2))

    def double(self, pt):
        return self.add(pt, pt)

    def mul(self, pt, k: int):
        """pt * k with k an Fr element (reduced mod r like Fr.t)."""
        k %= R
        acc = None
        addend = pt
        while k:
            if k & 1:
                acc = self.add(acc, addend)
            addend = self.add(addend, addend)
            k >>= 1
        return acc

    def eq(self, p1, p2):
        return p1 == p2

    def of_Fr(self, k: int):          # curve.ml:180  of_Fr = ( * ) one
        return self.mul(self.one, k)

    def sum(self, pts):               # curve.ml:179  List.fold_left (+) zero
        acc = None
        for q in pts:
            acc = self.add(acc, q)
        return acc

    def in_subgroup(self, pt):
        return self._mul_raw(pt, R) is None

    def _mul_raw(self, pt, k):
        acc = None
        addend = pt
        while k:
            if k & 1:
                acc = self.add(acc, addend)
            addend = self.add(addend, addend)
            k >>= 1
        return acc


G1 = Group("G1", _FP, (G1_X, G1_Y))
G2 = Group("G2", _FP2, ((G2_X0, G2_X1), (G2_Y0, G2_Y1)))


# --------------------------------------------------------------------------
# zcash / blst serialisation (the byte form bit-identity is judged in)
# --------------------------------------------------------------------------
def _fp_be(x: int) -> bytes:
    return x.to_bytes(48, "big")


def g1_to_uncompressed(pt) -> bytes:
    """96 B = x || y big-endian; identity = 0x40 then zeros."""
    if pt is None:
        return bytes([0x40]) + bytes(95)
    return _fp_be(pt[0]) + _fp_be(pt[1])


def g1_from_uncompressed(b: bytes):
    if len(b) != 96:
        raise ValueError("G1 uncompressed length")
    if b[0] & 0x40:
        if any(b[1:]) or b[0] != 0x40:
            raise ValueError("bad G1 infinity encoding")
        return None
    if b[0] & 0xE0:
        raise ValueError("bad G1 flags")
    x = int.from_bytes(b[:48], "big")
    y = int.from_bytes(b[48:], "big")
    if x >= P or y >= P:
        raise ValueError("G1 coordinate not canonical")
    pt = (x, y)
    if not G1.is_on_curve(pt):
        raise ValueError("G1 point not on curve")
    return pt


def g1_compress(pt) -> bytes:
    """48 B big-endian x; bit7 compressed, bit6 infinity, bit5 y > (p-1)/2."""
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    out = bytearray(_fp_be(x))
    out[0] |= 0x80
    if y > (P - 1) // 2:
        out[0] |= 0x20
    return bytes(out)


def g1_decompress(b: bytes):
    if len(b) != 48 or not (b[0] & 0x80):
        raise ValueError("G1 compressed encoding")
    if b[0] & 0x40:
        if (b[0] & 0x3F) or any(b[1:]):
            raise ValueError("bad G1 infinity encoding")
        return None
    sign = bool(b[0] & 0x20)
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if x >= P:
        raise ValueError("G1 x not canonical")
    y2 = (x * x * x + B1) % P
    y = pow(y2, (P + 1) // 4, P)
    if y * y % P != y2:
        raise ValueError("G1 x not on curve")
    if (y > (P - 1) // 2) != sign:
        y = P - y
    return (x, y)


def _f2_gt_half(y) -> bool:
    """Lexicographic 'largest' rule for Fp2: compare c1 first, then c0."""
    if y[1] != 0:
        return y[1] > (P - 1) // 2
    return y[0] > (P - 1) // 2


def g2_to_uncompressed(pt) -> bytes:
    """192 B = x.c1 || x.c0 || y.c1 || y.c0, each 48 B big-endian."""
    if pt is None:
        return bytes([0x40]) + bytes(191)
    (x0, x1), (y0, y1) = pt
    return _fp_be(x1) + _fp_be(x0) + _fp_be(y1) + _fp_be(y0)


def g2_from_uncompressed(b: bytes):
    if len(b) != 192:
        raise ValueError("G2 uncompressed length")
    if b[0] & 0x40:
        if any(b[1:]) or b[0] != 0x40:
            raise ValueError("bad G2 infinity encoding")
        return None
    if b[0] & 0xE0:
        raise ValueError("bad G2 flags")
    x1, x0, y1, y0 = (int.from_bytes(b[i * 48:(i + 1) * 48], "big") for i in range(4))
    if max(x0, x1, y0, y1) >= P:
        raise ValueError("G2 coordinate not canonical")
    pt = ((x0, x1), (y0, y1))
    if not G2.is_on_curve(pt):
        raise ValueError("G2 point not on curve")
    return pt


def g2_compress(pt) -> bytes:
    if pt is None:
        return bytes([0xC0]) + bytes(95)
    (x0, x1), y = pt
    out = bytearray(_fp_be(x1) + _fp_be(x0))
    out[0] |= 0x80
    if _f2_gt_half(y):
        out[0] |= 0x20
    return bytes(out)


def g2_decompress(b: bytes):
    if len(b) != 96 or not (b[0] & 0x80):
        raise ValueError("G2 compressed encoding")
    if b[0] & 0x40:
        if (b[0] & 0x3F) or any(b[1:]):
            raise ValueError("bad G2 infinity encoding")
        return None
    sign = bool(b[0] & 0x20)
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    if x0 >= P or x1 >= P:
        raise ValueError("G2 x not canonical")
    x = (x0, x1)
    y2 = f2_add(f2_mul(f2_sqr(x), x), B2)
    y = f2_sqrt(y2)
    if y is None:
        raise ValueError("G2 x not on curve")
    if _f2_gt_half(y) != sign:
        y = f2_neg(y)
    return (x, y)


# convenience: group → (to_uncompressed, compress) tables used by tests
SER = {
    "G1": (g1_to_uncompressed, g1_from_uncompressed, g1_compress, g1_decompress),
    "G2": (g2_to_uncompressed, g2_from_uncompressed, g2_compress, g2_decompress),
}

G1_GEN_COMPRESSED = bytes.fromhex(
    "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
    "6c55e83ff97a1aeffb3af00adb22c6bb"
)
G2_GEN_COMPRESSED = bytes.fromhex(
    "93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049"
    "334cf11213945d57e5ac7d055d042b7e024aa2b2f08f0a91260805272dc51051"
    "c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8"
)


def _is_prime(n: int) -> bool:
    """Deterministic-enough Miller–Rabin (40 fixed bases)."""
    if n < 2:
        return False
    small = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71,
             73, 79, 83, 89, 97, 101, 103, 107, 109, 113, 127, 131, 137, 139, 149, 151,
             157, 163, 167, 173]
    for q in small:
        if n % q == 0:
            return n == q
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in small:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def self_check() -> None:
    """Pins the restatement against public facts (SURVEY.md §8c (i),(ii),(vi))."""
    z = Z_PARAM
    assert _is_prime(P) and _is_prime(R)
    assert R == z ** 4 - z ** 2 + 1
    assert P == (z - 1) ** 2 * R // 3 + z
    assert P.bit_length() == 381 and R.bit_length() == 255
    assert G1.is_on_curve(G1.one) and G2.is_on_curve(G2.one)
    assert G1.in_subgroup(G1.one) and G2.in_subgroup(G2.one)
    assert g1_compress(G1.one) == G1_GEN_COMPRESSED
    assert g2_compress(G2.one) == G2_GEN_COMPRESSED
    assert g1_decompress(G1_GEN_COMPRESSED) == G1.one
    assert g2_decompress(G2_GEN_COMPRESSED) == G2.one
    assert (R - 1) % (1 << 32) == 0 and ((R - 1) >> 32) % 2 == 1
    w = FR_ROOT_2_32
    assert pow(w, 1 << 32, R) == 1 and pow(w, 1 << 31, R) == R - 1
    assert w == 0x212D79E5B416B6F0FD56DC8D168D6C0C4024FF270B3E0941B788F500B912F1F
    # curve.ml:224-239: g^(ab+cd) = g^(ab) + g^(cd)
    a, b, c, d = 1234, 5678, 4321, 8765
    assert G1.of_Fr(a * b + c * d) == G1.add(G1.of_Fr(a * b), G1.of_Fr(c * d))
    assert G1.mul(G1.one, a) == G1.of_Fr(a)


if __name__ == "__main__":
    self_check()
    print("oracle/bls12_381.py self-check OK")
