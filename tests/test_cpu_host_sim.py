"""Host simulation of the device limb algorithms: the SAME headers the kernels use
(mont.cuh, ec.cuh) compiled by g++ with -DZK_HOST_SIM, where ptx.cuh emulates each PTX carry
instruction.  A development check of the formulas on the CPU — tests only, never shipped."""
import ctypes
import os
import random
import subprocess
import tempfile

import pytest

from oracle.bls12_381 import G1, G2, P, R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "host_sim")


@pytest.fixture(scope="module")
def sim():
    out = tempfile.mkdtemp(prefix="zk_host_sim_")
    libs = {}
    for name in ("sim_field", "sim_ec"):
        so = os.path.join(out, name + ".so")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-DZK_HOST_SIM", "-shared", "-fPIC", "-x", "c++",
                               os.path.join(SIM, name + ".cpp"), "-o", so])
        libs[name] = ctypes.CDLL(so)
    return libs


def _field(fn, n, op, a, b=0):
    A = (ctypes.c_uint32 * n)(*[(a >> (32 * i)) & 0xFFFFFFFF for i in range(n)])
    B = (ctypes.c_uint32 * n)(*[(b >> (32 * i)) & 0xFFFFFFFF for i in range(n)])
    Out = (ctypes.c_uint32 * n)()
    fn(op, A, B, Out)
    return sum(int(Out[i]) << (32 * i) for i in range(n))


@pytest.mark.parametrize("which,mod,n", [("sim_fp_op", P, 12), ("sim_fr_op", R, 8)])
def test_montgomery_limb_algorithms(sim, which, mod, n):
    fn = getattr(sim["sim_field"], which)
    rng = random.Random(5)
    Rm = 1 << (32 * n)
    Ri = pow(Rm, -1, mod)
    vals = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, Rm % mod, (1 << (32 * n - 3)) % mod, mod - (1 << 200)] \
        + [rng.randrange(mod) for _ in range(400)]
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        assert _field(fn, n, 0, a, b) == a * b * Ri % mod
        assert _field(fn, n, 1, a, b) == (a + b) % mod
        assert _field(fn, n, 2, a, b) == (a - b) % mod
        assert _field(fn, n, 3, a) == (-a) % mod
        assert _field(fn, n, 4, a) == a * Rm % mod
        assert _field(fn, n, 5, a) == a * Ri % mod
        assert _field(fn, n, 9, a) == a * a * Ri % mod               # dedicated squaring
        both = (a * b + b * b) * Ri % mod
        assert _field(fn, n, 10, a, b) == both                       # two products, rows interleaved
        assert _field(fn, n, 11, a, b) == both                       # ... with the row loop rolled
        assert _field(fn, n, 12, a, b) == both
    for a in vals[:40]:
        exp = (pow(a, -1, mod) * Rm % mod) if a else 0
        assert _field(fn, n, 6, a * Rm % mod) == exp                 # binary extended Euclid
        if a in vals[:10]:
            assert _field(fn, n, 8, a * Rm % mod) == exp             # Fermat cross-check


RM = (1 << 384) % P


def _fpl(x):
    x = x * RM % P
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(12)]


def _unfp(l):
    return sum(int(v) << (32 * i) for i, v in enumerate(l)) * pow(RM, -1, P) % P


def _enc_f(G, c):
    return _fpl(c) if G is G1 else _fpl(c[0]) + _fpl(c[1])


def _dec_f(G, l):
    return _unfp(l) if G is G1 else (_unfp(l[:12]), _unfp(l[12:]))


def _enc_aff(G, pt, w):
    return [0] * (2 * w) if pt is None else _enc_f(G, pt[0]) + _enc_f(G, pt[1])


def _enc_xyzz(G, pt, z, w):
    if pt is None:
        return [0] * (4 * w)
    F = G.F
    zf = z if G is G1 else (z, (z * 7 + 1) % P)
    zz = F.mul(zf, zf)
    zzz = F.mul(zz, zf)
    return _enc_f(G, F.mul(pt[0], zz)) + _enc_f(G, F.mul(pt[1], zzz)) + _enc_f(G, zz) + _enc_f(G, zzz)


def _ec(sim, G, op, a, b, k=0):
    w = 12 if G is G1 else 24
    fn = sim["sim_ec"].sim_g1_op if G is G1 else sim["sim_ec"].sim_g2_op
    A = (ctypes.c_uint32 * (4 * w))(*a)
    B = (ctypes.c_uint32 * (4 * w))(*(b + [0] * (4 * w - len(b))))
    K = (ctypes.c_uint32 * 8)(*[(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
    Out = (ctypes.c_uint32 * (6 * w))()
    fn(op, A, B, K, Out)
    o = list(Out)
    x, y = _dec_f(G, o[:w]), _dec_f(G, o[w:2 * w])
    zero = 0 if G is G1 else (0, 0)
    return None if (x == zero and y == zero) else (x, y)


@pytest.mark.parametrize("G", [G1, G2], ids=["G1", "G2"])
def test_xyzz_formulas(sim, G):
    w = 12 if G is G1 else 24
    rng = random.Random(9)
    pts = [None] + [G.mul(G.one, rng.randrange(1, R)) for _ in range(4)]
    for p in pts:
        for q in pts:
            z = rng.randrange(1, P)
            assert _ec(sim, G, 0, _enc_xyzz(G, p, z, w), _enc_aff(G, q, w)) == G.add(p, q)
            assert _ec(sim, G, 5, _enc_xyzz(G, p, z, w), _enc_aff(G, q, w)) == G.add(p, q)   # paired products
            assert _ec(sim, G, 1, _enc_xyzz(G, p, z, w), _enc_xyzz(G, q, z + 3, w)) == G.add(p, q)
        z = rng.randrange(1, P)
        assert _ec(sim, G, 0, _enc_xyzz(G, p, z, w), _enc_aff(G, p, w)) == G.add(p, p)          # P + P
        assert _ec(sim, G, 5, _enc_xyzz(G, p, z, w), _enc_aff(G, p, w)) == G.add(p, p)
        assert _ec(sim, G, 5, _enc_xyzz(G, p, z, w), _enc_aff(G, G.neg(p), w)) is None
        assert _ec(sim, G, 0, _enc_xyzz(G, p, z, w), _enc_aff(G, G.neg(p), w)) is None           # P - P
        assert _ec(sim, G, 1, _enc_xyzz(G, p, z, w), _enc_xyzz(G, G.neg(p), z + 5, w)) is None
        assert _ec(sim, G, 2, _enc_xyzz(G, p, z, w), []) == G.add(p, p)
    for k in (0, 1, R - 1, rng.randrange(R)):
        assert _ec(sim, G, 3, _enc_xyzz(G, pts[1], 12345, w), [], k) == G.mul(pts[1], k)


# ---- Fp12 tower / pairing / square roots (pairing.cuh) ------------------------------------------
@pytest.fixture(scope="module")
def simp():
    out = tempfile.mkdtemp(prefix="zk_host_sim_")
    so = os.path.join(out, "sim_pairing.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DZK_HOST_SIM", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(SIM, "sim_pairing.cpp"), "-o", so])
    return ctypes.CDLL(so)


def _tower_to_limbs(t):
    """oracle Fp12 (coefficients of w^0..w^11, w^12 = 2 w^6 - 2) -> device tower limbs.
    Slot a_j of c_i is the Fp2 coefficient x + y u of w^(2j+i); u = w^6 - 1."""
    out = []
    for i in range(2):
        for j in range(3):
            k = 2 * j + i
            y = t[k + 6]
            x = (t[k] + y) % P
            out += _fpl(x) + _fpl(y)
    return out


def _limbs_to_tower(l):
    t = [0] * 12
    o = 0
    for i in range(2):
        for j in range(3):
            k = 2 * j + i
            x, y = _unfp(l[o:o + 12]), _unfp(l[o + 12:o + 24])
            o += 24
            t[k] = (x - y) % P
            t[k + 6] = y
    return tuple(t)


def _f12(simp, op, a, b=None):
    A = (ctypes.c_uint32 * 144)(*_tower_to_limbs(a))
    B = (ctypes.c_uint32 * 144)(*_tower_to_limbs(b if b is not None else a))
    Out = (ctypes.c_uint32 * 144)()
    simp.sim_f12_op(op, A, B, Out)
    return _limbs_to_tower(list(Out))


def test_fp12_tower_matches_the_oracle_field(simp):
    from oracle import pairing as OP
    rng = random.Random(21)
    a = tuple(rng.randrange(P) for _ in range(12))
    b = tuple(rng.randrange(P) for _ in range(12))
    assert _f12(simp, 0, a, b) == OP.f12_mul(a, b)
    assert _f12(simp, 1, a) == OP.f12_mul(a, a)
    assert OP.f12_mul(_f12(simp, 2, a), a) == OP.F12_ONE
    assert _f12(simp, 3, a) == OP.f12_pow(a, P)                       # Frobenius
    assert _f12(simp, 4, a) == OP.f12_pow(a, P ** 6)                  # conjugation
    u = OP.f12_pow(a, (P ** 6 - 1) * (P ** 2 + 1))                    # a unitary element
    assert OP.f12_mul(_f12(simp, 5, u), OP.f12_pow(u, OP.ATE_LOOP)) == OP.F12_ONE   # u^z, z < 0


def test_miller_loop_and_final_exponentiation_match_the_oracle(simp):
    from oracle import pairing as OP
    rng = random.Random(22)
    p = G1.mul(G1.one, rng.randrange(1, R))
    q = G2.mul(G2.one, rng.randrange(1, R))
    Pl = (ctypes.c_uint32 * 24)(*_enc_aff(G1, p, 12))
    Ql = (ctypes.c_uint32 * 48)(*_enc_aff(G2, q, 24))
    F = (ctypes.c_uint32 * 144)()
    simp.sim_miller(Pl, Ql, F)
    f = _limbs_to_tower(list(F))
    assert f == OP.miller_loop(p, q)                                  # same lines, same loop
    E = (ctypes.c_uint32 * 144)()
    simp.sim_final_exp(F, E)
    e = _limbs_to_tower(list(E))
    assert e == OP.f12_pow(OP.pairing(p, q), 3)                       # hard part carries the factor 3
    assert e != OP.F12_ONE and OP.f12_pow(e, R) == OP.F12_ONE


def test_square_roots_and_subgroup_check(simp):
    rng = random.Random(23)
    for _ in range(6):
        a = rng.randrange(P)
        sq = a * a % P
        Out = (ctypes.c_uint32 * 12)()
        assert simp.sim_fp_sqrt((ctypes.c_uint32 * 12)(*_fpl(sq)), Out) == 1
        assert _unfp(list(Out)) in (a, P - a)
        nonsq = sq * (P - 1) % P                                      # -1 is a non-residue (p = 3 mod 4)
        assert simp.sim_fp_sqrt((ctypes.c_uint32 * 12)(*_fpl(nonsq)), Out) == (1 if sq == 0 else 0)
    F2 = G2.F
    n_roots = n_none = 0
    cases = [(rng.randrange(P), rng.randrange(P)) for _ in range(8)]
    cases += [(rng.randrange(P), 0), (0, rng.randrange(P)), (0, 0), (P - 4, 0), (P - 1, 0)]
    for c in cases:
        for a in (F2.mul(c, c), c):
            Out = (ctypes.c_uint32 * 24)()
            ok = simp.sim_fp2_sqrt((ctypes.c_uint32 * 24)(*(_fpl(a[0]) + _fpl(a[1]))), Out)
            if ok:
                y = (_unfp(list(Out)[:12]), _unfp(list(Out)[12:]))
                assert F2.mul(y, y) == (a[0] % P, a[1] % P)
                n_roots += 1
            else:
                # a has no square root: its norm is a non-residue in Fp
                nrm = (a[0] * a[0] + a[1] * a[1]) % P
                assert pow(nrm, (P - 1) // 2, P) == P - 1
                n_none += 1
    assert n_roots >= len(cases) and n_none >= 1
    # subgroup membership: generator multiples pass, a curve point outside the subgroup fails
    assert simp.sim_in_subgroup(0, (ctypes.c_uint32 * 24)(*_enc_aff(G1, G1.mul(G1.one, 77), 12))) == 1
    assert simp.sim_in_subgroup(1, (ctypes.c_uint32 * 48)(*_enc_aff(G2, G2.mul(G2.one, 77), 24))) == 1
    x = 1
    while True:                                                       # first x with x^3 + 4 a square
        y2 = (x ** 3 + 4) % P
        y = pow(y2, (P + 1) // 4, P)
        if y * y % P == y2 and G1.add(G1.mul((x, y), R - 1), (x, y)) is not None:   # [r]P != O
            break
        x += 1
    assert simp.sim_in_subgroup(0, (ctypes.c_uint32 * 24)(*_enc_aff(G1, (x, y), 12))) == 0


# ---- experimental FP64-pipe Montgomery product (tools/exp/mont_f64.cuh; round-2 candidate) --------
def test_f64_limb_product_montgomery():
    out = tempfile.mkdtemp(prefix="zk_host_sim_")
    so = os.path.join(out, "sim_f64.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DZK_HOST_SIM", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(SIM, "sim_f64.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.sim_f64_fma_rz.restype = ctypes.c_double
    lib.sim_f64_fma_rz.argtypes = [ctypes.c_double] * 3
    # the emulated fma.rz.f64 itself: exact a*b + c truncated toward zero to 53 bits
    rng = random.Random(31)
    for _ in range(2000):
        a, b = rng.getrandbits(48), rng.getrandbits(48)
        hi = lib.sim_f64_fma_rz(float(a), float(b), float(1 << 104))
        assert int(hi) - (1 << 104) == (a * b >> 52) << 52
        lo = lib.sim_f64_fma_rz(float(a), float(b), float((1 << 104) + (1 << 52)) - hi)
        assert int(lo) - (1 << 52) == a * b & ((1 << 52) - 1)
    l48 = lambda x: (ctypes.c_uint64 * 8)(*[(x >> (48 * i)) & ((1 << 48) - 1) for i in range(8)])
    n0 = (-pow(P, -1, 1 << 48)) % (1 << 48)
    Ri = pow(1 << 384, -1, P)
    vals = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (1 << 380) % P, RM] + [rng.randrange(P) for _ in range(3000)]
    o = (ctypes.c_uint64 * 8)()
    for i, a in enumerate(vals):
        b = vals[(i * 13 + 5) % len(vals)]
        lib.sim_f64_mul(l48(a), l48(b), l48(P), ctypes.c_uint64(n0), o)
        got = sum(int(o[k]) << (48 * k) for k in range(8))
        assert all(int(o[k]) < (1 << 48) for k in range(8))
        assert got == a * b * Ri % P, (hex(a), hex(b))
