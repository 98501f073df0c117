"""Device Fp / Fr Montgomery arithmetic against Python big integers (bit-exact)."""
import ctypes
import random

import pytest

from oracle.bls12_381 import P, R

pytestmark = pytest.mark.gpu

OPS = {"mul": 0, "add": 1, "sub": 2, "neg": 3, "to_mont": 4, "from_mont": 5, "inverse": 6, "dbl": 7, "sqr": 8,
       "mul2_first": 9, "mul2_second": 10}


def _run(zk, field, op, a_vals, b_vals):
    from zukelang_b200 import _lib
    nl = 12 if field == 0 else 8
    n = len(a_vals)
    flat = lambda vals: (ctypes.c_uint32 * (nl * n))(*[(v >> (32 * i)) & 0xFFFFFFFF for v in vals for i in range(nl)])
    A, B, Out = flat(a_vals), flat(b_vals), (ctypes.c_uint32 * (nl * n))()
    _lib.check(zk.zk_test_field_op(field, op, A, B, Out, n))
    return [sum(int(Out[j * nl + i]) << (32 * i) for i in range(nl)) for j in range(n)]


@pytest.mark.parametrize("field,mod,nl", [(0, P, 12), (1, R, 8)])
def test_field_ops_match_bigint(zk, field, mod, nl):
    rng = random.Random(1234 + field)
    Rm = 1 << (32 * nl)
    Ri = pow(Rm, -1, mod)
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, Rm % mod, (1 << (32 * nl - 4)) % mod]
    a = edge + [rng.randrange(mod) for _ in range(2000)]
    b = [a[(i * 7 + 3) % len(a)] for i in range(len(a))]
    assert _run(zk, field, OPS["mul"], a, b) == [x * y * Ri % mod for x, y in zip(a, b)]
    assert _run(zk, field, OPS["add"], a, b) == [(x + y) % mod for x, y in zip(a, b)]
    assert _run(zk, field, OPS["sub"], a, b) == [(x - y) % mod for x, y in zip(a, b)]
    assert _run(zk, field, OPS["neg"], a, b) == [(-x) % mod for x in a]
    assert _run(zk, field, OPS["dbl"], a, b) == [2 * x % mod for x in a]
    assert _run(zk, field, OPS["to_mont"], a, b) == [x * Rm % mod for x in a]
    assert _run(zk, field, OPS["from_mont"], a, b) == [x * Ri % mod for x in a]
    inv_in = [x * Rm % mod for x in a[:64]]
    exp = [(pow(x, -1, mod) * Rm % mod) if x else 0 for x in a[:64]]
    assert _run(zk, field, OPS["inverse"], inv_in, inv_in) == exp


def _edge_values(mod, nl):
    """Inputs that stress the carry chains: the ends of the range, all-ones limb patterns, values with
    the top bits of the top limb set (as large as the modulus allows: the dedicated square needs
    2a + p < 2^(32 nl), csrc/mont.cuh)."""
    ones = (1 << (32 * nl)) - 1
    top = mod.bit_length()
    vals = [0, 1, 2, 3, mod - 1, mod - 2, mod - 3, (mod - 1) // 2, (mod + 1) // 2, ones % mod,
            (1 << (top - 1)), (1 << (top - 1)) - 1, (1 << (top - 1)) + 1, mod - (1 << 32), mod - (1 << 64) + 1,
            (mod >> 32 << 32) - 1,                      # top limbs of p, low limb all ones
            (mod >> 32 << 32) | 0xFFFFFFFE if ((mod >> 32 << 32) | 0xFFFFFFFE) < mod else mod - 4]
    for k in range(nl):                                  # one limb all ones / one limb zero
        vals.append((0xFFFFFFFF << (32 * k)) % mod)
        vals.append((mod - 1) & ~(0xFFFFFFFF << (32 * k)))
    for k in range(1, 4):                                # top k bits of the modulus' bit length set
        vals.append((((1 << k) - 1) << (top - k)) % mod)
    return [v % mod for v in vals]


@pytest.mark.parametrize("field,mod,nl", [(0, P, 12), (1, R, 8)])
def test_sqr_and_interleaved_pair_match_bigint(zk, field, mod, nl):
    """Mont::sqr (78 of 144 partial products for Fp) and Mont::mul2 (two products with interleaved
    CIOS rows) are what k_accumulate runs; compare both with Python integers on edge and random
    inputs, and with the plain product of the same library."""
    rng = random.Random(4321 + field)
    Ri = pow(1 << (32 * nl), -1, mod)
    a = _edge_values(mod, nl) + [rng.randrange(mod) for _ in range(4000)]
    b = [a[(i * 11 + 5) % len(a)] for i in range(len(a))]
    sq = _run(zk, field, OPS["sqr"], a, a)
    assert sq == [x * x * Ri % mod for x in a]
    assert sq == _run(zk, field, OPS["mul"], a, a)
    assert _run(zk, field, OPS["mul2_first"], a, b) == [x * y * Ri % mod for x, y in zip(a, b)]
    assert _run(zk, field, OPS["mul2_second"], a, b) == [(x + y) * (x - y) * Ri % mod for x, y in zip(a, b)]


def test_mixed_add_variants_match_the_oracle(zk):
    """XYZZ::madd, XYZZ::madd_paired (the bucket accumulation's) and XYZZ::add on the same operands:
    2p + q with q generic, q = identity, q = 2p (doubling branch), q = -2p (cancellation), p = identity."""
    import ctypes
    from oracle import bls12_381 as O
    from tests import helpers as H
    from zukelang_b200 import _lib
    rng = random.Random(77)
    ps, qs = [], []
    for i in range(48):
        p = O.G1.mul(O.G1.one, rng.randrange(1, R))
        q = O.G1.mul(O.G1.one, rng.randrange(1, R))
        ps.append(p)
        qs.append(q)
    p0 = ps[0]
    two_p0 = O.G1.add(p0, p0)
    ps += [p0, p0, p0, None, None]
    qs += [None, two_p0, O.G1.neg(two_p0), qs[0], None]
    expect = b"".join(H.expect_g1(O.G1.add(O.G1.add(p, p), q)) for p, q in zip(ps, qs))
    n = len(ps)
    for variant in (0, 1, 2):
        out = (ctypes.c_uint8 * (144 * n))()
        _lib.check(zk.zk_test_g1_madd(H.g1_bytes(ps), H.g1_bytes(qs), variant, n, out))
        assert bytes(out) == expect, "variant %d" % variant
