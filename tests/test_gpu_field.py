"""Device Fp / Fr Montgomery arithmetic against Python big integers (bit-exact)."""
import ctypes
import random

import pytest

from oracle.bls12_381 import P, R

pytestmark = pytest.mark.gpu

OPS = {"mul": 0, "add": 1, "sub": 2, "neg": 3, "to_mont": 4, "from_mont": 5, "inverse": 6, "dbl": 7}


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
