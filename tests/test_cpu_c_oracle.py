"""The C restatement of the reference's fold-MSM (oracle/c/zkoracle.c — the CPU baseline of
bench.py) against the Python big-integer oracle: same bytes, single- and multi-threaded."""
import ctypes
import os
import random
import subprocess

from oracle import bls12_381 as O
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "c")], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "c", "libzkoracle.so"))
    lib.zkoracle_g1_msm_fold.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    lib.zkoracle_g1_mul.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.zkoracle_g1_msm_pippenger.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_void_p]
    return lib


def test_c_fold_msm_equals_python_oracle():
    lib = _lib()
    rng = random.Random(2)
    n = 24
    pts = [O.G1.mul(O.G1.one, rng.randrange(1, O.R)) for _ in range(n)]
    pts[3] = None                                   # identity base
    pts[7] = pts[6]                                 # repeated base
    ks = [rng.randrange(O.R) for _ in range(n)]
    ks[5], ks[6], ks[8] = 0, O.R - 1, 1
    exp = O.g1_to_uncompressed(H.oracle_msm(O.G1, pts, ks))
    out = (ctypes.c_uint8 * 96)()
    for threads in (1, 3, 64):
        lib.zkoracle_g1_msm_fold(H.g1_bytes(pts), H.scalars_bytes(ks), n, threads, out)
        assert bytes(out) == exp, threads
    # empty-ish: all scalars zero -> identity
    lib.zkoracle_g1_msm_fold(H.g1_bytes(pts), H.scalars_bytes([0] * n), n, 2, out)
    assert bytes(out) == O.g1_to_uncompressed(None)


def test_c_scalar_mul_known_answers():
    lib = _lib()
    out = (ctypes.c_uint8 * 96)()
    g = O.g1_to_uncompressed(O.G1.one)
    for k in (1, 2, 3, 0xDEADBEEF, O.R - 1, (O.R - 1) // 2):
        lib.zkoracle_g1_mul(g, O.fr_to_bytes(k), out)
        assert bytes(out) == O.g1_to_uncompressed(O.G1.mul(O.G1.one, k)), k
    lib.zkoracle_g1_mul(g, O.fr_to_bytes(0), out)
    assert bytes(out) == O.g1_to_uncompressed(None)


def test_c_pippenger_equals_the_fold():
    """The "fair CPU" bucket method against the reference's fold: same point."""
    lib = _lib()
    rng = random.Random(4)
    n = 300
    dl = [rng.randrange(1, 1 << 30) for _ in range(n)]
    base = O.G1.mul(O.G1.one, 12345)
    pts, cur = [], base
    for _ in range(n):                                # cheap distinct points: an addition chain
        pts.append(cur)
        cur = O.G1.add(cur, base)
    pts[3] = None
    pts[7] = pts[6]
    pts[9] = O.G1.neg(pts[8])
    ks = [rng.randrange(O.R) for _ in range(n)]
    ks[:8] = [0, 1, O.R - 1, (O.R - 1) // 2, 1 << 15, (1 << 16) - 1, 1 << 254, (1 << 255) % O.R]
    ks[8] = ks[9]                                     # k P + k (-P) = O
    bases, scal = H.g1_bytes(pts), H.scalars_bytes(ks)
    ref = (ctypes.c_uint8 * 96)()
    lib.zkoracle_g1_msm_fold(bases, scal, n, 8, ref)
    out = (ctypes.c_uint8 * 96)()
    for c, threads in ((2, 1), (4, 3), (13, 40), (16, 16), (16, 1), (20, 2)):
        assert lib.zkoracle_g1_msm_pippenger(bases, scal, n, c, threads, out) == 0
        assert bytes(out) == bytes(ref), (c, threads)
    # a handful of points against the Python oracle directly, and the all-zero case
    m = 5
    lib.zkoracle_g1_msm_pippenger(bases, scal, m, 16, 4, out)
    assert bytes(out) == O.g1_to_uncompressed(H.oracle_msm(O.G1, pts[:m], ks[:m]))
    lib.zkoracle_g1_msm_pippenger(bases, H.scalars_bytes([0] * n), n, 16, 4, out)
    assert bytes(out) == O.g1_to_uncompressed(None)
    assert lib.zkoracle_g1_msm_pippenger(bases, scal, n, 1, 1, out) == -1
