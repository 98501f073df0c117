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
