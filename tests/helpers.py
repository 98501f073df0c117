"""Shared helpers for the parity tests: oracle <-> C-ABI byte conversions."""
import ctypes
import random

from oracle import bls12_381 as O

R, P = O.R, O.P


def scalars_bytes(ks):
    return b"".join(O.fr_to_bytes(k) for k in ks)


def g1_bytes(pts):
    return b"".join(O.g1_to_uncompressed(p) for p in pts)


def g2_bytes(pts):
    return b"".join(O.g2_to_uncompressed(p) for p in pts)


def out_buf(n):
    return (ctypes.c_uint8 * n)()


def expect_g1(pt):
    return O.g1_to_uncompressed(pt) + O.g1_compress(pt)


def expect_g2(pt):
    return O.g2_to_uncompressed(pt) + O.g2_compress(pt)


def oracle_msm(G, pts, ks):
    """curve.ml:91-103 — left fold of scalar_mul + add."""
    acc = None
    for p, k in zip(pts, ks):
        acc = G.add(G.mul(p, k), acc)
    return acc


def random_points(G, n, rng, small=True):
    """n points with known discrete logs (dlogs returned too)."""
    dl = [rng.randrange(1, 1 << 20) if small else rng.randrange(1, R) for _ in range(n)]
    return [G.mul(G.one, d) for d in dl], dl
