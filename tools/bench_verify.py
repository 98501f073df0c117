#!/usr/bin/env python3
"""Latency of the verifier-side entry points (not a headline metric):
zk_pairing_product at 1 / 3 / 4 pairs, Groth16 verify, and zk_g*_decompress throughput.
Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from zukelang_b200 import _lib  # noqa: E402
from zukelang_b200.curve import Bls12_381 as C  # noqa: E402


def timed(fn, reps=5):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps * 1e3


def main():
    _lib.lib()
    g1 = C.G1.fixed_base([3, 5, 7, 11])
    g2 = C.G2.fixed_base([13, 17, 19, 23])
    res = {}
    for n in (1, 3, 4):
        res["pairing_product_%d_ms" % n] = round(timed(lambda: C.Pairing.product(list(zip(g1[:n], g2[:n])))), 2)
    n = 4096
    pts1 = C.G1.fixed_base(list(range(1, n + 1)))
    pts2 = C.G2.fixed_base(list(range(1, n + 1)))
    c1 = [C.G1.to_compressed_bytes(p) for p in pts1]
    c2 = [C.G2.to_compressed_bytes(p) for p in pts2]
    assert C.G1.of_compressed_bytes_many(c1) == pts1 and C.G2.of_compressed_bytes_many(c2) == pts2
    res["g1_decompress_4096_ms"] = round(timed(lambda: C.G1.of_compressed_bytes_many(c1), 3), 2)
    res["g2_decompress_4096_ms"] = round(timed(lambda: C.G2.of_compressed_bytes_many(c2), 3), 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
