"""G1 / G2 MSM through the C ABI against the oracle's restatement of the
reference folds (curve.ml:91-118).  Bit-exact on uncompressed + compressed bytes."""
import ctypes
import random

import pytest

from oracle import bls12_381 as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
R = O.R

GROUPS = {
    "G1": (O.G1, H.g1_bytes, H.expect_g1, 144, "zk_g1_msm", "zk_g1_table_load", "zk_g1_table_msm", "zk_g1_fixed_base_mul", 96),
    "G2": (O.G2, H.g2_bytes, H.expect_g2, 288, "zk_g2_msm", "zk_g2_table_load", "zk_g2_table_msm", "zk_g2_fixed_base_mul", 192),
}


def _msm(zk, gname, pts, ks, inf_flags=None):
    from zukelang_b200 import _lib
    G, enc, _, outn, fn = GROUPS[gname][:5]
    out = H.out_buf(outn)
    flags = bytes(inf_flags) if inf_flags is not None else None
    _lib.check(getattr(zk, fn)(enc(pts), flags, H.scalars_bytes(ks), len(pts), out))
    return bytes(out)


@pytest.mark.parametrize("gname", ["G1", "G2"])
@pytest.mark.parametrize("n", [1, 2, 3, 17, 255, 256])
def test_msm_random(zk, gname, n):
    G, _, expect = GROUPS[gname][:3]
    rng = random.Random(n * 31 + len(gname))
    if gname == "G2" and n > 64:
        n = 64 + (n % 7)
    pts, _ = H.random_points(G, n, rng)
    ks = [rng.randrange(R) for _ in range(n)]
    assert _msm(zk, gname, pts, ks) == expect(H.oracle_msm(G, pts, ks))


@pytest.mark.parametrize("gname", ["G1", "G2"])
def test_msm_edge_scalars_and_bases(zk, gname):
    """Edge scalars {0, 1, r-1, 2^k} and edge bases {O, P and -P, repeated P}."""
    G, _, expect = GROUPS[gname][:3]
    rng = random.Random(99)
    p = G.mul(G.one, 7)
    q = G.mul(G.one, 11)
    pts = [p, G.neg(p), p, p, None, q, q, G.one, q, p]
    ks = [5, 5, 0, 1, 1234, R - 1, 1 << 254 if (1 << 254) < R else 1 << 200, 1 << 64, 1 << 15, (1 << 16) - 1]
    ks = [k % R for k in ks]
    assert _msm(zk, gname, pts, ks) == expect(H.oracle_msm(G, pts, ks))
    # sum cancels to the identity
    assert _msm(zk, gname, [p, G.neg(p)], [9, 9]) == expect(None)
    # all scalars zero
    assert _msm(zk, gname, [p, q], [0, 0]) == expect(None)
    # identity given through inf_flags rather than the 0x40 encoding
    assert _msm(zk, gname, [p, q], [3, 4], inf_flags=[0, 1]) == expect(G.mul(p, 3))
    # repeated point with many equal scalars hits the P + P (doubling) branch of the mixed add
    assert _msm(zk, gname, [p] * 8, [1] * 8) == expect(G.mul(p, 8))


def test_msm_empty_is_identity(zk):
    from zukelang_b200 import _lib
    out = H.out_buf(144)
    _lib.check(zk.zk_g1_msm(None, None, None, 0, out))
    assert bytes(out) == H.expect_g1(None)


def test_msm_rejects_bad_inputs(zk):
    from zukelang_b200 import _lib
    out = H.out_buf(144)
    bad_point = bytes(95) + b"\x01"                       # (0, 1) is not on the curve
    assert zk.zk_g1_msm(bad_point, None, O.fr_to_bytes(1), 1, out) == _lib.ZK_EPOINT
    good = O.g1_to_uncompressed(O.G1.one)
    assert zk.zk_g1_msm(good, None, (R).to_bytes(32, "little"), 1, out) == _lib.ZK_EPOINT  # scalar == r
    with pytest.raises(_lib.InvalidArgument):
        _lib.check(zk.zk_g1_msm(good, None, None, 1, out))


@pytest.mark.parametrize("gname,n", [("G1", 1000), ("G1", 4096), ("G2", 300)])
@pytest.mark.parametrize("precompute", [0, 1])
def test_table_msm_known_dlog(zk, gname, n, precompute):
    """Resident table (with and without the precomputed windows): bases k_i * G from the
    fixed-base kernel, result checked against (sum s_i k_i) * G — exact at any size."""
    from zukelang_b200 import _lib
    G, enc, expect, outn, _, load, msm, fixed, raw = GROUPS[gname]
    rng = random.Random(n + precompute)
    dl = [rng.randrange(R) for _ in range(n)]
    dl[3] = 0                                              # an identity base inside the table
    bases = (ctypes.c_uint8 * (raw * n))()
    _lib.check(getattr(zk, fixed)(H.scalars_bytes(dl), n, bases))
    # spot-check the fixed-base kernel itself against the oracle
    for i in (0, 3, n - 1):
        pt = G.mul(G.one, dl[i])
        exp = O.g1_to_uncompressed(pt) if gname == "G1" else O.g2_to_uncompressed(pt)
        assert bytes(bases[i * raw:(i + 1) * raw]) == exp
    h = ctypes.c_uint64()
    _lib.check(getattr(zk, load)(bases, None, n, precompute, 0, ctypes.byref(h)))
    try:
        for trial in range(2):
            ks = [rng.randrange(R) for _ in range(n)]
            if trial == 1:                                 # witness-like: mostly 0 / 1
                ks = [k if rng.random() < 0.1 else rng.randrange(2) for k in ks]
            out = H.out_buf(outn)
            _lib.check(getattr(zk, msm)(h.value, H.scalars_bytes(ks), n, out))
            total = sum(s * d for s, d in zip(ks, dl)) % R
            assert bytes(out) == expect(G.mul(G.one, total))
        # prefix MSM (fewer scalars than table points)
        m = n // 3
        ks = [rng.randrange(R) for _ in range(m)]
        out = H.out_buf(outn)
        _lib.check(getattr(zk, msm)(h.value, H.scalars_bytes(ks), m, out))
        assert bytes(out) == expect(G.mul(G.one, sum(s * d for s, d in zip(ks, dl)) % R))
    finally:
        _lib.check(zk.zk_table_free(h.value))


def test_table_msm_batch_and_pipelined_dev(zk):
    """zk_g1_table_msm_batch (double-buffered uploads, batched tails) and the pipelined *_dev calls
    give the same points as one-at-a-time calls; more MSMs than the tail queue holds."""
    import numpy as np
    import torch
    from zukelang_b200 import _lib
    n, count = 3000, 35            # the batch call queues up to 32 tails: 35 overflows the queue once
    rng = random.Random(77)
    dl = [rng.randrange(R) for _ in range(n)]
    bases = (ctypes.c_uint8 * (96 * n))()
    _lib.check(zk.zk_g1_fixed_base_mul(H.scalars_bytes(dl), n, bases))
    h = ctypes.c_uint64()
    _lib.check(zk.zk_g1_table_load(bases, None, n, 1, 0, ctypes.byref(h)))
    try:
        vecs, expect = [], []
        for i in range(count):
            ks = [rng.randrange(R) for _ in range(n)]
            vecs.append(ctypes.create_string_buffer(H.scalars_bytes(ks), 32 * n))
            expect.append(H.expect_g1(O.G1.mul(O.G1.one, sum(a * b for a, b in zip(ks, dl)) % R)))
        ptrs = (ctypes.c_void_p * count)(*[ctypes.addressof(v) for v in vecs])
        out = (ctypes.c_uint8 * (144 * count))()
        _lib.check(zk.zk_g1_table_msm_batch(h.value, ptrs, n, count, out))
        assert [bytes(out[i * 144:(i + 1) * 144]) for i in range(count)] == expect
        # pipelined device calls + join, queue depth 4: the 5th, 9th, ... call finds the queue full
        _lib.check(zk.zk_table_pipeline(h.value, 4))
        d_sc = [torch.from_numpy(np.frombuffer(v.raw, dtype=np.uint8).copy()).cuda() for v in vecs]
        d_out = torch.zeros(count, 144, dtype=torch.uint8, device="cuda")
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for i in range(count):
                _lib.check(zk.zk_g1_table_msm_dev(h.value, d_sc[i].data_ptr(), n, d_out[i].data_ptr(), st.cuda_stream))
            _lib.check(zk.zk_table_join(h.value, st.cuda_stream))
        st.synchronize()
        assert [bytes(d_out[i].cpu().numpy()) for i in range(count)] == expect
        _lib.check(zk.zk_table_pipeline(h.value, 0))
        # stage events: every join is recorded with the MSMs it held (zk_table_profile_totals)
        _lib.check(zk.zk_table_profile(h.value, 1, None))
        _lib.check(zk.zk_table_pipeline(h.value, 4))
        d_out.zero_()
        with torch.cuda.stream(st):
            for i in range(6):                      # a full queue of 4 is joined by the 5th call, the rest by the join
                _lib.check(zk.zk_g1_table_msm_dev(h.value, d_sc[i].data_ptr(), n, d_out[i].data_ptr(), st.cuda_stream))
            _lib.check(zk.zk_table_join(h.value, st.cuda_stream))
        st.synchronize()
        assert [bytes(d_out[i].cpu().numpy()) for i in range(6)] == expect[:6]
        totals, counts, last = (ctypes.c_float * 4)(), (ctypes.c_uint64 * 2)(), (ctypes.c_float * 4)()
        _lib.check(zk.zk_table_profile_totals(h.value, totals, counts))
        _lib.check(zk.zk_table_profile(h.value, 0, last))
        assert (counts[0], counts[1]) == (6, 2)
        assert all(t > 0 for t in totals) and all(0 < last[i] <= totals[i] for i in range(4))
        _lib.check(zk.zk_table_pipeline(h.value, 0))
    finally:
        _lib.check(zk.zk_table_free(h.value))


def test_g1_msm_full_size_2e20_known_dlog(zk):
    """BASELINE's headline size: 2^20 points, uniform scalars, resident precomputed table.  No CPU
    oracle can fold 2^20 scalar multiplications in seconds, so the check is the size-independent
    identity  sum s_i (d_i G) = (sum s_i d_i mod r) G  plus linearity  MSM(s) + MSM(t) = MSM(s + t)."""
    import numpy as np
    from zukelang_b200 import _lib
    n = 1 << 20
    rng = np.random.Generator(np.random.PCG64(20))

    def scalars():
        w = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        w[:, 3] &= np.uint64((1 << 61) - 1)                  # < 2^253 < r : canonical, sums stay canonical
        ints = [int(a) | (int(b) << 64) | (int(c) << 128) | (int(d) << 192) for a, b, c, d in w.tolist()]
        return w, ints

    dl_w, dl = scalars()
    bases = np.empty(n * 96, dtype=np.uint8)
    _lib.check(zk.zk_g1_fixed_base_mul(dl_w.ctypes.data, n, bases.ctypes.data))
    h = ctypes.c_uint64()
    _lib.check(zk.zk_g1_table_load(bases.ctypes.data, None, n, 1, 0, ctypes.byref(h)))
    try:
        s_w, s = scalars()
        t_w, t = scalars()
        u = [(a + b) % R for a, b in zip(s, t)]
        u_w = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in u), dtype=np.uint64).copy()
        outs = []
        for w in (s_w, t_w, u_w):
            out = H.out_buf(144)
            _lib.check(zk.zk_g1_table_msm(h.value, w.ctypes.data, n, out))
            outs.append(bytes(out))
        exp = np.empty(96 * 3, dtype=np.uint8)
        tots = [sum(a * b for a, b in zip(v, dl)) % R for v in (s, t, u)]
        _lib.check(zk.zk_g1_fixed_base_mul(H.scalars_bytes(tots), 3, exp.ctypes.data))
        for i in range(3):
            assert outs[i][:96] == bytes(exp[96 * i:96 * (i + 1)])
        both = H.out_buf(144)                                # linearity through the point adder
        _lib.check(zk.zk_g1_sum(outs[0][:96] + outs[1][:96], 2, both))
        assert bytes(both) == outs[2]
    finally:
        _lib.check(zk.zk_table_free(h.value))


def test_point_sums(zk):
    from zukelang_b200 import _lib
    rng = random.Random(12)
    for gname, G, enc, expect, fn, outn in (("G1", O.G1, H.g1_bytes, H.expect_g1, "zk_g1_sum", 144),
                                            ("G2", O.G2, H.g2_bytes, H.expect_g2, "zk_g2_sum", 288)):
        pts = [G.mul(G.one, rng.randrange(R)) for _ in range(5)] + [None]
        pts.append(G.neg(pts[0]))
        pts.append(pts[1])                                   # repeated point: doubling branch
        out = H.out_buf(outn)
        _lib.check(getattr(zk, fn)(enc(pts), len(pts), out))
        assert bytes(out) == expect(G.sum(pts))
    bad = bytes(95) + b"\x01"
    assert zk.zk_g1_sum(bad, 1, H.out_buf(144)) == _lib.ZK_EPOINT
