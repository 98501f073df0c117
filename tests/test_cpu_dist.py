"""Multi-process sharding logic on CPU: world_size 2 over gloo.  The group arithmetic is supplied by
an oracle-backed engine (test-only substitution); what is under test is the host logic of
zukelang_b200/dist.py: shard bounds, the gather, and the combination of partial sums."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bls12_381 as O
from tests import helpers as H


class OracleEngine:
    def msm_g1(self, bases, scalars, n):
        pts = [O.g1_from_uncompressed(bases[i * 96:(i + 1) * 96]) for i in range(n)]
        ks = [O.fr_from_bytes(scalars[i * 32:(i + 1) * 32]) for i in range(n)]
        return H.expect_g1(H.oracle_msm(O.G1, pts, ks))

    def sum_g1(self, points, k):
        return H.expect_g1(O.G1.sum([O.g1_from_uncompressed(points[i * 96:(i + 1) * 96]) for i in range(k)]))

    def sum_g2(self, points, k):
        return H.expect_g2(O.G2.sum([O.g2_from_uncompressed(points[i * 192:(i + 1) * 192]) for i in range(k)]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zukelang_b200 import dist as D
    rng = random.Random(11)
    pts, _ = H.random_points(O.G1, n, rng)
    ks = [rng.randrange(O.R) for _ in range(n)]
    got = D.sharded_g1_msm(H.g1_bytes(pts), H.scalars_bytes(ks), n, rank, world, engine=OracleEngine())
    q.put((rank, got))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 2])
def test_sharded_msm_world2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = random.Random(11)
    pts, _ = H.random_points(O.G1, n, rng)
    ks = [rng.randrange(O.R) for _ in range(n)]
    exp = H.expect_g1(H.oracle_msm(O.G1, pts, ks))
    assert res[0] == exp and res[1] == exp


def test_shard_bounds_cover_and_match_the_library_rule():
    from zukelang_b200.dist import shard_bounds
    for n in (0, 1, 5, 1 << 20, (1 << 20) - 1):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_combine_groth16_adds_partials():
    from zukelang_b200 import dist as D
    rng = random.Random(4)
    parts, tot = [], [None, None, None]
    for _ in range(3):
        a, b, c = (O.G1.mul(O.G1.one, rng.randrange(O.R)), O.G2.mul(O.G2.one, rng.randrange(O.R)),
                   O.G1.mul(O.G1.one, rng.randrange(O.R)))
        parts.append(H.expect_g1(a) + H.expect_g2(b) + H.expect_g1(c))
        tot = [O.G1.add(tot[0], a), O.G2.add(tot[1], b), O.G1.add(tot[2], c)]
    got = D.combine_groth16(parts, engine=OracleEngine())
    assert got == H.expect_g1(tot[0]) + H.expect_g2(tot[1]) + H.expect_g1(tot[2])
