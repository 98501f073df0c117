"""Multi-GPU plumbing for the prover path: one process per GPU, every MSM sharded
by base-point range, ONE tiny gather of partial sums per proof.

``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests) only moves the partial points;
the additions run in the CUDA library (``zk_g1_sum`` / ``zk_g2_sum``).  ``engine`` is the object
that performs group arithmetic: the default is the C-ABI library; tests/ inject an oracle-backed
engine to exercise the sharding logic on CPU-only machines (never done in the product)."""

from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

from . import _lib

G1_RAW, G2_RAW = 96, 192


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of shard `rank` — the same rule csrc/prove.cu:slice applies to every key field."""
    return n * rank // world, n * (rank + 1) // world


class CudaEngine:
    """Group arithmetic through libzkb200."""

    def msm_g1(self, bases: bytes, scalars: bytes, n: int) -> bytes:
        out = (ctypes.c_uint8 * _lib.G1_OUT)()
        _lib.check(_lib.lib().zk_g1_msm(bases, None, scalars, n, out))
        return bytes(out)

    def sum_g1(self, points: bytes, k: int) -> bytes:
        out = (ctypes.c_uint8 * _lib.G1_OUT)()
        _lib.check(_lib.lib().zk_g1_sum(points, k, out))
        return bytes(out)

    def sum_g2(self, points: bytes, k: int) -> bytes:
        out = (ctypes.c_uint8 * _lib.G2_OUT)()
        _lib.check(_lib.lib().zk_g2_sum(points, k, out))
        return bytes(out)


def all_gather_bytes(local: bytes, group=None) -> List[bytes]:
    """Gather equal-length byte strings from every rank (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [local]
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.frombuffer(bytearray(local), dtype=torch.uint8).to(dev)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, t, group=group)
    return [bytes(o.cpu().numpy()) for o in outs]


def sharded_g1_msm(bases: bytes, scalars: bytes, n: int, rank: int, world: int, engine=None, group=None) -> bytes:
    """sum_i scalars[i] * bases[i] with the index range split over `world` ranks.  Every rank passes
    the full arrays (or at least its own slice at the right offsets) and gets the full result."""
    engine = engine or CudaEngine()
    lo, hi = shard_bounds(n, rank, world)
    if hi > lo:
        part = engine.msm_g1(bases[lo * G1_RAW:hi * G1_RAW], scalars[lo * 32:hi * 32], hi - lo)[:G1_RAW]
    else:
        part = bytes([0x40]) + bytes(G1_RAW - 1)
    parts = all_gather_bytes(part, group)
    return engine.sum_g1(b"".join(parts), len(parts))


# offsets of (a, b, c) inside a Groth16 "proof out" buffer: uncompressed | compressed each
_G16 = ((0, G1_RAW, False), (_lib.G1_OUT, G2_RAW, True), (_lib.G1_OUT + _lib.G2_OUT, G1_RAW, False))


def combine_groth16(partials: Sequence[bytes], engine=None) -> bytes:
    """Add the shards' partial (a, b, c) — zk_groth16_prove outputs with shard_count > 1 — into the
    proof-out buffer of the whole proof."""
    if engine is None:                                    # the library adds all three elements in one call
        buf = b"".join(partials)
        if len(buf) != len(partials) * _lib.GROTH16_PROOF_OUT:
            raise ValueError("combine_groth16: every partial is a %d-byte proof-out buffer" % _lib.GROTH16_PROOF_OUT)
        res = (ctypes.c_uint8 * _lib.GROTH16_PROOF_OUT)()
        _lib.check(_lib.lib().zk_groth16_combine(buf, len(partials), res))
        return bytes(res)
    out = b""
    for off, raw, is_g2 in _G16:
        pts = b"".join(p[off:off + raw] for p in partials)
        out += (engine.sum_g2 if is_g2 else engine.sum_g1)(pts, len(partials))
    return out


def combine_pinocchio(partials: Sequence[bytes], engine=None) -> bytes:
    from .pinocchio import PROOF_IS_G2
    engine = engine or CudaEngine()
    out, off = b"", 0
    for is_g2 in PROOF_IS_G2:
        raw, tot = (G2_RAW, _lib.G2_OUT) if is_g2 else (G1_RAW, _lib.G1_OUT)
        pts = b"".join(p[off:off + raw] for p in partials)
        out += (engine.sum_g2 if is_g2 else engine.sum_g1)(pts, len(partials))
        off += tot
    return out
