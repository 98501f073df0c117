#!/usr/bin/env python3
"""Groth16 prove throughput on BASELINE.json's synthetic circuits (configs[2] = 2^16, configs[4] = 2^20)
through zk_groth16_prove_r1cs (the evaluation-form prover):

  mulchain   c_0 = x, c_{i+1} = c_i * x          (the named multiply chain; its B-query scalars W(j) = x
                                                  are all equal, so the G2 MSM piles into one bucket per window)
  r1cs       seeded random R1CS, ~3 non-zeros per row: every gate has its own W(j)

Three ways to use the GPUs:
  one process, one device           python tools/bench_groth16.py --logn 16 20
  one process, all devices (C ABI)  ZKB200_DEVICES=0,1,2,3,4,5,6,7 python tools/bench_groth16.py --logn 20
  one process per GPU (torchrun)    torchrun --nproc-per-node 8 tools/bench_groth16.py --logn 20
    (every rank holds shard (rank, world) of the key; the 576-byte partial results are all-gathered and added)

The last timed proof of every run is checked, byte for byte, against the closed-form trapdoor identity
(SURVEY.md §8c iv).  One JSON line per (size, circuit).  bench.py imports run() for its `groth16` field."""
import argparse
import ctypes
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from zukelang_b200 import _lib, sparse as S
from zukelang_b200.curve import R, fr_vector

ONE = ("ONE", 1)
SEED = 0x47524F54            # "GROT", SURVEY.md §8d config 5


def mulchain(n):
    x = ("input", 2)
    cs = [x] + [("_tmp", 3 + i) for i in range(n - 1)]
    out = ("v", 3 + n)
    gates = [({cs[i + 1]: 1}, {cs[i]: 1}, {x: 1}) for i in range(n - 1)]
    gates.append(({out: 1}, {cs[-1]: 1, ONE: 3}, {ONE: 1}))
    circ = S.SparseCircuit(gates, [ONE], [out], cs)

    def witness(xv):
        sol = {ONE: 1, x: xv}
        cur = xv
        for i in range(1, n):
            cur = cur * xv % R
            sol[cs[i]] = cur
        sol[out] = (cur + 3) % R
        return sol
    return circ, witness


def random_r1cs(n, seed=SEED):
    """Gate i: z_i = (a x_p + b x_q + k ONE) * (c x_s + d x_u), operands among the variables defined
    before gate i, coefficients uniform in Fr.  Same draws as oracle/zk.py:circuit_random_r1cs (the
    parity tests build the circuit from the oracle's generator; tests/test_cpu_wire.py checks that
    the two agree)."""
    rng = random.Random(seed + n)
    xs = [("input", 2), ("input", 3)]
    zs = [("_tmp", 4 + i) for i in range(n - 1)] + [("v", 3 + n)]
    defined = list(xs)
    gates, plan = [], []
    for i in range(n):
        p, q, s_, u = (defined[rng.randrange(len(defined))] for _ in range(4))
        a, b, c, d, k = (rng.randrange(1, R) for _ in range(5))
        l, r = {}, {}
        for var, co in ((p, a), (q, b), (ONE, k)):
            l[var] = (l.get(var, 0) + co) % R
        for var, co in ((s_, c), (u, d)):
            r[var] = (r.get(var, 0) + co) % R
        gates.append(({zs[i]: 1}, l, r))
        plan.append((zs[i], sorted(l.items()), sorted(r.items())))
        defined.append(zs[i])
    circ = S.SparseCircuit(gates, [ONE], [zs[-1]], xs + zs[:-1])

    def witness(seed2):
        r2 = random.Random(seed2)
        sol = {ONE: 1, xs[0]: r2.randrange(R), xs[1]: r2.randrange(R)}
        for z, l, r in plan:
            sol[z] = sum(c * sol[v] for v, c in l) % R * (sum(c * sol[v] for v, c in r) % R) % R
        live = set(circ.variables)
        return {k: v for k, v in sol.items() if k in live}
    return circ, witness


CIRCUITS = {"mulchain": mulchain, "r1cs": random_r1cs}


def fixed_base(zk, group, scalars):
    raw = 96 if group == "g1" else 192
    if not scalars:
        return b""
    out = (ctypes.c_uint8 * (raw * len(scalars)))()
    _lib.check(getattr(zk, "zk_%s_fixed_base_mul" % group)(fr_vector(scalars), len(scalars), out))
    return bytes(out)


def shard_slice(length, idx, cnt):
    """The rule of csrc/prove.cu:slice."""
    return length * idx // cnt, length * (idx + 1) // cnt


def load_key(zk, circ, td, w, shard):
    """Key generation for shard (idx, cnt): the trapdoor scalars on the host (Groth16Sparse.keygen_scalars),
    every group element from the fixed-base kernel — and only the points this shard keeps."""
    idx, cnt = shard
    n = circ.n
    a, b, gm, d, t = td
    sc = S.Groth16Sparse().keygen_scalars(td, circ, w)
    mids = sc["mids"]
    ltd = [sc["ltd"][k] for k in mids]
    lo_t, hi_t = shard_slice(n, idx, cnt)
    lo_h, hi_h = shard_slice(n, idx, cnt)          # n_h = n in the evaluation-form key
    lo_m, hi_m = shard_slice(len(mids), idx, cnt)
    g1 = fixed_base(zk, "g1", [a, b, d] + sc["lag"][lo_t:hi_t] + sc["hk"][lo_h:hi_h] + ltd[lo_m:hi_m])
    g2 = fixed_base(zk, "g2", [b, d] + sc["lag"][lo_t:hi_t])
    pos = {k: i for i, k in enumerate(circ.variables)}
    idx_arr = (ctypes.c_uint32 * max(len(mids), 1))(*[pos[k] for k in mids])
    keep = []

    def field(blob, first, count, size, lo):
        """A list-valued key field of which only [lo, lo + count) is materialised: the C side reads
        exactly that slice (csrc/prove.cu), so the pointer is biased by -lo elements."""
        buf = ctypes.create_string_buffer(blob[first * size:(first + count) * size], max(count * size, 1))
        keep.append(buf)
        return ctypes.addressof(buf) - lo * size

    nt, nh, nm = hi_t - lo_t, hi_h - lo_h, hi_m - lo_m
    st = _lib.Groth16PKeyStruct(
        n=n, m=len(circ.variables), n_mid=len(mids), n_h=n, mid_index=ctypes.addressof(idx_arr),
        a=field(g1, 0, 1, 96, 0), b1=field(g1, 1, 1, 96, 0), d1=field(g1, 2, 1, 96, 0),
        ti1=field(g1, 3, nt, 96, lo_t), tiztd=field(g1, 3 + nt, nh, 96, lo_h), ltd_mid=field(g1, 3 + nt + nh, nm, 96, lo_m),
        b2=field(g2, 0, 1, 192, 0), d2=field(g2, 1, 1, 192, 0), ti2=field(g2, 2, nt, 192, lo_t))
    h = ctypes.c_uint64()
    _lib.check(zk.zk_groth16_pk_load(ctypes.byref(st), idx, cnt, ctypes.byref(h)))
    return h.value, len(mids)


def run(zk, logn, iters, circuit="mulchain", dist=None, quiet=False):
    """Times `iters` proofs (after one warm-up proof).  dist = torch.distributed when the key is sharded
    over the ranks of a torchrun job; every rank then returns the same record.  Raises on a proof that
    differs from the closed form."""
    import torch
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    n = 1 << logn
    t0 = time.time()
    circ, witness = CIRCUITS[circuit](n)
    dom = S.EvalDomain(circ)
    rng = random.Random(SEED + logn)
    td = tuple(rng.randrange(1, R) for _ in range(5))
    h, n_mid = load_key(zk, circ, td, dom.w, (rank, world))
    setup_s = time.time() - t0
    sols = []
    m = len(circ.variables)
    per = (m + world - 1) // world                     # scalars of the witness each rank uploads when sharded
    for i in range(2):
        sol = witness(rng.randrange(R))
        # the witness the caller hands to prove lives in PINNED host memory (the upload is then one
        # asynchronous DMA at link speed; from pageable memory the same 32 MB take ~3 ms at 2^20)
        raw = bytearray(fr_vector(sol[k] for k in circ.variables)) + bytearray(32 * (per * world - m))
        pinned = torch.frombuffer(raw, dtype=torch.uint8).pin_memory()
        sols.append((sol, pinned))
    if dist:
        # one process per GPU: every rank uploads ITS 1/N of the witness and the ranks all-gather it on
        # the devices (NVLink) instead of pushing the whole witness down all N PCIe links at once
        d_part = torch.empty(32 * per, dtype=torch.uint8, device="cuda")
        d_full = torch.empty(32 * per * world, dtype=torch.uint8, device="cuda")
    out = (ctypes.c_uint8 * _lib.GROTH16_PROOF_OUT)()
    wall, dev, ok = [], [], True
    stages = []
    ms = ctypes.c_float()
    st_ms = (ctypes.c_float * 8)()
    if dist:
        from zukelang_b200 import dist as D
    for it in range(iters + 1):
        sol, sol_b = sols[it & 1]
        r, s = rng.randrange(R), rng.randrange(R)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if dist:
            d_part.copy_(sol_b[32 * per * rank:32 * per * (rank + 1)], non_blocking=True)
            dist.all_gather_into_tensor(d_full, d_part)
            torch.cuda.current_stream().synchronize()      # the library runs on its own stream
            sol_ptr = d_full.data_ptr()
        else:
            sol_ptr = sol_b.data_ptr()
        _lib.check(zk.zk_groth16_prove_r1cs(h, dom.handle, sol_ptr, r.to_bytes(32, "little"), s.to_bytes(32, "little"), out))
        proof = bytes(out)
        if dist:
            proof = D.combine_groth16(D.all_gather_bytes(proof))
        dt = time.perf_counter() - t1
        _lib.check(zk.zk_groth16_last_device_ms(h, ctypes.byref(ms)))
        if dist:
            t = torch.tensor([dt, ms.value], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dms = float(t[0].item()), float(t[1].item())
        else:
            dms = ms.value
        _lib.check(zk.zk_groth16_last_stage_ms(h, st_ms))
        if it:
            wall.append(dt * 1e3)
            dev.append(dms)
            stages.append([round(float(x), 3) for x in st_ms])
        if it == iters:                    # the last TIMED proof is the one checked (closed form: ~10 s of host work at 2^20)
            A, B, C = S.closed_form_scalars(td, r, s, circ, dom.w, sol)
            ok &= proof[0:96] + proof[432:528] == fixed_base(zk, "g1", [A, C]) and proof[144:336] == fixed_base(zk, "g2", [B])
    med = lambda xs: sorted(xs)[len(xs) // 2]
    ndev = zk.zk_device_count()
    rec = {"probe": "groth16", "log_n": logn, "constraints": n, "variables": len(circ.variables), "circuit": circuit,
           "processes": world, "devices_per_process": ndev,
           "prove_ms": med(wall), "prove_ms_best": min(wall), "prove_ms_all": wall,
           "device_ms": med(dev), "device_ms_all": dev, "proofs_per_s": 1e3 / med(wall), "exact_ok": bool(ok),
           "stages_ms": dict(zip(("upload", "qap_values", "B_g2_scalars_sort_accumulate", "quotient_h", "AC_g1_scalars_sort", "AC_g1_accumulate", "tails", "combine_download"),
                                 stages[len(stages) // 2])),
           "stages_note": "CUDA events on this process' primary stream, one timed proof (rank 0's when sharded)",
           "setup_s": setup_s, "witness_memory": "pinned host" + ("; each rank uploads 1/%d of it, all_gather on the devices" % world if dist else ""),
           "h2d_bytes_per_proof": 32 * (per if dist else m) + 64, "d2h_bytes_per_proof": 576,
           "timing": "wall clock around the C-ABI prove call%s, median of %d (max over ranks); device_ms = CUDA events inside "
                     "the call, first upload to last download" % (" (and, before it, this rank's 1/N witness upload + the witness all_gather) + all_gather of the 576-B partials + zk_g*_sum" if dist else "", iters),
           "msm_points": {"A_g1": n + 3, "C_g1": 3 + 2 * n + n_mid, "B_g2": n + 2}}
    if not quiet and rank == 0:
        print(json.dumps(rec), flush=True)
    _lib.check(zk.zk_key_free(h))
    dom.free()
    if not ok:
        raise AssertionError("Groth16 proof differs from the closed form (2^%d, %s)" % (logn, circuit))
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, nargs="*", default=[16])
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--circuit", nargs="*", default=["mulchain"], choices=sorted(CIRCUITS))
    args = ap.parse_args()
    dist = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        lr = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(lr)
        os.environ.setdefault("ZKB200_DEVICE", str(lr))
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    zk = _lib.lib()
    for ln in args.logn:
        for c in args.circuit:
            run(zk, ln, args.iters, c, dist)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
