#!/usr/bin/env python3
"""Groth16 prove throughput on synthetic multiply-chain circuits (BASELINE configs[2] / [4]):
c_0 = x, c_{i+1} = c_i * x, n gates, through zk_groth16_prove_r1cs (evaluation-form prover).
Every run is checked exactly against the closed-form trapdoor identity.  One JSON line per size."""
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


def fixed_base(zk, group, scalars):
    raw = 96 if group == "g1" else 192
    out = (ctypes.c_uint8 * (raw * len(scalars)))()
    _lib.check(getattr(zk, "zk_%s_fixed_base_mul" % group)(fr_vector(scalars), len(scalars), out))
    return bytes(out)


def run(zk, logn, iters, shard=(0, 1), quiet=False):
    n = 1 << logn
    t0 = time.time()
    circ, witness = mulchain(n)
    dom = S.EvalDomain(circ)
    P = S.Groth16Sparse()
    rng = random.Random(0x47524F54 + logn)
    td = tuple(rng.randrange(R) for _ in range(5))
    a, b, gm, d, t = td
    sc = P.keygen_scalars(td, circ, dom.w)
    mids = sc["mids"]
    g1 = fixed_base(zk, "g1", [a, b, d] + sc["lag"] + sc["hk"] + [sc["ltd"][k] for k in mids])
    g2 = fixed_base(zk, "g2", [b, d] + sc["lag"])
    pos = {k: i for i, k in enumerate(circ.variables)}
    idx = (ctypes.c_uint32 * len(mids))(*[pos[k] for k in mids])
    o1 = lambda i, cnt: ctypes.create_string_buffer(g1[i * 96:(i + cnt) * 96], cnt * 96)
    o2 = lambda i, cnt: ctypes.create_string_buffer(g2[i * 192:(i + cnt) * 192], cnt * 192)
    bufs = dict(a=o1(0, 1), b1=o1(1, 1), d1=o1(2, 1), ti1=o1(3, n), tiztd=o1(3 + n, n), ltd_mid=o1(3 + 2 * n, len(mids)),
                b2=o2(0, 1), d2=o2(1, 1), ti2=o2(2, n))
    st = _lib.Groth16PKeyStruct(n=n, m=len(circ.variables), n_mid=len(mids), n_h=n, mid_index=ctypes.addressof(idx),
                                **{k: ctypes.addressof(v) for k, v in bufs.items()})
    h = ctypes.c_uint64()
    _lib.check(zk.zk_groth16_pk_load(ctypes.byref(st), shard[0], shard[1], ctypes.byref(h)))
    setup_s = time.time() - t0
    sols = []
    for i in range(2):
        sol = witness(rng.randrange(R))
        sols.append((sol, fr_vector(sol[k] for k in circ.variables)))
    out = (ctypes.c_uint8 * _lib.GROTH16_PROOF_OUT)()
    times = []
    ok = True
    for it in range(iters + 1):
        sol, sol_b = sols[it & 1]
        r, s = rng.randrange(R), rng.randrange(R)
        t1 = time.perf_counter()
        _lib.check(zk.zk_groth16_prove_r1cs(h.value, dom.handle, sol_b, r.to_bytes(32, "little"), s.to_bytes(32, "little"), out))
        dt = time.perf_counter() - t1
        if it:
            times.append(dt)
        if it <= 1 and shard == (0, 1):
            A, B, C = S.closed_form_scalars(td, r, s, circ, dom.w, sol)
            exp = fixed_base(zk, "g1", [A, C])
            expb = fixed_base(zk, "g2", [B])
            b_ = bytes(out)
            ok &= b_[0:96] == exp[:96] and b_[432:528] == exp[96:] and b_[144:336] == expb
    best = min(times)
    rec = {"probe": "groth16", "log_n": logn, "constraints": n, "variables": len(circ.variables),
           "prove_ms_best": best * 1e3, "prove_ms_all": [x * 1e3 for x in times], "proofs_per_s": 1.0 / best,
           "exact_ok": bool(ok), "setup_s": setup_s, "circuit": "multiply chain c_{i+1} = c_i * x",
           "h2d_bytes_per_proof": 32 * len(circ.variables) + 64, "d2h_bytes_per_proof": 576,
           "msm_points": {"A_g1": n + 3, "C_g1": 3 + 2 * n + len(mids), "B_g2": n + 2}}
    if not quiet:
        print(json.dumps(rec), flush=True)
    _lib.check(zk.zk_key_free(h.value))
    dom.free()
    if quiet:
        if not ok:
            raise AssertionError('Groth16 proof differs from the closed form')
        return rec
    return ok


def run_sharded(zk, logn, iters):
    """BASELINE configs[4]: one proof sharded by base range over the ranks of a torchrun job.
    Every rank evaluates the QAP redundantly and proves with shard (rank, world) of the key; the
    576-byte partial results are all-gathered and added (zk_g1_sum / zk_g2_sum).  Rank 0 checks the
    proof against the closed form and prints one JSON line."""
    import torch
    import torch.distributed as dist
    from zukelang_b200 import dist as D
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    n = 1 << logn
    circ, witness = mulchain(n)
    dom = S.EvalDomain(circ)
    P = S.Groth16Sparse()
    rng = random.Random(0x47524F54 + logn)
    td = tuple(rng.randrange(R) for _ in range(5))
    a, b, gm, d, t = td
    sc = P.keygen_scalars(td, circ, dom.w)
    mids = sc["mids"]
    g1 = fixed_base(zk, "g1", [a, b, d] + sc["lag"] + sc["hk"] + [sc["ltd"][k] for k in mids])
    g2 = fixed_base(zk, "g2", [b, d] + sc["lag"])
    pos = {k: i for i, k in enumerate(circ.variables)}
    idx = (ctypes.c_uint32 * len(mids))(*[pos[k] for k in mids])
    o1 = lambda i, cnt: ctypes.create_string_buffer(g1[i * 96:(i + cnt) * 96], cnt * 96)
    o2 = lambda i, cnt: ctypes.create_string_buffer(g2[i * 192:(i + cnt) * 192], cnt * 192)
    bufs = dict(a=o1(0, 1), b1=o1(1, 1), d1=o1(2, 1), ti1=o1(3, n), tiztd=o1(3 + n, n), ltd_mid=o1(3 + 2 * n, len(mids)),
                b2=o2(0, 1), d2=o2(1, 1), ti2=o2(2, n))
    st = _lib.Groth16PKeyStruct(n=n, m=len(circ.variables), n_mid=len(mids), n_h=n, mid_index=ctypes.addressof(idx),
                                **{k: ctypes.addressof(v) for k, v in bufs.items()})
    h = ctypes.c_uint64()
    _lib.check(zk.zk_groth16_pk_load(ctypes.byref(st), rank, world, ctypes.byref(h)))
    del g1, g2, bufs
    sol = witness(rng.randrange(R))
    sol_b = fr_vector(sol[k] for k in circ.variables)
    out = (ctypes.c_uint8 * _lib.GROTH16_PROOF_OUT)()
    times, ok = [], True
    for it in range(iters + 1):
        r, s = rng.randrange(R), rng.randrange(R)
        dist.barrier()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        _lib.check(zk.zk_groth16_prove_r1cs(h.value, dom.handle, sol_b, r.to_bytes(32, "little"), s.to_bytes(32, "little"), out))
        parts = D.all_gather_bytes(bytes(out))
        proof = D.combine_groth16(parts)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if it:
            times.append(float(dt.item()))
        if it <= 1 and rank == 0:
            A, B, C = S.closed_form_scalars(td, r, s, circ, dom.w, sol)
            exp = fixed_base(zk, "g1", [A, C])
            expb = fixed_base(zk, "g2", [B])
            ok &= proof[0:96] == exp[:96] and proof[432:528] == exp[96:] and proof[144:336] == expb
    if rank == 0:
        best = min(times)
        print(json.dumps({"probe": "groth16_sharded", "n_gpus": world, "log_n": logn, "constraints": n,
                          "prove_ms_best": best * 1e3, "prove_ms_all": [x * 1e3 for x in times],
                          "proofs_per_s": 1.0 / best, "exact_ok": bool(ok),
                          "timing": "wall clock around prove + gather + combine, max over ranks"}), flush=True)
    _lib.check(zk.zk_key_free(h.value))
    dom.free()
    return ok


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, nargs="*", default=[16])
    ap.add_argument("--iters", type=int, default=4)
    args = ap.parse_args()
    ok = True
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
            ok &= run_sharded(zk, ln, args.iters)
        dist.barrier()
        dist.destroy_process_group()
    else:
        zk = _lib.lib()
        for ln in args.logn:
            ok &= run(zk, ln, args.iters)
    sys.exit(0 if ok else 1)
