#!/usr/bin/env python3
"""GPU probe: integer-pipe microbenchmarks and G1/G2 table-MSM timings with exact
known-dlog verification.  Prints one JSON object per line.  (Development tool; the
contract benchmark is bench.py.)"""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from zukelang_b200 import _lib

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def rand_scalars(n, seed):
    """n uniform scalars mod r as (n, 4) uint64 little-endian words + python ints on demand."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    w[:, 3] &= np.uint64((1 << 62) - 1)      # < 2^254 < r : uniform enough for timing, canonical
    return w


def words_to_ints(w):
    return [int(a) | (int(b) << 64) | (int(c) << 128) | (int(d) << 192) for a, b, c, d in w]


def microbench(zk):
    names = {0: "mad.lo.u32", 1: "mad.cc chain", 2: "mad.wide.u32", 3: "Fp mul", 4: "Fr mul", 5: "G1 madd", 6: "G1 madd call/4", 7: "G1 madd inline/4", 8: "G1 madd call/5", 9: "G1 madd call/3", 10: "G1 madd paired/4", 11: "G1 madd paired/3", 12: "G1 madd paired/2"}
    iters = {0: 4096, 1: 4096, 2: 4096, 3: 256, 4: 512, 5: 64, 6: 64, 7: 64, 8: 64, 9: 64, 10: 64, 11: 64, 12: 64}
    for kind in range(13):
        ops, ms = ctypes.c_double(), ctypes.c_double()
        best = 0.0
        for _ in range(3):
            _lib.check(zk.zk_bench_intpipe(kind, iters[kind], ctypes.byref(ops), ctypes.byref(ms)))
            best = max(best, ops.value)
        print(json.dumps({"probe": "intpipe", "kind": names[kind], "ops_per_s": best, "ms": ms.value}), flush=True)


def msm_run(zk, group, logn, precompute, iters, c, seed=0x5A554B45, dist="uniform"):
    n = 1 << logn
    raw, outn = (96, 144) if group == "g1" else (192, 288)
    fixed = getattr(zk, "zk_%s_fixed_base_mul" % group)
    load = getattr(zk, "zk_%s_table_load" % group)
    msm_dev = getattr(zk, "zk_%s_table_msm_dev" % group)
    dl_w = rand_scalars(n, 0x42415345)
    t0 = time.time()
    bases = np.empty(n * raw, dtype=np.uint8)
    _lib.check(fixed(dl_w.ctypes.data, n, bases.ctypes.data))
    t_fixed = time.time() - t0
    h = ctypes.c_uint64()
    t0 = time.time()
    _lib.check(load(bases.ctypes.data, None, n, precompute, c, ctypes.byref(h)))
    t_load = time.time() - t0
    info = (ctypes.c_uint64 * 8)()
    _lib.check(zk.zk_table_info(h.value, info))
    sc_w = rand_scalars(n, seed)
    if dist == "witness":          # 90 % of the scalars in {0, 1}, 10 % uniform
        rng = np.random.Generator(np.random.PCG64(seed + 1))
        small = rng.random(n) < 0.9
        bits = rng.integers(0, 2, size=n, dtype=np.uint64)
        sc_w[small] = 0
        sc_w[small, 0] = bits[small]
    d_sc = torch.from_numpy(sc_w.view(np.int64)).cuda()
    d_out = torch.zeros(outn, dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    times = []
    for it in range(iters + 1):
      with torch.cuda.stream(side):
        stream = side.cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        _lib.check(msm_dev(h.value, d_sc.data_ptr(), n, d_out.data_ptr(), stream))
        e1.record(side)
        torch.cuda.synchronize()
        if it:
            times.append(e0.elapsed_time(e1))
    # stage split of one more run (digits+sort, accumulate+fix-up, bucket reduction, finalize)
    stage = (ctypes.c_float * 4)()
    _lib.check(zk.zk_table_profile(h.value, 1, None))
    with torch.cuda.stream(side):
        _lib.check(msm_dev(h.value, d_sc.data_ptr(), n, d_out.data_ptr(), side.cuda_stream))
    torch.cuda.synchronize()
    _lib.check(zk.zk_table_profile(h.value, 0, stage))
    got = bytes(d_out.cpu().numpy())
    # exact check: sum s_i d_i mod r times the generator, via a 1-point MSM on the device
    tot = sum(a * b for a, b in zip(words_to_ints(sc_w), words_to_ints(dl_w))) % R
    exp = np.empty(raw, dtype=np.uint8)
    _lib.check(fixed(tot.to_bytes(32, "little"), 1, exp.ctypes.data))
    ok = got[:raw] == bytes(exp)
    _lib.check(zk.zk_table_free(h.value))
    ms = min(times)
    print(json.dumps({"probe": "msm", "group": group, "log_n": logn, "precompute": precompute, "c": int(info[0]),
                      "W": int(info[1]), "S": int(info[4]), "table_MB": int(info[5]) >> 20, "ms": ms,
                      "ms_all": times, "stages_ms": [round(float(x), 4) for x in stage],
                      "env": {k: v for k, v in os.environ.items() if k.startswith("ZKB200_") and k != "ZKB200_DEVICE"}, "Mpts_per_s": n / ms / 1e3, "exact_ok": bool(ok), "scalars": dist,
                      "fixed_base_s": t_fixed, "load_s": t_load}), flush=True)
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--micro", action="store_true")
    ap.add_argument("--group", default="g1")
    ap.add_argument("--logn", type=int, nargs="*", default=[])
    ap.add_argument("--precompute", type=int, nargs="*", default=[0, 1])
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--c", type=int, default=0)
    ap.add_argument("--dist", default="uniform", choices=["uniform", "witness"])
    args = ap.parse_args()
    zk = _lib.lib()
    info = ctypes.create_string_buffer(256)
    _lib.check(zk.zk_device_info(info, 256))
    print(json.dumps({"probe": "device", "info": info.value.decode()}), flush=True)
    if args.micro:
        microbench(zk)
    ok = True
    for logn in args.logn:
        for pre in args.precompute:
            ok &= msm_run(zk, args.group, logn, pre, args.iters, args.c, dist=args.dist)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
