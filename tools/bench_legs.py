"""Secondary legs of bench.py (kept out of bench.py so that the headline path stays readable):

  sweep       BASELINE configs[3]: G1 MSM at 2^16 .. 2^24 points on the N ranks of the job
  oneshot     the non-precomputed shapes: zk_g1_msm (host bases + scalars in, what Curve.G.dot binds
              to) and a resident table without precomputed windows
  g2          a 2^20-point G2 table MSM with its own roofline object (134 400 MAC32 per point)

Every MSM result is checked exactly against its known discrete log.  Product-side code: nothing
here touches oracle/."""
from __future__ import annotations

import ctypes
import time

import numpy as np

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
MAC32_PER_POINT_G2 = 134_400       # 16 windows x 8 400 MAC32 (G2 XYZZ mixed add)
SWEEP_A, SWEEP_B = 0x5A554B4531, 0x42415345AB   # bases P_i = (A i + B) G, discrete logs below 2^64


def _sum_words(w):
    """sum_i s_i for scalars given as (n, 4) little-endian uint64 words, exact."""
    tot = 0
    for k in range(4):
        col = w[:, k]
        lo = int((col & np.uint64(0xFFFFFFFF)).sum(dtype=np.uint64))
        hi = int((col >> np.uint64(32)).sum(dtype=np.uint64))
        tot += (lo + (hi << 32)) << (64 * k)
    return tot


def _sum_index_words(idx, w):
    """sum_i idx_i * s_i, exact (idx < 2^24, so 16-bit pieces of the words keep every partial sum below 2^64)."""
    tot = 0
    for k in range(4):
        col = w[:, k]
        for p in range(4):
            piece = (col >> np.uint64(16 * p)) & np.uint64(0xFFFF)
            tot += int((piece * idx).sum(dtype=np.uint64)) << (64 * k + 16 * p)
    return tot


def masked_scalars(n, seed):
    """n scalars uniform in [0, 2^254) (canonical: 2^254 < r) as (n, 4) uint64 words — the sweep's
    scalars; the headline step uses uniform_scalars (reduced mod r) of bench.py."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    w[:, 3] &= np.uint64((1 << 62) - 1)
    return w


def _fixed_base(zk, lib, group, words):
    raw = 96 if group == "g1" else 192
    n = words.shape[0]
    out = np.empty(n * raw, dtype=np.uint8)
    lib.check(getattr(zk, "zk_%s_fixed_base_mul" % group)(words.ctypes.data, n, out.ctypes.data))
    return out


def _expect(zk, lib, group, dlog):
    raw = 96 if group == "g1" else 192
    exp = np.empty(raw, dtype=np.uint8)
    lib.check(getattr(zk, "zk_%s_fixed_base_mul" % group)((dlog % R).to_bytes(32, "little"), 1, exp.ctypes.data))
    return bytes(exp)


def sweep(zk, lib, torch, dist, rank, world, side, logns, seed):
    """One record per size: this rank reduces its base range, the partial sums are all-gathered and
    added (N > 1).  ms = CUDA events on the launching stream, max over ranks: `single` = one MSM at
    a time, `pipelined` = 8 MSMs queued behind one batched tail (what the headline step does)."""
    out = []
    Q = 8
    for logn in logns:
        n_total = 1 << logn
        lo, hi = n_total * rank // world, n_total * (rank + 1) // world
        n = hi - lo
        idx = np.arange(lo, hi, dtype=np.uint64)
        dl = np.zeros((n, 4), dtype=np.uint64)
        dl[:, 0] = idx * np.uint64(SWEEP_A) + np.uint64(SWEEP_B)        # < 2^64 for i < 2^24
        t0 = time.time()
        bases = _fixed_base(zk, lib, "g1", dl)
        h = ctypes.c_uint64()
        lib.check(zk.zk_g1_table_load(bases.ctypes.data, None, n, 1, 0, ctypes.byref(h)))
        del bases
        info = (ctypes.c_uint64 * 8)()
        lib.check(zk.zk_table_info(h.value, info))
        setup_s = time.time() - t0
        w = masked_scalars(n, seed + 977 * logn + rank)
        d_sc = torch.from_numpy(w.view(np.int64)).cuda()
        d_out = torch.zeros(Q, 144, dtype=torch.uint8, device="cuda")
        parts = torch.zeros(Q, 96, dtype=torch.uint8, device="cuda")
        gath = torch.zeros(world, Q, 96, dtype=torch.uint8, device="cuda")
        sums = torch.zeros(Q, 144, dtype=torch.uint8, device="cuda")

        def steps(k):
            for q in range(k):
                lib.check(zk.zk_g1_table_msm_dev(h.value, d_sc.data_ptr(), n, d_out[q].data_ptr(), side.cuda_stream))
            lib.check(zk.zk_table_join(h.value, side.cuda_stream))
            if world > 1:
                parts.copy_(d_out[:, :96])
                dist.all_gather_into_tensor(gath.view(-1), parts.view(-1))
                lib.check(zk.zk_g1_sum_strided_dev(gath.data_ptr(), world, Q, sums.data_ptr(), side.cuda_stream))

        def timed(k, reps):
            best = None
            for _ in range(reps):
                if dist is not None:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(side):
                    e0.record(side)
                    steps(k)
                    e1.record(side)
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
                if dist is not None:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                best = float(t.item()) if best is None else min(best, float(t.item()))
            return best

        with torch.cuda.stream(side):
            steps(1)                                       # warm-up + exact check of this rank's partial
        torch.cuda.synchronize()
        dlog = SWEEP_A * _sum_index_words(idx, w) + SWEEP_B * _sum_words(w)
        ok = bytes(d_out[0].cpu().numpy())[:96] == _expect(zk, lib, "g1", dlog)
        ms_single = timed(1, 3)
        lib.check(zk.zk_table_pipeline(h.value, 1))
        timed(Q, 1)
        ms_pipe = timed(Q, 2) / Q
        lib.check(zk.zk_table_pipeline(h.value, 0))
        okt = torch.tensor([1 if ok else 0], device="cuda")
        if dist is not None:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        out.append({"log_n": logn, "points": n_total, "window_bits": int(info[0]), "windows": int(info[1]),
                    "ms_single": ms_single, "mpts_single": n_total / ms_single / 1e3,
                    "ms_pipelined": ms_pipe, "mpts_pipelined": n_total / ms_pipe / 1e3,
                    "exact_ok": bool(int(okt.item())), "setup_s": round(setup_s, 2), "table_MiB_per_gpu": int(info[5]) >> 20})
        lib.check(zk.zk_table_free(h.value))
        del d_sc
    return out


def oneshot(zk, lib, torch, side, logn, seed, peak_mac32):
    """N = 1: the shapes without a precomputed table."""
    n = 1 << logn
    idx = np.arange(n, dtype=np.uint64)
    dl = np.zeros((n, 4), dtype=np.uint64)
    dl[:, 0] = idx * np.uint64(SWEEP_A) + np.uint64(SWEEP_B)
    bases = _fixed_base(zk, lib, "g1", dl)
    w = masked_scalars(n, seed + 31)
    exp = _expect(zk, lib, "g1", SWEEP_A * _sum_index_words(idx, w) + SWEEP_B * _sum_words(w))
    rec = {"log_n": logn}
    # (1) zk_g1_msm: host bases and scalars in (parse + curve check + upload + MSM), point out
    outb = (ctypes.c_uint8 * 144)()
    times = []
    for it in range(3):
        t0 = time.perf_counter()
        lib.check(zk.zk_g1_msm(bases.ctypes.data, None, w.ctypes.data, n, outb))
        times.append((time.perf_counter() - t0) * 1e3)
    rec["zk_g1_msm"] = {"ms": min(times[1:]), "mpts": n / min(times[1:]) / 1e3, "exact_ok": bytes(outb)[:96] == exp,
                        "h2d_bytes": n * 128, "note": "one call per MSM: bases parsed, curve-checked and uploaded every time (what Curve.G.dot binds to)"}
    # (2) resident table without precomputed windows, device-resident scalars
    h = ctypes.c_uint64()
    lib.check(zk.zk_g1_table_load(bases.ctypes.data, None, n, 0, 0, ctypes.byref(h)))
    info = (ctypes.c_uint64 * 8)()
    lib.check(zk.zk_table_info(h.value, info))
    d_sc = torch.from_numpy(w.view(np.int64)).cuda()
    d_out = torch.zeros(144, dtype=torch.uint8, device="cuda")
    stage = (ctypes.c_float * 4)()
    ms = []
    lib.check(zk.zk_table_profile(h.value, 1, None))
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record(side)
            lib.check(zk.zk_g1_table_msm_dev(h.value, d_sc.data_ptr(), n, d_out.data_ptr(), side.cuda_stream))
            e1.record(side)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    lib.check(zk.zk_table_profile(h.value, 0, stage))
    best = min(ms[1:])
    W = int(info[1])
    rec["table_no_precompute"] = {"ms": best, "mpts": n / best / 1e3, "window_bits": int(info[0]), "windows": W,
                                  "exact_ok": bytes(d_out.cpu().numpy())[:96] == exp, "stages_ms": [float(x) for x in stage],
                                  "roofline_frac_whole_msm": n * 48_000 / (best * 1e-3) / peak_mac32,
                                  "roofline_frac_accumulate": n * 48_000 / (float(stage[1]) * 1e-3) / peak_mac32}
    lib.check(zk.zk_table_free(h.value))
    return rec


def g2_roofline(zk, lib, torch, side, logn, seed, peak_mac32):
    """N = 1: a 2^logn-point G2 MSM over a precomputed table; roofline object for k_accumulate<Fp2>."""
    n = 1 << logn
    idx = np.arange(n, dtype=np.uint64)
    dl = np.zeros((n, 4), dtype=np.uint64)
    dl[:, 0] = idx * np.uint64(SWEEP_A) + np.uint64(SWEEP_B)
    bases = _fixed_base(zk, lib, "g2", dl)
    h = ctypes.c_uint64()
    lib.check(zk.zk_g2_table_load(bases.ctypes.data, None, n, 1, 0, ctypes.byref(h)))
    del bases
    info = (ctypes.c_uint64 * 8)()
    lib.check(zk.zk_table_info(h.value, info))
    w = masked_scalars(n, seed + 63)
    exp = _expect(zk, lib, "g2", SWEEP_A * _sum_index_words(idx, w) + SWEEP_B * _sum_words(w))
    d_sc = torch.from_numpy(w.view(np.int64)).cuda()
    d_out = torch.zeros(288, dtype=torch.uint8, device="cuda")
    stage = (ctypes.c_float * 4)()
    ms, acc = [], []
    lib.check(zk.zk_table_profile(h.value, 1, None))
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record(side)
            lib.check(zk.zk_g2_table_msm_dev(h.value, d_sc.data_ptr(), n, d_out.data_ptr(), side.cuda_stream))
            e1.record(side)
        torch.cuda.synchronize()
        lib.check(zk.zk_table_profile(h.value, 1, stage))
        if it:
            ms.append(e0.elapsed_time(e1))
            acc.append(float(stage[1]))
    lib.check(zk.zk_table_profile(h.value, 0, None))
    lib.check(zk.zk_table_free(h.value))
    a = sum(acc) / len(acc)
    achieved = n * MAC32_PER_POINT_G2 / (a * 1e-3) / 1e12
    return {"bound": "int32-imad", "kernel": "k_accumulate<Fp2>", "log_n": logn, "achieved": achieved, "peak": peak_mac32 / 1e12,
            "unit": "TMAC32/s", "frac": achieved / (peak_mac32 / 1e12), "kernel_ms": a, "msm_ms": min(ms),
            "mpts": n / min(ms) / 1e3, "algorithmic_mac32_per_point": MAC32_PER_POINT_G2, "window_bits": int(info[0]),
            "windows": int(info[1]), "exact_ok": bytes(d_out.cpu().numpy())[:192] == exp, "traffic": None,
            "table_MiB": int(info[5]) >> 20}
