#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200 prover hot path.

Metric (BASELINE.json): BLS12-381 G1 MSM throughput, Mpts/s, at 2^20 points.
A "step" is ONE multi-scalar multiplication of 2^20 uniformly random Fr scalars against a
resident proving-key-style base table (scalars from PCG64(0x5A554B45),
bases P_i = d_i * G with known discrete logs so the result is checked exactly every run).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1: one process.  N > 1: launched by torchrun, one rank per GPU; the 2^20 points are
sharded by base range across ranks ("strong" scaling), each rank reduces its
shard, the N partial sums (96 B each) are all-gathered over NCCL and summed on every rank.

Fields beyond the base contract:
  roofline      dominant kernel (k_accumulate) against the integer-multiply pipe: algorithmic
                MAC32 per launch (48 000 per point) / CUDA-event duration of that
                kernel in the timed steps / the IMAD-chain peak measured in this same process.
  cpu_baseline  the oracle's C restatement of the reference's fold-MSM on the host cores
                (bounded sample), rank 0 / N = 1 only.
  cpu_pippenger informational "fair CPU" line: a multi-threaded bucket method in
                portable C on the same host cores, checked against the fold; NOT the reference's
                algorithm and not the baseline.
  e2e           same metric through the host-buffer C-ABI call (pinned host scalars in, point out).
--impl reference runs only the CPU restatement (the reference itself needs OCaml, absent here).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
MAC32_PER_POINT = 48_000          # 16 windows x 3 000 MAC32 (G1 XYZZ mixed add)
LOG_N = 20
NCU_DRAM_BYTES_PER_MSM = 3_052_184_568      # k_accumulate, one 2^20-point MSM, c = 17 (profiles/r02_ncu_accumulate.md)
SEED_SCALARS, SEED_BASES = 0x5A554B45, 0x42415345


def words_to_ints(w):
    return [int(a) | (int(b) << 64) | (int(c) << 128) | (int(d) << 192) for a, b, c, d in w]


def uniform_scalars(n, seed):
    """n scalars uniform in [0, r): four u64 words from PCG64(seed), reduced mod r.
    Returns (uint64 array (n, 4) little-endian words, list of python ints)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    w = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    ints = [v % R for v in words_to_ints(w.tolist())]
    out = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in ints), dtype=np.uint64).reshape(n, 4).copy()
    return out, ints


# ------------------------------------------------------------------------------------------
# host placement
# ------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs NVML reports as local to the GPU, BEFORE any pinned host memory
    is allocated (first-touch places it on that NUMA node; a remote node costs ~4x H2D bandwidth)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "cpus %d-%d (%d)" % (min(cpus), max(cpus), len(cpus))
    except Exception as e:
        return "unchanged (%s)" % type(e).__name__
    return "unchanged"


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows, self.proc, self.idx = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_start=None, t_end=None):
        """Summary of the samples that arrived inside [t_start, t_end] (the timed region); falls back
        to the busiest half of all samples when the region was shorter than the sampling period."""
        if self.proc:
            self.proc.terminate()
        inside = [r for ts, r in self.rows if t_start is not None and t_start - 0.05 <= ts <= t_end + 0.15]
        rows = inside if inside else [r for _, r in self.rows]
        sm, mx, reasons = [], 0, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        load = sm if inside else (sm[len(sm) // 2:] if sm else [])
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(inside)}


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ------------------------------------------------------------------------------------------
def load_c_oracle():
    so = os.path.join(ROOT, "oracle", "c", "libzkoracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "c")], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(so)
    lib.zkoracle_g1_msm_fold.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    lib.zkoracle_g1_msm_fold.restype = ctypes.c_int
    if hasattr(lib, "zkoracle_g1_msm_pippenger"):         # absent from a stale build: only cpu_pippenger is lost
        lib.zkoracle_g1_msm_pippenger.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_void_p]
        lib.zkoracle_g1_msm_pippenger.restype = ctypes.c_int
    return lib


def cpu_fold_msm(bases_raw, scalars_raw, n, threads):
    """The oracle's restatement of curve.ml:91-118 on `threads` host cores; returns (seconds, out96)."""
    lib = load_c_oracle()
    if n * 96 > len(bases_raw) or n * 32 > len(scalars_raw):
        raise ValueError("cpu_fold_msm: %d points asked of buffers holding %d" % (n, len(bases_raw) // 96))
    out = (ctypes.c_uint8 * 96)()
    t0 = time.perf_counter()
    lib.zkoracle_g1_msm_fold(bases_raw, scalars_raw, n, threads, out)
    return time.perf_counter() - t0, bytes(out)


def cpu_pippenger_msm(bases_raw, scalars_raw, n, threads, c=16):
    """NOT the reference's algorithm: a plain multi-threaded bucket method in portable C (no assembly),
    the "fair CPU" line of SURVEY.md §8(d); returns (seconds, out96)."""
    lib = load_c_oracle()
    out = (ctypes.c_uint8 * 96)()
    t0 = time.perf_counter()
    rc = lib.zkoracle_g1_msm_pippenger(bases_raw, scalars_raw, n, c, threads, out)
    if rc != 0:
        raise RuntimeError("zkoracle_g1_msm_pippenger: bad arguments")
    return time.perf_counter() - t0, bytes(out)


def cpu_pippenger_line(bases_raw, scalars_raw, n_check, fold_out, n_timed, threads):
    """The informational ``cpu_pippenger`` object of the JSON line: checked against the fold on the
    cpu_baseline sample, timed on a larger prefix of the same workload."""
    if max(n_check, n_timed) * 96 > len(bases_raw) or max(n_check, n_timed) * 32 > len(scalars_raw):
        raise ValueError("cpu_pippenger_line: %d / %d points asked of buffers holding %d" % (n_check, n_timed, len(bases_raw) // 96))
    _, chk = cpu_pippenger_msm(bases_raw[:n_check * 96], scalars_raw[:n_check * 32], n_check, threads)
    secs, _ = cpu_pippenger_msm(bases_raw[:n_timed * 96], scalars_raw[:n_timed * 32], n_timed, threads)
    return {"value": n_timed / secs / 1e6, "unit": "Mpts/s", "cores": threads, "window_bits": 16,
            "agrees_with_fold": chk == fold_out,
            "sample": "first %d of the 2^20 points, oracle/c bucket method (signed 16-bit windows, portable C, "
                      "no assembly) on all host threads, %.1f s; not the reference's algorithm" % (n_timed, secs)}


def oracle_bases(n, seed):
    """n bases d_i * G on the CPU (python oracle, small n only) -> raw bytes, dlogs."""
    from oracle import bls12_381 as O
    import random
    rng = random.Random(seed)
    dl = [rng.randrange(1, 1 << 40) for _ in range(n)]
    # one running addition chain keeps this O(n) group operations
    step = O.G1.mul(O.G1.one, 1 << 20)
    pts, cur, acc_d = [], O.G1.mul(O.G1.one, dl[0]), dl[0]
    dl_out = []
    for i in range(n):
        pts.append(cur)
        dl_out.append(acc_d)
        cur = O.G1.add(cur, step)
        acc_d += 1 << 20
    return b"".join(O.g1_to_uncompressed(p) for p in pts), dl_out


def run_reference_arm(args):
    """--impl reference: the reference's own algorithm on the host cores (oracle port; the OCaml
    reference cannot be built in this image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    threads = os.cpu_count() or 1
    sample = 1 << 12                              # bounded sample of the 2^20 workload per step
    bases_raw, dl = oracle_bases(sample, SEED_BASES)
    _, ints = uniform_scalars(sample, SEED_SCALARS)
    sc_raw = b"".join(v.to_bytes(32, "little") for v in ints)
    from oracle import bls12_381 as O
    expect = O.g1_to_uncompressed(O.G1.mul(O.G1.one, sum(a * b for a, b in zip(ints, dl)) % R))
    times = []
    for it in range(args.warmup + args.steps):
        dt, out = cpu_fold_msm(bases_raw, sc_raw, sample, threads)
        assert out == expect, "CPU restatement disagrees with the closed form"
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    val = sample * len(times) / total / 1e6
    line = {"impl": "reference", "metric": "bls12_381_g1_msm_throughput", "value": val, "unit": "Mpts/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32-limb modular (Fp 381-bit)",
            "data": "synthetic", "config": {"workload": "G1 MSM, 2^20 points, uniform 255-bit scalars (config 4)",
                                            "sample": "first 2^12 points per step", "algorithm": "reference fold of double-and-add scalar muls (curve.ml:91-118)"},
            "cpu_baseline": {"value": val, "unit": "Mpts/s", "cores": threads, "kind": "port",
                             "sample": "2^12 of the 2^20 points per step, oracle/c fold-MSM on all host threads"},
            "e2e": {"value": val, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def tail_tree_levels(c, bucket_windows, q_join, sm_count):
    """Launches of k_reduce_tree in one batched tail: the chunk width L of the bucket reduction as
    csrc/msm_impl.cuh:tail() picks it (the narrowest of 4..32 that keeps the chunk threads of the
    q_join queued MSMs within one warp per scheduler), then 64-wide trees down to one element."""
    buckets = 1 << (c - 1)
    L = 4
    while L < 32 and q_join * buckets * bucket_windows // L > sm_count * 4 * 32:
        L <<= 1
    L = min(L, buckets)
    levels, cnt = 0, buckets // L
    while True:
        levels += 1
        cnt = (cnt + 63) // 64
        if cnt <= 1:
            break
    return levels


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import numpy as np
    import torch
    from zukelang_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
    affinity = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("ZKB200_DEVICE", str(local_rank))
    dist = None
    # keep stdout to the single JSON line: NCCL's "NCCL version ..." banner goes to stdout
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", ""):
        os.environ["NCCL_DEBUG"] = "WARN"
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=600))
    zk = _lib.lib()
    sampler = ClockSampler(local_rank)
    sampler.start()                                     # nvidia-smi takes a second to come up: start it first
    n_total = 1 << args.logn
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    n = hi - lo

    # ---- setup (untimed): bases with known discrete logs, resident precomputed table ----------
    t0 = time.time()
    dl_w_all, dl_all = uniform_scalars(n_total, SEED_BASES)
    dl_w = np.ascontiguousarray(dl_w_all[lo:hi])
    bases = np.empty(n * 96, dtype=np.uint8)
    _lib.check(zk.zk_g1_fixed_base_mul(dl_w.ctypes.data, n, bases.ctypes.data))
    handle = ctypes.c_uint64()
    _lib.check(zk.zk_g1_table_load(bases.ctypes.data, None, n, 1, args.window_bits, ctypes.byref(handle)))
    info = (ctypes.c_uint64 * 8)()
    _lib.check(zk.zk_table_info(handle.value, info))
    setup_s = time.time() - t0

    # scalar batches: distinct per step (rotating pool), resident in HBM for `value`,
    # in pinned host memory for `e2e`
    pool = min(4, args.steps + args.warmup)
    batches = []
    for b in range(pool):
        w, ints = uniform_scalars(n_total, SEED_SCALARS + b)
        tot = sum(a * d for a, d in zip(ints, dl_all)) % R
        host = torch.from_numpy(np.ascontiguousarray(w[lo:hi]).view(np.int64)).pin_memory()
        batches.append({"host": host, "dev": host.cuda(), "expect_dlog": tot})
    QUEUE = int(os.environ.get("ZKB200_BENCH_QUEUE", "32"))   # zk_table_pipeline batches up to 32 tails per join
    d_ring = torch.zeros(QUEUE, 144, dtype=torch.uint8, device="cuda")    # one result slot per queued MSM
    p_ring = torch.zeros(QUEUE, 96, dtype=torch.uint8, device="cuda")     # partial sums to gather (N > 1)
    g_ring = torch.zeros(world, QUEUE, 96, dtype=torch.uint8, device="cuda")
    s_ring = torch.zeros(QUEUE, 144, dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    pipelined = not args.no_pipeline

    def run_steps(batch_ids):
        """One MSM per entry of batch_ids, device-resident scalars, on stream `side`.
        Pipelined: up to QUEUE MSMs are queued and zk_table_join runs the group as ONE launch
        sequence (one counting sort and one balanced accumulation over all their buckets, one batched
        tail); for N > 1 the group's partial sums travel in one all_gather and are added by one
        batched kernel.  Returns the last result tensor."""
        group = QUEUE if pipelined else 1
        last = None
        for g0 in range(0, len(batch_ids), group):
            ids = batch_ids[g0:g0 + group]
            for q, b in enumerate(ids):
                _lib.check(zk.zk_g1_table_msm_dev(handle.value, batches[b]["dev"].data_ptr(), n, d_ring[q].data_ptr(),
                                                  side.cuda_stream))
            _lib.check(zk.zk_table_join(handle.value, side.cuda_stream))
            last = d_ring[len(ids) - 1]
            if world > 1:
                p_ring.copy_(d_ring[:, :96])
                dist.all_gather_into_tensor(g_ring.view(-1), p_ring.view(-1))
                _lib.check(zk.zk_g1_sum_strided_dev(g_ring.data_ptr(), world, QUEUE, s_ring.data_ptr(), side.cuda_stream))
                last = s_ring[len(ids) - 1]
        return last

    def combine(res):
        return bytes(res.cpu().numpy())[:96]

    def expected_point(tot):
        exp = np.empty(96, dtype=np.uint8)
        _lib.check(zk.zk_g1_fixed_base_mul(tot.to_bytes(32, "little"), 1, exp.ctypes.data))
        return bytes(exp)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- integer-pipe peak, measured in this process (denominator of the roofline) -------------
    ops, ms = ctypes.c_double(), ctypes.c_double()
    peak_imad = 0.0
    for _ in range(3):
        _lib.check(zk.zk_bench_intpipe(1, 4096, ctypes.byref(ops), ctypes.byref(ms)))
        peak_imad = max(peak_imad, ops.value)
    peak_mac32 = peak_imad / 2.0                       # one MAC32 = mad.lo + mad.hi

    # ---- warm-up, with the exact known-dlog check -------------------------------------------
    _lib.check(zk.zk_table_pipeline(handle.value, QUEUE if pipelined else 0))
    with torch.cuda.stream(side):
        for it in range(args.warmup):
            b = it % pool
            res = run_steps([b])
            side.synchronize()
            assert combine(res) == expected_point(batches[b]["expect_dlog"]), "MSM result differs from the known-dlog closed form"

    # ---- timed region: `value` (inputs resident in HBM) -----------------------------------------
    stage = (ctypes.c_float * 4)()
    slots = [(args.warmup + it) % pool for it in range(args.steps)]   # scalar batch of every step
    _lib.check(zk.zk_table_profile(handle.value, 1, None))            # stage events around every join (5 records each)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.time()
    with torch.cuda.stream(side):
        e0.record(side)
        res = run_steps(slots)
        e1.record(side)
    barrier()
    t_region1 = time.time()
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_region0, t_region1)
    assert combine(res) == expected_point(batches[slots[-1]]["expect_dlog"])
    # the dominant kernel INSIDE the timed region: summed launch durations of k_accumulate (CUDA events
    # on the launching stream) and the MSMs those launches processed
    totals, counts = (ctypes.c_float * 4)(), (ctypes.c_uint64 * 2)()
    _lib.check(zk.zk_table_profile_totals(handle.value, totals, counts))
    _lib.check(zk.zk_table_profile(handle.value, 0, None))
    timed_stage_ms = [float(x) for x in totals]
    timed_msms, timed_joins = int(counts[0]), int(counts[1])
    # stage times of one step on its own (single-MSM launches), for reference
    acc_ms, stages_last = [], None
    _lib.check(zk.zk_table_profile(handle.value, 1, None))
    with torch.cuda.stream(side):
        for it in range(min(args.steps, 5)):
            run_steps([it % pool])
            side.synchronize()
            _lib.check(zk.zk_table_profile(handle.value, 1, stage))
            acc_ms.append(float(stage[1]))
            stages_last = [float(x) for x in stage]
    _lib.check(zk.zk_table_profile(handle.value, 0, None))
    _lib.check(zk.zk_table_pipeline(handle.value, 0))
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = n_total * args.steps / (dev_ms * 1e-3) / 1e6

    # ---- e2e: host buffers through the public C-ABI call (H2D + MSM + D2H per step) ---------------
    # context for e2e: this box's pinned host -> device bandwidth (e2e cannot beat n*32 B / this)
    probe = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    dprobe = torch.empty_like(probe, device="cuda")
    dprobe.copy_(probe, non_blocking=True)
    torch.cuda.synchronize()
    tp = time.perf_counter()
    dprobe.copy_(probe, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = probe.numel() / (time.perf_counter() - tp) / 1e9
    del probe, dprobe
    # one call for all K steps: host scalar vectors in (pinned), K point results out; inside, the
    # upload of step i+1 overlaps the accumulation of step i (zk_g1_table_msm_batch).  The call is
    # warmed once (staging buffers, copy stream and events live in the table handle and are created
    # on first use), then the K-step region is repeated E2E_REPS times; the MEDIAN repetition is
    # reported, with the minimum and the per-step upload / stall / compute split of the last one.
    E2E_REPS = 3
    outs_host = torch.zeros(args.steps, 144, dtype=torch.uint8).pin_memory()
    ptrs = (ctypes.c_void_p * args.steps)(*[batches[b]["host"].data_ptr() for b in slots])
    _lib.check(zk.zk_table_batch_timing(handle.value, 1, None, 0, None))
    _lib.check(zk.zk_g1_table_msm_batch(handle.value, ptrs, n, args.steps, outs_host.data_ptr()))   # warm-up (sizes the staging ring)
    gathered = torch.empty(world, args.steps, 96, dtype=torch.uint8, device="cuda")
    sums = torch.empty(args.steps, 144, dtype=torch.uint8, device="cuda")
    rep_s = []
    for rep in range(E2E_REPS):
        outs_host.zero_()
        barrier()
        t0 = time.perf_counter()
        _lib.check(zk.zk_g1_table_msm_batch(handle.value, ptrs, n, args.steps, outs_host.data_ptr()))
        if world > 1:
            with torch.cuda.stream(side):
                parts = outs_host[:, :96].contiguous().cuda(non_blocking=True)
                dist.all_gather_into_tensor(gathered.view(-1), parts.view(-1))
                _lib.check(zk.zk_g1_sum_strided_dev(gathered.data_ptr(), world, args.steps, sums.data_ptr(), side.cuda_stream))
                outs_host.copy_(sums)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert bytes(outs_host[-1].numpy())[:96] == expected_point(batches[slots[-1]]["expect_dlog"]), "e2e result mismatch"
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rep_s.append(float(t.item()))
    e2e_s = sorted(rep_s)[len(rep_s) // 2]
    e2e_val = n_total * args.steps / e2e_s / 1e6
    # per-step split of the last repetition (this rank): upload, compute, and the time the compute
    # stream sat waiting for an upload
    tbuf = (ctypes.c_float * (4 * 64))()
    tsteps = ctypes.c_size_t()
    _lib.check(zk.zk_table_batch_timing(handle.value, 0, tbuf, 4 * 64, ctypes.byref(tsteps)))
    tm = [[float(tbuf[4 * i + j]) for j in range(4)] for i in range(tsteps.value)]
    e2e_split = None
    if tm:
        med = lambda xs: sorted(xs)[len(xs) // 2]
        groups = sorted({(x[2], x[3]) for x in tm})                 # kernel window of every group of steps
        sizes = [sum(1 for x in tm if (x[2], x[3]) == g) for g in groups]
        e2e_split = {"steps_timed": len(tm), "h2d_ms_per_step_median": med([x[1] - x[0] for x in tm]),
                     "groups": [{"steps": k, "kernels_start_ms": g[0], "kernels_ms": g[1] - g[0]} for g, k in zip(groups, sizes)],
                     "compute_ms_per_step": sum(g[1] - g[0] for g in groups) / len(tm),
                     "compute_stream_stall_ms_total": groups[0][0] + sum(max(0.0, groups[i][0] - groups[i - 1][1]) for i in range(1, len(groups))),
                     "first_kernel_after_ms": groups[0][0], "last_kernel_done_ms": groups[-1][1]}

    # ---- secondary legs (every rank takes part; each is exact-checked) ---------------------------
    _lib.check(zk.zk_table_free(handle.value))
    host0 = batches[0]["host"]                             # kept for the CPU baseline leg
    del batches, d_ring
    torch.cuda.empty_cache()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    legs = {}

    def leg(name, fn):
        t0 = time.time()
        try:
            legs[name] = fn()
        except Exception as e:                                 # never lose the headline line
            legs[name] = {"error": repr(e)}
        legs.setdefault("leg_seconds", {})[name] = round(time.time() - t0, 1)

    if args.groth16:
        # second half of BASELINE.json's metric: Groth16 proofs/s on the synthetic circuits of configs[2]
        # (2^16) and configs[4] (2^20, G2 B-query included; sharded by base range over the N ranks):
        # host witness in, proof bytes out, every proof checked against the closed-form trapdoor identity
        import bench_groth16

        def g16():
            out = []
            for spec in args.groth16:
                ln, circ = spec.split(":")
                out.append(bench_groth16.run(zk, int(ln), 3, circ, dist, quiet=True))
            return out
        leg("groth16", g16)
    if args.sweep:
        import bench_legs
        leg("sweep", lambda: bench_legs.sweep(zk, _lib, torch, dist, rank, world, side, args.sweep, SEED_SCALARS))
    if world == 1 and not args.no_shapes:
        import bench_legs
        leg("oneshot", lambda: bench_legs.oneshot(zk, _lib, torch, side, args.logn, SEED_SCALARS, peak_mac32))
        leg("roofline_g2", lambda: bench_legs.g2_roofline(zk, _lib, torch, side, min(args.logn, 20), SEED_SCALARS, peak_mac32))

    if rank == 0:
        c, W = int(info[0]), int(info[1])
        # kernels of libzkb200 launched in the timed region, per group of queued MSMs: 5 for the sort and
        # the accumulation (digit count pass with the scalar check, 2 scan kernels, digit scatter pass,
        # k_accumulate) + the batched tail (2 partial fix-ups, reduce chunks, tree levels,
        # combine+finalize) + the shard sum for N > 1
        group = QUEUE if pipelined else 1
        groups = max(timed_joins, (args.steps + group - 1) // group)     # a table beyond 2^21 points joins more often
        levels = tail_tree_levels(c, int(info[2]), min(group, args.steps),
                                  torch.cuda.get_device_properties(local_rank).multi_processor_count)
        gpu_launches = groups * (5 + 4 + levels + (1 if world > 1 else 0))
        assert timed_msms == args.steps or timed_joins == 64, (timed_joins, groups, timed_msms)   # the ring keeps 64 joins
        acc_avg = timed_stage_ms[1] / timed_msms                 # k_accumulate time per MSM inside the timed region
        acc_single = sum(acc_ms) / len(acc_ms)                   # ... of a single-MSM launch after it
        achieved = n * MAC32_PER_POINT / (acc_avg * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        algo_bytes = n * (W * 96 + 32)                                   # gathered bases + scalars
        line = {
            "metric": "bls12_381_g1_msm_throughput", "value": value, "unit": "Mpts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32-limb modular (Fp 381-bit, Fr 255-bit)",
            "data": "synthetic",
            "config": {"workload": "G1 MSM, 2^%d points, uniform 255-bit scalars (BASELINE configs[3], SURVEY §8d config 4)" % args.logn,
                       "points": n_total, "points_per_gpu": n, "window_bits": c, "windows": W,
                       "precomputed_table": bool(info[7]), "table_MiB_per_gpu": int(info[5]) >> 20,
                       "segments": int(info[4]), "true_mixed_adds_per_point": W,
                       "pipelined_steps": bool(pipelined),
                       "l2": "inputs larger than L2 (table %d MiB, %d rotating scalar batches of %d MiB)" % (int(info[5]) >> 20, pool, (n * 32) >> 20),
                       "parallelism": "base-range shards x%d, all_gather of 96-B partial sums (one per group of %d steps)" % (world, QUEUE if pipelined else 1), "setup_s": setup_s, "host_affinity": affinity},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "Mpts/s", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 144,
                    "api": "zk_g1_table_msm_batch: K host scalar vectors (pinned) in, K points out; a group of vectors is uploaded while the previous group is accumulated",
                    "h2d_gbs_measured": h2d_gbs, "h2d_bound_mpts": h2d_gbs * 1e9 / 32 / 1e6 * world,
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "timing": "K-step call repeated %d times after one warm-up call, median reported; wall clock, max over ranks" % E2E_REPS,
                    "reps_ms_per_step": [x / args.steps * 1e3 for x in rep_s],
                    "best_value": n_total * args.steps / min(rep_s) / 1e6, "split_last_rep": e2e_split},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "int32-imad", "kernel": "k_accumulate<Fp>", "achieved": achieved,
                         "peak": peak_mac32 / 1e12, "unit": "TMAC32/s", "frac": achieved / (peak_mac32 / 1e12),
                         "traffic": (NCU_DRAM_BYTES_PER_MSM if (world == 1 and args.logn == 20 and c == 17) else None),
                         "traffic_source": "profiles/r02_ncu_accumulate.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of a one-MSM launch; algorithmic: 15 x 96 B + 32 B per point = 1.54 GB)",
                         "kernel_ms": acc_avg, "kernel_ms_note": "k_accumulate launch durations summed over the timed region (CUDA events on the launching stream) / MSMs processed; one launch accumulates all the MSMs of a join",
                         "launches_in_timed_region": timed_joins, "msms_per_launch": timed_msms / max(timed_joins, 1),
                         "timed_region_stage_ms_per_step": [x / timed_msms for x in timed_stage_ms],
                         "kernel_ms_single_msm_launch": acc_single, "stages_ms_last_step": stages_last,
                         "algorithmic_mac32_per_point": MAC32_PER_POINT,
                         "peak_source": "mad.lo.cc/madc.hi.cc chains measured in this process (zk_bench_intpipe), /2",
                         "whole_step_frac": n_total / world * MAC32_PER_POINT / (dev_ms / args.steps * 1e-3) / peak_mac32,
                         "hbm": {"achieved_gbs": algo_bytes / (dev_ms / args.steps * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "frac": algo_bytes / (dev_ms / args.steps * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        }
        line.update({k: v for k, v in legs.items()})
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            sample = min(1 << 19, n_total)     # ~10 s of the reference's algorithm on 16 cores
            sc_raw = host0.numpy().tobytes()[:sample * 32]
            secs, out = cpu_fold_msm(bases[:sample * 96].tobytes(), sc_raw, sample, threads)
            line["cpu_baseline"] = {"value": sample / secs / 1e6, "unit": "Mpts/s", "cores": threads, "kind": "port",
                                    "sample": "first 2^%d of the 2^%d points, oracle/c fold of double-and-add scalar muls "
                                              "(curve.ml:91-118) on all host threads, %.1f s" % (sample.bit_length() - 1, args.logn, secs)}
            try:                                                   # informational; never lose the headline line
                n_timed = min(max(1 << 18, sample), n_total)      # never shorter than the sample it is checked on
                line["cpu_pippenger"] = cpu_pippenger_line(bases[:n_timed * 96].tobytes(),
                                                           host0.numpy().tobytes()[:n_timed * 32],
                                                           sample, out, n_timed, threads)
            except Exception as e:
                line["cpu_pippenger"] = {"error": repr(e)}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_JSON_OUT = sys.stdout


def main():
    global _JSON_OUT
    # stdout carries exactly ONE line, the JSON record: keep a private handle on the real stdout and
    # point fd 1 at stderr so that library chatter (e.g. NCCL's version banner) cannot precede it
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--logn", type=int, default=LOG_N)
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--groth16", nargs="*", default=["16:mulchain", "20:mulchain", "20:r1cs"],
                    help="also time Groth16 prove: log2(constraints):circuit (mulchain | r1cs), on all N ranks")
    ap.add_argument("--sweep", type=int, nargs="*", default=list(range(16, 25)),
                    help="G1 MSM sweep sizes (log2 points, BASELINE configs[3]); empty = skip")
    ap.add_argument("--no-shapes", action="store_true", help="skip the one-shot and G2 legs (N = 1)")
    ap.add_argument("--no-pipeline", action="store_true", help="plain stream order between steps")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
