#!/usr/bin/env python3
"""Small, exact-checked pass over the MSM and evaluation-form prover kernels, meant to run under
compute-sanitizer:

  compute-sanitizer --tool memcheck  python tools/sanitize_small.py
  compute-sanitizer --tool racecheck python tools/sanitize_small.py

Covers: one-shot and table G1 / G2 MSMs (with and without precomputed windows) on uniform and on
witness-like scalars (the latter drive k_fix_heavy's shared-memory tree), the pipelined queue with
overflows (depth 4, 11 MSMs), the batched host API (cp.async staging, copy stream), and the
evaluation-form Groth16 prover at 2^10 constraints on both circuits (sparse mat-vec, fused
shared-memory NTT passes, three MSMs with batched tails).  Every result is compared with its closed
form, so a sanitizer-clean run is also a correct one.  The dense-QAP prover and the pairing kernels
are covered by running `python __graft_entry__.py smoke` under the same tools.  Prints one JSON line."""
import ctypes
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np
import torch

from zukelang_b200 import _lib
from zukelang_b200.curve import R

import bench_groth16


def words(ints):
    return np.frombuffer(b"".join(v.to_bytes(32, "little") for v in ints), dtype=np.uint64).reshape(len(ints), 4).copy()


def fixed(zk, group, ints):
    raw = 96 if group == "g1" else 192
    out = np.empty(raw * len(ints), dtype=np.uint8)
    w = words(ints)
    _lib.check(getattr(zk, "zk_%s_fixed_base_mul" % group)(w.ctypes.data, len(ints), out.ctypes.data))
    return out


def msm_family(zk, group, n, rng, report):
    raw, outn = (96, 144) if group == "g1" else (192, 288)
    dl = [rng.randrange(1, R) for _ in range(n)]
    bases = fixed(zk, group, dl)
    uniform = [rng.randrange(R) for _ in range(n)]
    witness = [rng.randrange(2) if rng.random() < 0.9 else rng.randrange(R) for _ in range(n)]
    for name, ks in (("uniform", uniform), ("witness", witness)):
        exp = bytes(fixed(zk, group, [sum(a * b for a, b in zip(ks, dl)) % R]))
        sc = words(ks)
        out = (ctypes.c_uint8 * outn)()
        _lib.check(getattr(zk, "zk_%s_msm" % group)(bases.ctypes.data, None, sc.ctypes.data, n, out))
        report["%s_oneshot_%s" % (group, name)] = bytes(out)[:raw] == exp
        for pre in (0, 1):
            h = ctypes.c_uint64()
            _lib.check(getattr(zk, "zk_%s_table_load" % group)(bases.ctypes.data, None, n, pre, 0, ctypes.byref(h)))
            _lib.check(getattr(zk, "zk_%s_table_msm" % group)(h.value, sc.ctypes.data, n, out))
            report["%s_table_pre%d_%s" % (group, pre, name)] = bytes(out)[:raw] == exp
            if pre and name == "witness":
                # pipelined queue, depth 4, 11 MSMs (two overflows + a final join), then the batched host API
                count = 11
                _lib.check(zk.zk_table_pipeline(h.value, 4))
                d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
                d_out = torch.zeros(count, outn, dtype=torch.uint8, device="cuda")
                st = torch.cuda.Stream()
                with torch.cuda.stream(st):
                    for i in range(count):
                        _lib.check(getattr(zk, "zk_%s_table_msm_dev" % group)(h.value, d_sc.data_ptr(), n, d_out[i].data_ptr(), st.cuda_stream))
                    _lib.check(zk.zk_table_join(h.value, st.cuda_stream))
                st.synchronize()
                report["%s_pipelined" % group] = all(bytes(d_out[i].cpu().numpy())[:raw] == exp for i in range(count))
                _lib.check(zk.zk_table_pipeline(h.value, 0))
                ptrs = (ctypes.c_void_p * 5)(*[sc.ctypes.data] * 5)
                outs = (ctypes.c_uint8 * (outn * 5))()
                _lib.check(getattr(zk, "zk_%s_table_msm_batch" % group)(h.value, ptrs, n, 5, outs))
                report["%s_batch" % group] = all(bytes(outs[i * outn:i * outn + raw]) == exp for i in range(5))
            _lib.check(zk.zk_table_free(h.value))


def main():
    zk = _lib.lib()
    rng = random.Random(0x53414E49)
    report = {}
    msm_family(zk, "g1", int(os.environ.get("SANITIZE_G1_POINTS", "4096")), rng, report)
    msm_family(zk, "g2", int(os.environ.get("SANITIZE_G2_POINTS", "1024")), rng, report)
    for circ in ("mulchain", "r1cs"):
        rec = bench_groth16.run(zk, 10, 1, circ, None, quiet=True)
        report["groth16_2e10_%s" % circ] = bool(rec["exact_ok"])
    ok = all(report.values())
    print(json.dumps({"sanitize_small": report, "ok": ok}), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
