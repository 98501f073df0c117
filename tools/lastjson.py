#!/usr/bin/env python3
"""Print a compact summary of the last JSON line on stdin (bench.py output)."""
import json
import sys

lines = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
if not lines:
    print("no JSON line")
    sys.exit(1)
d = json.loads(lines[-1])
r = d.get("roofline", {})
print("N=%s value %.1f %s  ms/step %.3f  e2e %.1f  c=%s W=%s  stages %s  acc_frac %.3f  step_frac %.3f  clocks %s" % (
    d.get("n_gpus"), d["value"], d["unit"], d["ms_per_step"], d.get("e2e", {}).get("value", float("nan")),
    d["config"].get("window_bits"), d["config"].get("windows"),
    [round(x, 3) for x in r.get("stages_ms_last_step", [])], r.get("frac", float("nan")),
    r.get("whole_step_frac", float("nan")), d.get("clocks", {}).get("sm_mhz")))
