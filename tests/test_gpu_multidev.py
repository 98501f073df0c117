"""The multi-GPU path BEHIND the C ABI (VERDICT r1 "missing" 1): after
zk_init_devices one process drives every visible GPU, key and table handles spread their base
ranges over the devices, and ONE zk_groth16_prove_r1cs / zk_g*_table_msm(_batch) call returns the
finished result — equal to the single-device bytes and to the oracle's closed form
(groth16.ml:123-161; protocol.mli:23 is the OCaml entry this serves).  With one visible GPU the
same code path runs on a one-device list."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_one_prove_call_on_all_visible_devices(zk):
    env = {k: v for k, v in os.environ.items() if k not in ("ZKB200_DEVICE", "ZKB200_DEVICES", "LOCAL_RANK")}
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multidev_worker.py"), "15"], capture_output=True,
                       text=True, timeout=900, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    rec = json.loads(p.stdout.strip().splitlines()[-1])
    assert rec["ok"] and rec["multi_equals_single_device"] and rec["multi_equals_oracle"], rec
    assert rec["device_count"] == rec["devices"] >= 1
