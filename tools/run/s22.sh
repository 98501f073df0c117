#!/bin/bash
# GPU session 22 (2 GPUs): the in-process multi-device path after the queue-sizing fix, quick parity subset
set -x
mkdir -p gpurun_out/s22
O=gpurun_out/s22
timeout 600 python -m pytest tests/test_gpu_multidev.py tests/test_gpu_msm.py tests/test_gpu_prove.py -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
ZKB200_DEVICES=0,1 timeout 600 python tools/bench_groth16.py --logn 16 20 --iters 3 > $O/groth16_inproc_2dev.jsonl 2> $O/err.log
python - <<'PY'
import json
for l in open('gpurun_out/s22/groth16_inproc_2dev.jsonl'):
    d=json.loads(l); print('in-process', d['devices_per_process'], 'devices', d['log_n'], 'prove_ms %.2f'%d['prove_ms'], d['exact_ok'], d['stages_ms'])
PY
tail -n 3 $O/err.log
echo done
