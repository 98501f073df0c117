#!/bin/bash
# GPU session 20: early-B order from 2^15 constraints on — parity of the affected tests, proofs at 2^16
set -x
mkdir -p gpurun_out/s20
O=gpurun_out/s20
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_prove.py tests/test_gpu_sparse.py tests/test_gpu_multidev.py -m gpu -x -q -k "not config5" > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 300 python tools/bench_groth16.py --logn 15 16 17 --iters 5 --circuit mulchain r1cs > $O/groth16.jsonl 2>$O/err.log
python - <<'PY'
import json
for l in open('gpurun_out/s20/groth16.jsonl'):
    d=json.loads(l); print('g16', d['log_n'], d['circuit'], 'prove_ms %.2f'%d['prove_ms'], d['exact_ok'], d['stages_ms'])
PY
echo done
