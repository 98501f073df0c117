#!/bin/bash
# GPU session 23 (2 GPUs): witness upload sharded over the ranks + all_gather on the devices
set -x
mkdir -p gpurun_out/s23
O=gpurun_out/s23
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/bench_groth16.py --logn 16 20 --iters 4 > $O/groth16_n2.jsonl 2> $O/err.log
python - <<'PY'
import json
for l in open('gpurun_out/s23/groth16_n2.jsonl'):
    if not l.startswith('{'): continue
    d=json.loads(l); print('N=2', d['log_n'], d['circuit'], 'prove_ms %.2f'%d['prove_ms'], 'dev %.2f'%d['device_ms'], d['exact_ok'], d['stages_ms'], d['witness_memory'], d['h2d_bytes_per_proof'])
PY
tail -n 4 $O/err.log
echo done
