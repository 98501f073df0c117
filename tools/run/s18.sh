#!/bin/bash
# GPU session 18: proof order B -> (quotient || B tail) -> A + C; queue_limit; full suite
set -x
mkdir -p gpurun_out/s18
O=gpurun_out/s18
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 600 python tools/bench_groth16.py --logn 16 20 --iters 5 --circuit mulchain r1cs 2>$O/err.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('g16', d['log_n'], d['circuit'], 'prove_ms %.2f'%d['prove_ms'], d['exact_ok'], d['stages_ms'])
"
timeout 300 python tools/bench_pinocchio.py > $O/pinocchio.json 2>>$O/err.log; cut -c1-400 $O/pinocchio.json
tail -n 3 $O/err.log
echo done
