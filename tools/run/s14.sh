#!/bin/bash
# GPU session 14: do the G1 and G2 tails of a proof overlap?  chunk width of the bucket reduction for small batches
set -x
mkdir -p gpurun_out/s14
O=gpurun_out/s14
for mode in 0 1 2 3; do for L in 0 8 16; do
ZKB200_TAIL_MODE=$mode ZKB200_REDUCE_CHUNK=$L timeout 300 python tools/bench_groth16.py --logn 16 --iters 5 --circuit r1cs 2>>$O/err.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('mode $mode L $L', d['log_n'], 'prove_ms %.2f'%d['prove_ms'], d['exact_ok'], d['stages_ms'])
"
done; done
for L in 0 8 16; do
ZKB200_REDUCE_CHUNK=$L timeout 300 python tools/bench_groth16.py --logn 20 --iters 3 --circuit r1cs 2>>$O/err.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('2e20 L $L', 'prove_ms %.2f'%d['prove_ms'], d['exact_ok'], d['stages_ms'])
"
done
tail -3 $O/err.log
echo done
