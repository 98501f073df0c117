#!/bin/bash
# GPU session 16: what bounds the tail kernels?  ncu --set full of k_reduce_chunks (Q = 20 join at 2^17 points; Q = 2 proof tail)
set -x
mkdir -p gpurun_out/s16
O=gpurun_out/s16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reduce_chunks -s 6 -c 1 -o $O/prof_reduce_chunks_q20 python bench.py --logn 17 --steps 20 --warmup 5 --no-cpu --groth16 --sweep --no-shapes > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reduce_tree -s 12 -c 1 -o $O/prof_reduce_tree_q20 python bench.py --logn 17 --steps 20 --warmup 5 --no-cpu --groth16 --sweep --no-shapes > $O/ncu2.log 2>&1
tail -3 $O/ncu1.log $O/ncu2.log
ls -la $O
echo done
