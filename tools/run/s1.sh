#!/bin/bash
# GPU session 1 (round 2): parity tests, accumulate variants A/B, bench at 2^20 and at the 8-GPU shard size, launch list
set -x
mkdir -p gpurun_out/s1
O=gpurun_out/s1
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -25 $O/pytest.log
# accumulate variants: direct loads (8 / 4) vs cp.async staging (9 / 5)
for v in 8 9; do ZKB200_ACC_VARIANT=$v timeout 300 python tools/gpu_probe.py --logn 20 --precompute 1 --iters 4 >> $O/ab_g1.jsonl 2>>$O/ab.err; done
for v in 8 9; do for c in 16 17; do ZKB200_ACC_VARIANT=$v ZKB200_WINDOW_BITS_PRE=$c timeout 300 python tools/gpu_probe.py --logn 17 --precompute 1 --iters 4 >> $O/ab_g1_2e17.jsonl 2>>$O/ab.err; done; done
for v in 4 5; do ZKB200_ACC_VARIANT_G2=$v timeout 300 python tools/gpu_probe.py --group g2 --logn 18 20 --precompute 1 --iters 3 >> $O/ab_g2.jsonl 2>>$O/ab.err; done
cat $O/ab_g1.jsonl $O/ab_g1_2e17.jsonl $O/ab_g2.jsonl | cut -c1-420
timeout 600 python bench.py --groth16-logn 16 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
cut -c1-3000 $O/bench_n1.json
timeout 300 python bench.py --logn 17 --no-cpu --groth16-logn > $O/bench_2e17.json 2> $O/bench_2e17.err
ZKB200_WINDOW_BITS_PRE=17 timeout 300 python bench.py --logn 17 --no-cpu --groth16-logn > $O/bench_2e17_c17.json 2> $O/bench_2e17_c17.err
cut -c1-1500 $O/bench_2e17.json; cut -c1-1500 $O/bench_2e17_c17.json
timeout 300 python bench.py --logn 17 --steps 3 --no-cpu --groth16-logn > $O/plain_2e17.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_2e17.csv python bench.py --logn 17 --steps 3 --no-cpu --groth16-logn > $O/ncu_2e17.log 2>&1
echo done
