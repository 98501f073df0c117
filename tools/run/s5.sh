#!/bin/bash
# GPU session 5: fused NTT passes (parity + A/B), paired G2 variants, bench legs, launch list at Q = 16 on a 2^17 shard,
# ncu --set full of the accumulation kernel
set -x
mkdir -p gpurun_out/s5
O=gpurun_out/s5
timeout 900 python -m pytest tests -m gpu -x -q -k "not config5" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -6 $O/pytest.log
for f in 0 1; do ZKB200_NTT_FUSED=$f timeout 600 python tools/bench_groth16.py --logn 16 20 --iters 3 >> $O/ntt_ab.jsonl 2>>$O/ab.err; done
cut -c1-600 $O/ntt_ab.jsonl
for v in 5 9; do ZKB200_ACC_VARIANT_G2=$v timeout 300 python tools/gpu_probe.py --group g2 --logn 18 20 --precompute 1 --iters 3 >> $O/ab_g2.jsonl 2>>$O/ab.err; done
timeout 300 python tools/gpu_probe.py --logn 17 20 --precompute 1 --iters 4 > $O/probe_g1.jsonl 2>>$O/ab.err
cat $O/ab_g2.jsonl $O/probe_g1.jsonl | grep msm | cut -c1-330
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
tail -3 $O/bench_n1.err
timeout 300 python bench.py --logn 17 --no-cpu --groth16 --sweep --no-shapes > $O/bench_2e17.json 2> $O/bench_2e17.err
python - <<'PY'
import json
for f in ('bench_n1.json','bench_2e17.json'):
    d=json.loads(open('gpurun_out/s5/'+f).read().strip().splitlines()[-1])
    print(f, 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['stages_ms_last_step'], 'e2e %.1f'%d['e2e']['value'])
    for k in ('groth16','sweep','oneshot','roofline_g2','leg_seconds'):
        if k in d: print('  ',k, json.dumps(d[k])[:1500])
PY
timeout 300 python bench.py --logn 17 --steps 16 --no-cpu --groth16 --sweep --no-shapes > $O/plain_2e17.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_2e17_q16.csv python bench.py --logn 17 --steps 16 --no-cpu --groth16 --sweep --no-shapes > $O/ncu_2e17.log 2>&1
timeout 300 python tools/gpu_probe.py --logn 20 --precompute 1 --iters 1 > $O/plain_probe.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -c 1 -o $O/prof_acc_r2 python tools/gpu_probe.py --logn 20 --precompute 1 --iters 1 > $O/ncu_full.log 2>&1
echo done
