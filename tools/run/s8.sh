#!/bin/bash
# GPU session 8: per-slice partial fix-up, L = 32 reduce chunks for big batches, first-group heuristic of the batch API
set -x
mkdir -p gpurun_out/s8
O=gpurun_out/s8
timeout 900 python -m pytest tests -m gpu -x -q -k "not config5 and not multidev" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
ZKB200_LIB=$PWD/zukelang_b200/libzkb200_checked.so timeout 600 python tools/sanitize_small.py > $O/checked_small.json 2> $O/checked_small.err; echo "checked small rc=$?"; cut -c1-300 $O/checked_small.json; tail -3 $O/checked_small.err
ZKB200_LIB=$PWD/zukelang_b200/libzkb200_checked.so timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_sparse.py -m gpu -x -q > $O/checked_pytest.log 2>&1; echo "checked pytest rc=$?"; tail -3 $O/checked_pytest.log
timeout 300 python bench.py --logn 17 --steps 20 --warmup 5 --no-cpu --groth16 --sweep --no-shapes > $O/bench_2e17.json 2> $O/bench_2e17.err
timeout 300 python bench.py --logn 20 --steps 20 --warmup 5 --no-cpu --groth16 --sweep --no-shapes > $O/bench_2e20.json 2> $O/bench_2e20.err
python - <<'PY'
import json
for f in ('bench_2e17.json','bench_2e20.json'):
    try:
        d=json.loads(open('gpurun_out/s8/'+f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%r['frac'], 'kernel_ms %.3f single %.3f'%(r['kernel_ms'], r['kernel_ms_single_msm_launch']), [round(x,3) for x in r['timed_region_stage_ms_per_step']], 'e2e %.1f'%d['e2e']['value'], [round(x,3) for x in d['e2e']['reps_ms_per_step']], d['e2e']['split_last_rep']['groups'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -n 3 $O/bench_2e17.err $O/bench_2e20.err
timeout 600 python tools/bench_groth16.py --logn 16 20 --iters 3 --circuit mulchain > $O/groth16.jsonl 2> $O/groth16.err
python - <<'PY'
import json
for l in open('gpurun_out/s8/groth16.jsonl'):
    d=json.loads(l); print(d['log_n'], d['circuit'], 'prove_ms %.2f dev %.2f ok %s'%(d['prove_ms'], d['device_ms'], d['exact_ok']), d['stages_ms'])
PY
timeout 300 python bench.py --logn 17 --steps 20 --warmup 5 --no-cpu --groth16 --sweep --no-shapes > $O/plain_2e17.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_2e17.csv python bench.py --logn 17 --steps 20 --warmup 5 --no-cpu --groth16 --sweep --no-shapes > $O/ncu_2e17.log 2>&1
echo done
