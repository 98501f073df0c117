#!/bin/bash
# GPU session 19: validation of the round's final kernels — full parity suite, checked build, driver bench command,
# 8-GPU shard size, proofs, launch list
set -x
mkdir -p gpurun_out/s19
O=gpurun_out/s19
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -9 $O/pytest.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
ZKB200_LIB=$PWD/zukelang_b200/libzkb200_checked.so timeout 600 python tools/sanitize_small.py > $O/checked_small.json 2> $O/checked_small.err; echo "checked small rc=$?"; cut -c1-200 $O/checked_small.json; tail -n 2 $O/checked_small.err
ZKB200_LIB=$PWD/zukelang_b200/libzkb200_checked.so timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/checked_smoke.log 2>&1; echo "checked smoke rc=$?"; tail -1 $O/checked_smoke.log
ZKB200_LIB=$PWD/zukelang_b200/libzkb200_checked.so timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_sparse.py tests/test_gpu_configs.py -m gpu -x -q -k "not config5" > $O/checked_pytest.log 2>&1; echo "checked pytest rc=$?"; tail -2 $O/checked_pytest.log
timeout 600 python tools/bench_groth16.py --logn 16 20 --iters 5 --circuit mulchain r1cs > $O/groth16.jsonl 2>$O/err.log
python - <<'PY'
import json
for l in open('gpurun_out/s19/groth16.jsonl'):
    d=json.loads(l); print('g16', d['log_n'], d['circuit'], 'prove_ms %.2f'%d['prove_ms'], d['exact_ok'], d['stages_ms'])
PY
T0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --logn 17 --steps 20 --warmup 5 --no-cpu --groth16 --sweep --no-shapes > $O/bench_2e17.json 2> $O/bench_2e17.err
python - <<'PY'
import json
for f in ('bench_n1.json','bench_2e17.json'):
    d=json.loads(open('gpurun_out/s19/'+f).read().strip().splitlines()[-1])
    r=d['roofline']
    print(f, 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%r['frac'], 'whole %.3f'%r['whole_step_frac'], 'kernel_ms %.3f'%r['kernel_ms'], [round(x,3) for x in r['timed_region_stage_ms_per_step']], 'e2e %.1f'%d['e2e']['value'], [round(x,3) for x in d['e2e']['reps_ms_per_step']])
    for g in d.get('groth16', []): print('  ', g['log_n'], g['circuit'], 'prove_ms %.2f'%g['prove_ms'], g['exact_ok'], g['stages_ms'], 'setup %.1f'%g['setup_s'])
    print('  ', d.get('leg_seconds')); print('  ', json.dumps(d.get('roofline_g2'))[:400]); print('  ', json.dumps(d.get('oneshot'))[:700])
PY
tail -n 3 $O/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench_n1.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --groth16 20:r1cs --sweep 20 --no-shapes > $O/ncu_bench.log 2>&1
echo done
