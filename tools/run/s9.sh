#!/bin/bash
# GPU session 9: full parity suite on the final kernels, the driver's N = 1 bench command, launch list of it
set -x
mkdir -p gpurun_out/s9
O=gpurun_out/s9
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -10 $O/pytest.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
/usr/bin/time -v timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
grep -E "Elapsed|Maximum resident" $O/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s9/bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%r['frac'], 'whole %.3f'%r['whole_step_frac'], 'kernel_ms %.3f'%r['kernel_ms'], [round(x,3) for x in r['timed_region_stage_ms_per_step']], 'e2e %.1f'%d['e2e']['value'], [round(x,3) for x in d['e2e']['reps_ms_per_step']])
for g in d['groth16']: print(g['log_n'], g['circuit'], 'prove_ms %.2f'%g['prove_ms'], g['exact_ok'], g['stages_ms'], 'setup %.1f'%g['setup_s'])
print(d['leg_seconds']); print(json.dumps(d['roofline_g2'])[:600]); print(json.dumps(d['oneshot'])[:800]); print(d.get('cpu_baseline')); print(d.get('cpu_pippenger'))
for s in d['sweep']: print(s['log_n'], 'single %.1f pipelined %.1f Mpts/s'%(s['mpts_single'], s['mpts_pipelined']), s['exact_ok'])
PY
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-600 $O/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench_n1.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --groth16 16:mulchain --sweep 20 --no-shapes > $O/ncu_bench.log 2>&1
echo done
