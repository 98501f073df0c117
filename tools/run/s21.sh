#!/bin/bash
# GPU session 21 (8 GPUs): the driver's N = 8 command on the round's final kernels
set -x
mkdir -p gpurun_out/s21
O=gpurun_out/s21
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s21/bench_n8.json').read().strip().splitlines()[-1])
r=d['roofline']
print('N', d['n_gpus'], 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%r['frac'], 'whole %.3f'%r['whole_step_frac'], [round(x,3) for x in r['timed_region_stage_ms_per_step']], 'e2e %.1f'%d['e2e']['value'], [round(x,3) for x in d['e2e']['reps_ms_per_step']])
for g in d.get('groth16', []): print(g.get('log_n'), g.get('circuit'), 'prove_ms %.2f'%g['prove_ms'], g['exact_ok'], g['stages_ms'], 'setup %.1f'%g['setup_s']) if 'prove_ms' in g else print(g)
print(d['leg_seconds'])
for s in d.get('sweep', []): print(s['log_n'], 'single %.1f pipelined %.1f Mpts/s'%(s['mpts_single'], s['mpts_pipelined']), s['exact_ok'])
PY
tail -n 3 $O/bench_n8.err
echo done
