#!/bin/bash
# GPU session 25: final N = 1 bench line (both arms)
set -x
mkdir -p gpurun_out/s25
O=gpurun_out/s25
T0=$(date +%s)
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s25/bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%r['frac'], 'whole %.3f'%r['whole_step_frac'], [round(x,3) for x in r['timed_region_stage_ms_per_step']], 'e2e %.1f'%d['e2e']['value'])
for g in d['groth16']: print(g['log_n'], g['circuit'], 'prove_ms %.2f'%g['prove_ms'], g['exact_ok'])
print(d['leg_seconds']); print(d['cpu_baseline']); print(d['cpu_pippenger'])
PY
tail -n 3 $O/bench_n1.err
echo done
