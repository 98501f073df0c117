#!/bin/bash
# GPU session 24: zk_groth16_combine through the shard-emulation tests, final N = 1 bench line
set -x
mkdir -p gpurun_out/s24
O=gpurun_out/s24
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_prove.py tests/test_gpu_sparse.py -m gpu -x -q -k "not config5" > $O/pytest.log 2>&1; tail -2 $O/pytest.log
T0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s24/bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%r['frac'], 'whole %.3f'%r['whole_step_frac'], [round(x,3) for x in r['timed_region_stage_ms_per_step']], 'e2e %.1f'%d['e2e']['value'])
for g in d['groth16']: print(g['log_n'], g['circuit'], 'prove_ms %.2f'%g['prove_ms'], g['exact_ok'])
print(d['leg_seconds']); print(d['cpu_baseline']); print(d['cpu_pippenger'])
PY
tail -n 3 $O/bench_n1.err
echo done
