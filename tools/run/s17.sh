#!/bin/bash
# GPU session 17: compact (rolled) products in the latency-bound point formulas
set -x
mkdir -p gpurun_out/s17
O=gpurun_out/s17
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 600 python tools/bench_groth16.py --logn 16 20 --iters 5 --circuit mulchain r1cs 2>$O/err.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('g16', d['log_n'], d['circuit'], 'prove_ms %.2f'%d['prove_ms'], d['exact_ok'], d['stages_ms'], 'setup %.1f'%d['setup_s'])
"
for ln in 20 17; do
timeout 200 python bench.py --logn $ln --steps 20 --warmup 5 --no-cpu --groth16 --sweep > $O/b_${ln}.json 2> $O/b_${ln}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s17/b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.4f'%r['kernel_ms'], [round(x,3) for x in r['timed_region_stage_ms_per_step']], 'e2e %.1f'%d['e2e']['value'])
        print('   oneshot', json.dumps(d.get('oneshot'))[:500]); print('   g2', json.dumps(d.get('roofline_g2'))[:400])
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 python tools/bench_pinocchio.py > $O/pinocchio.json 2>>$O/err.log; cut -c1-400 $O/pinocchio.json
timeout 300 python tools/bench_verify.py > $O/verify.json 2>>$O/err.log; cut -c1-400 $O/verify.json
tail -n 3 $O/err.log
echo done
