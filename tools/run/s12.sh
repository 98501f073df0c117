#!/bin/bash
# GPU session 12: streaming cache hints on entries / bucket sums
set -x
mkdir -p gpurun_out/s12
O=gpurun_out/s12
timeout 600 python -m pytest tests/test_gpu_msm.py -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
for ln in 20 17; do
timeout 200 python bench.py --logn $ln --steps 20 --warmup 5 --no-cpu --groth16 --sweep --no-shapes > $O/b_${ln}.json 2> $O/b_${ln}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s12/b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.4f single %.4f'%(r['kernel_ms'], r['kernel_ms_single_msm_launch']), [round(x,3) for x in r['timed_region_stage_ms_per_step']], 'e2e %.1f'%d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
echo done
