#!/bin/bash
# GPU session 11: accumulate time per MSM against the number of MSMs in the launch; accumulate variants in batched mode
set -x
mkdir -p gpurun_out/s11
O=gpurun_out/s11
for ln in 20 17; do for q in 2 5 10 20; do
ZKB200_BENCH_QUEUE=$q timeout 200 python bench.py --logn $ln --steps 20 --warmup 5 --no-cpu --groth16 --sweep --no-shapes > $O/b_${ln}_q$q.json 2> $O/b_${ln}_q$q.err
done; done
ZKB200_ACC_VARIANT=8 timeout 200 python bench.py --logn 20 --steps 20 --warmup 5 --no-cpu --groth16 --sweep --no-shapes > $O/b_20_q20_v8.json 2> $O/b_20_v8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s11/b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.4f single %.4f'%(r['kernel_ms'], r['kernel_ms_single_msm_launch']), [round(x,3) for x in r['timed_region_stage_ms_per_step']], 'e2e %.1f'%d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
echo done
