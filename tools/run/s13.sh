#!/bin/bash
# GPU session 13 (8 GPUs): the driver's N = 8 commands (both arms), in-process 8-device proof
set -x
mkdir -p gpurun_out/s13
O=gpurun_out/s13
nvidia-smi -L > $O/gpus.txt
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s13/bench_n8.json').read().strip().splitlines()[-1])
r=d['roofline']
print('N', d['n_gpus'], 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%r['frac'], 'whole %.3f'%r['whole_step_frac'], [round(x,3) for x in r['timed_region_stage_ms_per_step']], 'e2e %.1f'%d['e2e']['value'], [round(x,3) for x in d['e2e']['reps_ms_per_step']])
for g in d.get('groth16', []): print(g.get('log_n'), g.get('circuit'), 'prove_ms %.2f'%g['prove_ms'], g['exact_ok'], g['stages_ms'], 'setup %.1f'%g['setup_s']) if 'prove_ms' in g else print(g)
print(d['leg_seconds'])
for s in d.get('sweep', []): print(s['log_n'], 'single %.1f pipelined %.1f Mpts/s'%(s['mpts_single'], s['mpts_pipelined']), s['exact_ok'])
PY
tail -n 5 $O/bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err; cut -c1-200 $O/bench_ref_n8.json
ZKB200_DEVICES=0,1,2,3,4,5,6,7 timeout 600 python tools/bench_groth16.py --logn 20 --iters 3 > $O/groth16_inproc_8dev.jsonl 2> $O/groth16_inproc.err
python - <<'PY'
import json
for l in open('gpurun_out/s13/groth16_inproc_8dev.jsonl'):
    d=json.loads(l); print('in-process', d['devices_per_process'], 'devices', d['log_n'], 'prove_ms %.2f'%d['prove_ms'], d['exact_ok'], d['stages_ms'])
PY
echo done
