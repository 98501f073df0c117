#!/bin/bash
# GPU session 6: full -m gpu suite with durations, queue depth 32 on the 8-GPU shard size, Groth16 stage budget +
# launch list at 2^20, compute-sanitizer memcheck / racecheck on the small pass and on smoke()
set -x
mkdir -p gpurun_out/s6
O=gpurun_out/s6
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -20 $O/pytest.log
for q in 16 32; do ZKB200_BENCH_QUEUE=$q timeout 300 python bench.py --logn 17 --steps 32 --no-cpu --groth16 --sweep --no-shapes > $O/bench_2e17_q$q.json 2> $O/bench_2e17_q$q.err; done
timeout 300 python bench.py --logn 20 --steps 32 --no-cpu --groth16 --sweep --no-shapes > $O/bench_2e20_q32.json 2> $O/bench_2e20.err
python - <<'PY'
import json
for f in ('bench_2e17_q16.json','bench_2e17_q32.json','bench_2e20_q32.json'):
    try:
        d=json.loads(open('gpurun_out/s6/'+f).read().strip().splitlines()[-1])
        print(f, 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['stages_ms_last_step'], 'e2e %.1f'%d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
timeout 900 python tools/bench_groth16.py --logn 16 20 --iters 3 --circuit mulchain r1cs > $O/groth16.jsonl 2> $O/groth16.err
cut -c1-900 $O/groth16.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_groth16_2e20.csv python tools/bench_groth16.py --logn 20 --iters 1 > $O/ncu_groth16.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file $O/memcheck_small.log python tools/sanitize_small.py > $O/memcheck_small.out 2>&1; echo "memcheck small rc=$?"
tail -3 $O/memcheck_small.log; tail -2 $O/memcheck_small.out | cut -c1-600
timeout 900 compute-sanitizer --tool racecheck --log-file $O/racecheck_small.log python tools/sanitize_small.py > $O/racecheck_small.out 2>&1; echo "racecheck small rc=$?"
tail -3 $O/racecheck_small.log
timeout 900 compute-sanitizer --tool memcheck --log-file $O/memcheck_smoke.log python __graft_entry__.py smoke > $O/memcheck_smoke.out 2>&1; echo "memcheck smoke rc=$?"
tail -3 $O/memcheck_smoke.log; tail -2 $O/memcheck_smoke.out
echo done
