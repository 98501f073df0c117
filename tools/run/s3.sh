#!/bin/bash
# GPU session 3: folded MSM pipeline — parity suite, timings at 2^20 / 2^17, skewed scalars (heavy-bucket path), one-shot shape
set -x
mkdir -p gpurun_out/s3
O=gpurun_out/s3
timeout 900 python -m pytest tests -m gpu -x -q -k "not config5" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -6 $O/pytest.log
timeout 300 python tools/gpu_probe.py --logn 17 20 --precompute 1 --iters 4 > $O/probe_g1.jsonl 2>$O/probe.err
timeout 300 python tools/gpu_probe.py --logn 20 --precompute 1 --iters 3 --dist witness >> $O/probe_g1.jsonl 2>>$O/probe.err
timeout 300 python tools/gpu_probe.py --logn 20 --precompute 0 --iters 3 >> $O/probe_g1.jsonl 2>>$O/probe.err
timeout 300 python tools/gpu_probe.py --group g2 --logn 18 --precompute 1 --iters 3 >> $O/probe_g1.jsonl 2>>$O/probe.err
cut -c1-400 $O/probe_g1.jsonl
timeout 600 python bench.py --groth16-logn 16 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --logn 17 --no-cpu --groth16-logn > $O/bench_2e17.json 2> $O/bench_2e17.err
python - <<'PY'
import json
for f in ('bench_n1.json','bench_2e17.json'):
    d=json.loads(open('gpurun_out/s3/'+f).read().strip().splitlines()[-1])
    print(f, 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['stages_ms_last_step'], 'e2e %.1f'%d['e2e']['value'], d.get('groth16'))
PY
echo done
