#!/bin/bash
# GPU session 2: multi-device C ABI on 2 GPUs + regression of the single-device suites after the runtime refactor
set -x
mkdir -p gpurun_out/s2
O=gpurun_out/s2
nvidia-smi -L > $O/gpus.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python tests/multidev_worker.py 15 > $O/multidev.json 2> $O/multidev.err; echo "worker rc=$?"
cat $O/multidev.json; tail -5 $O/multidev.err
timeout 900 python -m pytest tests -m gpu -x -q -k "not config5" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -8 $O/pytest.log
