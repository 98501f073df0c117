#!/bin/bash
# GPU session 26: the full -m gpu suite on the final commit
mkdir -p gpurun_out/s26
timeout 260 python -m pytest tests -m gpu -x -q > gpurun_out/s26/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s26/pytest.log
tail -3 gpurun_out/s26/pytest.log
