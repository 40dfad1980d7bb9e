#!/bin/bash
# One gpurun call: GPU test suite, smoke, default bench.  Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu =="
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -4
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
echo "== bench =="
timeout 900 python bench.py "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"
tail -c 1500 gpurun_out/bench.log
tail -5 gpurun_out/bench.err
