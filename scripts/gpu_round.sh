#!/bin/bash
# One gpurun call: the GPU test suite, smoke, the default bench and (N > 1 boxes) a 2-GPU bench.  Everything is
# logged under gpurun_out/.  Each stage has its own timeout so a hung kernel cannot run into gpurun's limit.
# ncu evidence is refreshed separately by scripts/gpu_profile.sh.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu =="
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
echo "== smoke =="
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
echo "== bench =="
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"
tail -c 1500 gpurun_out/bench.log
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "== bench, 2 GPUs =="
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 2 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
  echo "bench n2 exit $?"
  tail -c 600 gpurun_out/bench_n2.log
fi
