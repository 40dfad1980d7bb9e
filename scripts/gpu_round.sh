#!/bin/bash
# One gpurun call: the GPU test suite, smoke, the default bench, a short train-step bench of the A/B arm given as
# "VAR=VALUE" in $1 (optional) and (N > 1 boxes) a 2-GPU bench.  Everything is logged under gpurun_out/.  Each stage has
# its own timeout so a hung kernel cannot run into gpurun's limit.  ncu evidence is refreshed by scripts/gpu_profile.sh.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu =="
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
echo "== bench =="
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"
tail -c 600 gpurun_out/bench.log
if [ -n "${1:-}" ]; then
  echo "== bench, $1 =="
  env "$1" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-render --no-skip-leg > gpurun_out/bench_ab.log 2> gpurun_out/bench_ab.err
  echo "bench ab exit $?"
fi
python - <<'P'
import json
for n in ('bench', 'bench_ab'):
    try:
        d = json.loads(open(f'gpurun_out/{n}.log').read().strip().splitlines()[-1])
        print(n, 'ms/step', round(d['ms_per_step'], 3), 'rays/s', round(d['value']), 'e2e', round(d['e2e']['value']),
              {k: round(v['ms_total'] / v['calls'], 3) for k, v in d.get('ops_ms', {}).items() if k.startswith('field')})
    except Exception as e:
        print(n, 'n/a', e)
P
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "== bench, 2 GPUs =="
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 2 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
  echo "bench n2 exit $?"
  tail -c 600 gpurun_out/bench_n2.log
fi
