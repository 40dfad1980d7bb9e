#!/bin/bash
# One gpurun --gpus N call: the world-size-2 GPU test (NCCL, CUDA-graph steps), then bench.py at N GPUs (weak `value` + the
# strong-scaling leg), optionally the cfg4 / cfg5 render workloads.  Logs under gpurun_out/.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -s > gpurun_out/pytest_multi_n$N.log 2>&1
grep "\[multi\]\|passed\|failed" gpurun_out/pytest_multi_n$N.log | cut -c1-900
P=29517
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $N --no-skip-leg > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench n$N exit $?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
    print('weak: ms/step', round(d['ms_per_step'], 3), 'rays/s', round(d['value']), '| e2e ms', round(d['e2e']['ms_per_step'], 3))
    print('strong:', d.get('strong'))
    r = d.get('render')
    if r: print('render ms/frame', round(r['ms_per_frame'], 2))
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/bench_n$N.err
if [ "${2:-}" = "render" ]; then
  for wl in cfg4 cfg5; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) \
        bench.py --gpus $N --workload $wl --steps 4 --warmup 3 > gpurun_out/bench_${wl}_n$N.log 2> gpurun_out/bench_${wl}_n$N.err
    echo "$wl exit $?"; tail -c 900 gpurun_out/bench_${wl}_n$N.log
  done
fi
