#!/bin/bash
# A/B of NCCL channel counts for the weak-scaling train step at N GPUs (train leg only)
N=${1:-8}
mkdir -p gpurun_out
P=29617
for v in "NCCL_MAX_NCHANNELS=2" "NCCL_MAX_NCHANNELS=4" "NCCL_X=0"; do
  env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --no-skip-leg --no-render --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/bench_ab.log 2> gpurun_out/bench_ab.err
  python - "$v" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/bench_ab.log').read().strip().splitlines()[-1])
    print(sys.argv[1], 'weak ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'strong', round(d['strong']['ms_per_step'], 3) if d.get('strong') else None)
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
  P=$((P+1))
done
