#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "vol_decoder" 2>&1 | grep -v Warning | tail -25
for dec in hn cudnn; do
  HN_DECODER=$dec timeout 600 python bench.py --no-cpu-baseline --no-eager-gpu-baseline --no-skip-leg --no-render > gpurun_out/bench_dec_$dec.log 2> gpurun_out/bench_dec_$dec.err
  python - <<P
import json
d = json.loads(open('gpurun_out/bench_dec_$dec.log').read().strip().splitlines()[-1])
print('$dec', 'ms/step', round(d['ms_per_step'], 3), {k: round(v['ms_total'] / v['calls'], 3) for k, v in d['ops_ms'].items() if 'decoder' in k or k.startswith('field')})
P
done
