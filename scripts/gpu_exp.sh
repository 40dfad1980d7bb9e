#!/bin/bash
# A/B of training-kernel variants: per-op times of the eager roofline re-run in bench.py
mkdir -p gpurun_out
for v in "$@"; do
  env $v timeout 600 python bench.py --no-cpu-baseline --no-eager-gpu-baseline --no-skip-leg --no-render --steps 10 > gpurun_out/bench_exp.log 2> gpurun_out/bench_exp.err
  python - "$v" <<P
import json, sys
d = json.loads(open('gpurun_out/bench_exp.log').read().strip().splitlines()[-1])
print(sys.argv[1], 'ms/step', round(d['ms_per_step'], 3), {k: round(v['ms_total'] / v['calls'], 3) for k, v in d['ops_ms'].items() if k.startswith('field')})
P
done
