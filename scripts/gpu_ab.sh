#!/bin/bash
# One short gpurun call for an A/B of a training-kernel switch:  bash scripts/gpu_ab.sh ENV_VAR [pytest -k expression]
# runs the selected GPU tests with the default setting, then the train-step bench with the default and with ENV_VAR=0.
set -u
VAR=${1:-HN_TC_RELU_BITS}
SEL=${2:-"relu_mask or field_bf16_backward or network_train_bf16 or graphed_train_step or skipping or full_size"}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -q -x -m gpu -k "$SEL" > gpurun_out/ab_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/ab_pytest.log
tail -5 gpurun_out/ab_pytest.log
B="--steps 30 --warmup 5 --no-cpu-baseline --no-render --no-skip-leg"
timeout 300 python bench.py $B > gpurun_out/ab_on.json 2> gpurun_out/ab_on.err; echo "bench default exit $?"
env $VAR=0 timeout 300 python bench.py $B > gpurun_out/ab_off.json 2> gpurun_out/ab_off.err; echo "bench $VAR=0 exit $?"
python - <<'P'
import json
for n in ('on', 'off'):
    try:
        d = json.loads(open(f'gpurun_out/ab_{n}.json').read().strip().splitlines()[-1])
        print(n, 'ms/step', round(d['ms_per_step'], 3), 'rays/s', round(d['value']), 'e2e', round(d['e2e']['value']),
              {k: round(v['ms_total'] / v['calls'], 3) for k, v in d.get('ops_ms', {}).items()})
    except Exception as e:
        print(n, 'failed', e)
P
