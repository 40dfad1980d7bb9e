#!/bin/bash
# One short gpurun call for an A/B of a training-kernel switch:
#   bash scripts/gpu_ab.sh ENV_VAR VALUE_A VALUE_B [pytest -k expression]
# runs the selected GPU tests with ENV_VAR=VALUE_A, then the train-step bench with VALUE_A and with VALUE_B.
set -u
VAR=${1:-HN_TC_LEADER_STORES}
A=${2:-4}
Bv=${3:-1}
SEL=${4:-"relu_mask or field_bf16_backward or network_train_bf16 or graphed_train_step or skipping"}
mkdir -p gpurun_out
env $VAR=$A timeout 400 python -m pytest tests -q -x -m gpu -k "$SEL" > gpurun_out/ab_pytest.log 2>&1
echo "pytest ($VAR=$A) exit $?" | tee -a gpurun_out/ab_pytest.log
tail -4 gpurun_out/ab_pytest.log
B="--steps 30 --warmup 5 --no-cpu-baseline --no-render --no-skip-leg"
for v in $A $Bv; do
  env $VAR=$v timeout 200 python bench.py $B > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; echo "bench $VAR=$v exit $?"
done
python - $A $Bv <<'P'
import json, sys
for n in sys.argv[1:]:
    try:
        d = json.loads(open(f'gpurun_out/ab_{n}.json').read().strip().splitlines()[-1])
        print(n, 'ms/step', round(d['ms_per_step'], 3), 'rays/s', round(d['value']), 'e2e', round(d['e2e']['value']),
              {k: round(v['ms_total'] / v['calls'], 3) for k, v in d.get('ops_ms', {}).items() if k.startswith('field')})
    except Exception as e:
        print(n, 'failed', e)
P
