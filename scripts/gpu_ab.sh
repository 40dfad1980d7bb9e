#!/bin/bash
# One short gpurun call: the training-path tests that cover the 1-bit ReLU mask images, then the train-step bench with
# the mask images (default), with the activation read (HN_TC_RELU_BITS=0) and at the 20 x 20 patch size of
# zju_mocap/387/single_gpu.yaml.  Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_tc_train.py -q -x -m gpu \
    -k "relu_mask or field_bf16_backward or network_train_bf16 or graphed_train_step or skipping" > gpurun_out/ab_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/ab_pytest.log
tail -5 gpurun_out/ab_pytest.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-render --no-skip-leg"
timeout 300 python bench.py $B > gpurun_out/ab_bits.json 2> gpurun_out/ab_bits.err; echo "bench bits exit $?"
HN_TC_RELU_BITS=0 timeout 300 python bench.py $B > gpurun_out/ab_readh.json 2> gpurun_out/ab_readh.err; echo "bench read-h exit $?"
timeout 300 python bench.py $B --patch 20 > gpurun_out/ab_p20.json 2> gpurun_out/ab_p20.err; echo "bench p20 exit $?"
python - <<'P'
import json
for n in ('bits', 'readh', 'p20'):
    try:
        d = json.loads(open(f'gpurun_out/ab_{n}.json').read().strip().splitlines()[-1])
        print(n, 'ms/step', round(d['ms_per_step'], 3), 'rays/s', round(d['value']), 'e2e', round(d['e2e']['value']),
              {k: round(v['ms_total'] / v['calls'], 3) for k, v in d.get('ops_ms', {}).items()})
    except Exception as e:
        print(n, 'failed', e)
P
