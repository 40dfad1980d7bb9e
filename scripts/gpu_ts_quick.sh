#!/bin/bash
# quick GPU check of the tcgen05 forward kernels: timings of the render / training variants, the bf16 field tests, a short bench
mkdir -p gpurun_out
python scripts/ts_trace.py > gpurun_out/ts_trace.log 2>&1
python scripts/ts_trace.py 786432 train > gpurun_out/ts_trace_train.log 2>&1
python scripts/ts_trace.py 786432 train >> gpurun_out/ts_trace_train.log 2>&1
head -3 gpurun_out/ts_trace.log; grep forward gpurun_out/ts_trace_train.log; grep -c "timed out" gpurun_out/ts_trace_train.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc_train.py -q -m gpu -x -k "bf16 or relu_mask or graphed or skipping" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-eager-gpu-baseline --no-skip-leg > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'], 3), 'e2e ms', round(d['e2e']['ms_per_step'], 3), 'roofline', round(d['roofline']['frac'], 4),
      {k: round(v['ms_total'] / v['calls'], 3) for k, v in d['ops_ms'].items() if k.startswith('field')})
r = d.get('render')
if r: print('render ms/frame', round(r['ms_per_frame'], 2), r['ops_ms_per_frame'], 'mlp frac', round(r['roofline_mlp']['frac'], 3))
P
