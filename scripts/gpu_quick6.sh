#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_train.py tests/test_gpu_cfg2_step.py -q -m gpu -x -k "not trajectory" 2>&1 | tail -3
bash scripts/gpu_exp.sh "$@"
