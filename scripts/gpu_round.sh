#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, then (only if the plain bench exited 0)
# the ncu launch list.  Everything is logged under gpurun_out/.  Each stage has its own timeout so a
# hung kernel cannot run into gpurun's limit.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest (fp32 + rays + composite + motion field) =="
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "not bf16 and not full_size" > gpurun_out/pytest_fp32.log 2>&1
echo "exit $?" >> gpurun_out/pytest_fp32.log
tail -5 gpurun_out/pytest_fp32.log
echo "== pytest (bf16 tcgen05) =="
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "bf16 or full_size" > gpurun_out/pytest_bf16.log 2>&1
echo "exit $?" >> gpurun_out/pytest_bf16.log
tail -5 gpurun_out/pytest_bf16.log
echo "== smoke =="
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
echo "== bench =="
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
rc=$?
echo "bench exit $rc"
tail -c 3000 gpurun_out/bench.log
tail -5 gpurun_out/bench.err
if [ $rc -eq 0 ] && [ "${HN_NCU:-1}" = "1" ]; then
  echo "== ncu launch list (bench) =="
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  echo "ncu exit $?"
  echo "== ncu full: tcgen05 field kernel =="
  python scripts/profile_render.py --rays 32768 > gpurun_out/prof_plain_render.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_tc_kernel -s 1 -c 1 \
      -o gpurun_out/prof_field_tc -f python scripts/profile_render.py --rays 32768 > gpurun_out/ncu_tc.log 2>&1
  echo "ncu exit $?"
  echo "== ncu full: motion field + sgemm (train) =="
  python scripts/profile_render.py --rays 6144 --train > gpurun_out/prof_plain_train.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:motion_field|sgemm_kernel" -s 40 -c 12 \
      -o gpurun_out/prof_train -f python scripts/profile_render.py --rays 6144 --train > gpurun_out/ncu_train.log 2>&1
  echo "ncu exit $?"
fi
