#!/bin/bash
# One gpurun call that refreshes the ncu evidence under gpurun_out/ (copy the summaries to profiles/):
#   1. launch list of the default bench command (plain run first, must exit 0),
#   2. ncu --set full of the bf16 training kernels + adam + motion field + pose front end,
#   3. ncu --set full of the forward render kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-r01f}
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-render --no-skip-leg > gpurun_out/prof_plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/prof_plain_bench.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-render --no-skip-leg > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
python scripts/profile_render.py --rays 6144 --train > gpurun_out/prof_plain_train.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on \
    -k "regex:field_tc|field_ts|tile_wgrad_kernel|adam_kernel|motion_field|pose_fk|pack_corners|composite|lpips_head|unpack_corners|select_patches|deconv_gemm" -s 0 -c 30 \
    -o gpurun_out/prof_train_${TAG} -f python scripts/profile_render.py --rays 6144 --train > gpurun_out/ncu_train.log 2>&1
echo "ncu train exit $?"
python scripts/profile_render.py --rays 32768 > gpurun_out/prof_plain_render.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:field_tc|field_ts|motion_field_fwd|composite_fwd" -s 0 -c 4 \
    -o gpurun_out/prof_render_${TAG} -f python scripts/profile_render.py --rays 32768 > gpurun_out/ncu_render.log 2>&1
echo "ncu render exit $?"
ls -la gpurun_out/*${TAG}*
