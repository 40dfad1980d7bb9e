#!/bin/bash
# Build libhumannerf_b200.so (sm_100a only) in-tree.  Used by __graft_entry__.build().
set -euo pipefail
cd "$(dirname "$0")"
SRC=humannerf_b200/csrc
OUT=humannerf_b200/lib
mkdir -p "$OUT" build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${HN_EXTRA_FLAGS:-} -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v"
objs=()
pids=()
for f in api rays patches motion_field composite field_fp32 mlp_tc_entry mlp_tc_2sm mlp_tc_ts mlp_tc_ts2 mlp_tc_bwd_entry mlp_tc_bwd_2sm tile_wgrad adam pose_fk lpips_head vol_decoder; do
  $NVCC $FLAGS -c $SRC/$f.cu -o build/$f.o > build/$f.log 2>&1 &
  pids+=($!)
  objs+=(build/$f.o)
done
fail=0
for p in "${pids[@]}"; do wait $p || fail=1; done
if [ $fail -ne 0 ]; then cat build/*.log; exit 1; fi
$NVCC -shared -o $OUT/libhumannerf_b200.so "${objs[@]}" -lcudart
echo "built $OUT/libhumannerf_b200.so"
