#!/bin/bash
# compute-sanitizer evidence for the hand-rolled cross-CTA mbarrier protocols of the tcgen05 kernels:
# memcheck, racecheck and synccheck over the bf16 field forward (training variant, image stores) + both backward
# chains + the weight-gradient kernel, on a ragged sample count (128 * 9 + 77) and with the non-rigid MLP skipped.
# Target: tests/relu_bits_ab.py (seeded inputs, one forward + backward through the C ABI).  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for arm in "1229 0" "2048 1"; do
    set -- $arm
    log=gpurun_out/sanitizer_${TAG}_${tool}_P$1_ignore$2.log
    timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python tests/relu_bits_ab.py $1 $2 /tmp/san_$1_$2.npz > $log 2>&1
    echo "$tool P=$1 ignore=$2 exit $?" | tee -a $log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Internal|not supported|Error" $log | head -5
  done
done
