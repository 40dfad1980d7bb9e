#!/bin/bash
# ncu --set full of the volume decoder kernels (one forward + backward of scripts/time_decoder.py)
mkdir -p gpurun_out
python scripts/time_decoder.py hn tf32 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:deconv_gemm|deconv_finish|softmax_prior|column_sum|linear_" -s 0 -c 40 \
    -o gpurun_out/prof_decoder -f python scripts/time_decoder.py hn tf32 > gpurun_out/ncu_decoder.log 2>&1
echo "ncu exit $?"
