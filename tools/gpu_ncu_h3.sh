#!/bin/bash
# ncu --set full (with source) of ONE launch of the second forward product (fp16 three-product GEMM, <160, 1, 8, 1>: the 4th GEMM launch
# of a step = layer 1, u -> z) inside a step;
# key metrics and the top stalled SASS instructions are reduced on the box (tools/ncu_hot.py).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:gemm_tf32_kernel" -s 3 -c 1 \
    -f -o gpurun_out/one_h3 python tools/profile_step.py > gpurun_out/ncu_one_h3.log 2>&1; echo "rc=$?"
python tools/ncu_hot.py gpurun_out/one_h3.ncu-rep 25 > gpurun_out/ncu_hot_h3.txt 2>&1; head -60 gpurun_out/ncu_hot_h3.txt
