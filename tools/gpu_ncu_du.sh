#!/bin/bash
# Source-level ncu capture of ONE launch of the masked g_u product from the microbenchmark; SASS page reduced to CSV on the box.
set -u
mkdir -p gpurun_out
CASE="${1:-dU masked+colsum}" ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm_tf32_kernel" -s 3 -c 1 \
    -f -o gpurun_out/one_${2:-du2} python tools/bench_gemm.py > gpurun_out/ncu_one_${2:-du2}.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/one_${2:-du2}.ncu-rep --page source --csv --print-source sass > gpurun_out/one_${2:-du2}_sass.csv 2>/dev/null
python tools/ncu_hot.py gpurun_out/one_${2:-du2}.ncu-rep 40
ls -la gpurun_out/one_${2:-du2}*
