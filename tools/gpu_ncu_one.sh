#!/bin/bash
# ncu --set full (with source) of ONE launch of a GEMM instance inside a step:  tools/gpu_ncu_one.sh "<256, 0, 0, 1>" tag
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:gemm_tf32_kernel" -s ${3:-6} -c 1 \
    -f -o gpurun_out/one_$2 python tools/profile_step.py > gpurun_out/ncu_one_$2.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/one_$2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]: print(r[h.index('Kernel Name')][:90], r[h.index('gpu__time_duration.sum')])
"
ls -la gpurun_out/one_$2.ncu-rep
