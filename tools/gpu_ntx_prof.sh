#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-p}
timeout 300 python tools/profile_ntx.py > gpurun_out/ntxprof_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail gpurun_out/ntxprof_plain_$TAG.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/ntx_launches_$TAG.csv python tools/profile_ntx.py > gpurun_out/ntxprof_l_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tf32_kernel -c 3 -f -o gpurun_out/ntx_full_$TAG python tools/profile_ntx.py > gpurun_out/ntxprof_f_$TAG.log 2>&1; echo "full rc=$?"
