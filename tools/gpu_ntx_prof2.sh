#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-p}
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ntx_bwd_fused -c 1 -f -o gpurun_out/ntx_fused_$TAG python tools/profile_ntx.py > gpurun_out/ntxprof_f_$TAG.log 2>&1; echo "full rc=$?"
