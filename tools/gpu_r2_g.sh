#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-g}
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
for B in 4096 512; do BATCH=$B timeout 300 python tools/cpu_overhead.py 2>&1 | tail -2; done
ONLY_TABLES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:colsum -c 4 -f -o gpurun_out/tables_$TAG python tools/bench_rowwise.py > gpurun_out/ncu_tables_$TAG.log 2>&1; echo "ncu tables rc=$?"
