#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-i}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
for B in 4096 512; do BATCH=$B timeout 300 python tools/cpu_overhead.py 2>&1 | tail -2; done
