#!/bin/bash
# One GPU-box call of round 2: new tests first (fail fast), then the whole GPU suite, a short bench line and the parity report.
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 600 python -m pytest tests/test_gpu_tables.py tests/test_gpu_config_sizes.py -m gpu -x -q > gpurun_out/pytest_new_$TAG.log 2>&1; echo "pytest new rc=$?"
tail -15 gpurun_out/pytest_new_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest all rc=$?"
tail -15 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.log; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python tests/parity_report.py --batches 128,512,4096 --out gpurun_out/parity_$TAG.json > gpurun_out/parity_$TAG.log 2>&1; echo "parity rc=$?"
grep -v "^   " gpurun_out/parity_$TAG.log | tail -8
