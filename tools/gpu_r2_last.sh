#!/bin/bash
# Last call of round 2: GPU suite, bench (forward products timed separately, range check every 8th forward), parity report of the
# default precision at 128 / 512 / 4096 pairs.
set -u
mkdir -p gpurun_out
TAG=${1:-r2last}
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; python tools/bench_line.py < gpurun_out/bench_$TAG.log
timeout 600 python tests/parity_report.py --batches 128,512,4096 --out gpurun_out/parity_fp16x3_$TAG.json > gpurun_out/parity_fp16x3_$TAG.log 2>&1; echo "parity rc=$?"; grep "^==" gpurun_out/parity_fp16x3_$TAG.log
