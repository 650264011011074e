#!/bin/bash
# GCN on the fp16 three-product form: GPU suite; in-situ aggregation timing with and without programmatic dependent launch.
set -u
mkdir -p gpurun_out
TAG=${1:-r2e}
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
for flag in "" "--no-pdl"; do
  timeout 600 python bench.py --no-extra --no-cpu-baseline $flag > gpurun_out/bench_$TAG$flag.log 2> gpurun_out/bench_$TAG$flag.err; echo "bench $flag rc=$?"
  python tools/bench_line.py < gpurun_out/bench_$TAG$flag.log
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$TAG$flag.log').read().strip().splitlines()[-1])
print('gemm', d['roofline_gemm']['us_per_call'], 'agg', d['roofline']['us_per_launch'], d['roofline']['frac'], 'loss', d['run']['loss'], d['clocks'])
PY
done
timeout 600 python bench.py --model gcn --no-extra --no-cpu-baseline 2>/dev/null | python tools/bench_line.py
timeout 600 python bench.py --model gcn --precision tf32x3 --no-extra --no-cpu-baseline 2>/dev/null | python tools/bench_line.py
