#!/bin/bash
# fp16 three-product compensated GEMM (compensate = 2 / precision "fp16x3"): unit tests, the whole GPU suite with it as the process
# default, microbenchmark against the TF32 + bf16 form, step bench in both modes.
set -u
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 300 python -m pytest tests/test_gpu_tables.py -x -q > gpurun_out/pytest_tables_$TAG.log 2>&1; echo "tables rc=$?"; tail -4 gpurun_out/pytest_tables_$TAG.log
CASE=step timeout 300 python tools/bench_gemm.py 2>&1 | tail -4
MOLCLR_B200_PRECISION=fp16x3 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_fp16x3_$TAG.log 2>&1; echo "pytest fp16x3 rc=$?"; tail -8 gpurun_out/pytest_fp16x3_$TAG.log
timeout 600 python bench.py --precision fp16x3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$TAG.err
python tools/bench_line.py < gpurun_out/bench_$TAG.log
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$TAG.log').read().strip().splitlines()[-1])
for k, v in d.get('extra', {}).items():
    print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
print('gemm', d['roofline_gemm']['us_per_call'], 'agg', d['roofline']['us_per_launch'], d['roofline']['frac'], 'loss', d['run']['loss'])
PY
