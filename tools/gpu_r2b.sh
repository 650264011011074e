#!/bin/bash
# Round-2 second session: parity tests with programmatic dependent launch (and without, if they fail: bisects a launch-order race),
# A/B of the early TF32-pass issue of the compensated GEMM (debug-switch build: MOLCLR_GEMM_DEBUG bit 4 = 16 waits for the
# conversion first, as before), then the step bench (its `extra` object carries the PDL A/B).
set -u
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_$TAG.log
if [ $rc -ne 0 ]; then
  MOLCLR_B200_PDL=0 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}_nopdl.log 2>&1; echo "pytest (no PDL) rc=$?"; tail -3 gpurun_out/pytest_${TAG}_nopdl.log
fi
for dbg in 0 16; do
  MOLCLR_B200_LIB=molclr_b200/libmolclr_b200_dbg.so MOLCLR_GEMM_DEBUG=$dbg CASE=step timeout 300 python tools/bench_gemm.py 2>&1 | tail -3
done
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$TAG.err
python tools/bench_line.py < gpurun_out/bench_$TAG.log
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$TAG.log').read().strip().splitlines()[-1])
for k, v in d.get('extra', {}).items():
    print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
print('gemm', d['roofline_gemm']['us_per_call'], 'agg', d['roofline']['us_per_launch'], d['roofline']['frac'])
PY
