#!/bin/bash
# fp16x3 as the default precision: whole GPU suite, smoke, per-parameter parity report, knock-outs of the compensated products
# (debug-switch build: 1 = no epilogue, 2 = no conversion, 4 = no correction MMAs), step bench with the CPU baseline.
set -u
mkdir -p gpurun_out
TAG=${1:-r2d}
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tests/parity_report.py --out gpurun_out/parity_fp16x3_$TAG.json > gpurun_out/parity_fp16x3_$TAG.log 2>&1; echo "parity rc=$?"; grep "^==" gpurun_out/parity_fp16x3_$TAG.log
for dbg in 0 1 2 4 7; do
  MOLCLR_B200_LIB=molclr_b200/libmolclr_b200_dbg.so MOLCLR_GEMM_DEBUG=$dbg CASE="step fwd" timeout 300 python tools/bench_gemm.py 2>&1 | tail -5
done
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$TAG.err
python tools/bench_line.py < gpurun_out/bench_$TAG.log
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$TAG.log').read().strip().splitlines()[-1])
for k, v in d.get('extra', {}).items():
    print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
print('gemm', d['roofline_gemm']['us_per_call'], 'agg', d['roofline']['us_per_launch'], d['roofline']['frac'], 'loss', d['run']['loss'])
PY
