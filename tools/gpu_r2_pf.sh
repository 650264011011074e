#!/bin/bash
# L2 prefetch of the A operand in the row GEMMs: GPU suite, then the step with and without it (debug-switch build, bit 32 = off).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_pf.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_pf.log
for dbg in 0 32; do
  MOLCLR_B200_LIB=molclr_b200/libmolclr_b200_dbg.so MOLCLR_GEMM_DEBUG=$dbg timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/bench_pf_$dbg.log 2>/dev/null
  python tools/bench_line.py < gpurun_out/bench_pf_$dbg.log
  python -c "
import json; d=json.loads(open('gpurun_out/bench_pf_$dbg.log').read().strip().splitlines()[-1]); print({k: round(v,1) for k,v in d['roofline_gemm']['us_per_call'].items()}, d['clocks'])"
done
