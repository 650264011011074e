#!/bin/bash
# the MOLCLR_* tuning switches exist only in the debug-switch build of the library
python -m molclr_b200.build --debug-switches > /dev/null && export MOLCLR_B200_LIB=$PWD/molclr_b200/libmolclr_b200_dbg.so
# One GPU-box call for the fp16 NT-Xent path: its parity tests first, the NT-Xent microbenchmark (fp16 vs TF32 operands, stripe
# widths), then the whole GPU suite and a short bench line.  Every stage under its own timeout.
set -u
mkdir -p gpurun_out
TAG=${1:-ntx}
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k ntxent > gpurun_out/pytest_ntx_$TAG.log 2>&1; echo "pytest ntxent rc=$?"; tail -5 gpurun_out/pytest_ntx_$TAG.log
timeout 300 python tools/bench_ntxent.py > gpurun_out/ntx_bench_$TAG.log 2>&1; echo "bench_ntxent rc=$?"; cat gpurun_out/ntx_bench_$TAG.log
MOLCLR_NTX_STRIPE16=2048 NTX_MODES=1 timeout 300 python tools/bench_ntxent.py > gpurun_out/ntx_bench_s2048_$TAG.log 2>&1; echo "bench_ntxent(2048) rc=$?"; cat gpurun_out/ntx_bench_s2048_$TAG.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_$TAG.log
