#!/bin/bash
# the MOLCLR_* tuning switches exist only in the debug-switch build of the library
python -m molclr_b200.build --debug-switches > /dev/null && export MOLCLR_B200_LIB=$PWD/molclr_b200/libmolclr_b200_dbg.so
# GEMM microbenchmarks (with and without the epilogue) + full ncu captures of the backward GEMM variants.
mkdir -p gpurun_out
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
python tools/bench_gemm.py > gpurun_out/gemm_bench_$TAG.log 2>&1; echo rc=$?
MOLCLR_GEMM_DEBUG=1 python tools/bench_gemm.py > gpurun_out/gemm_bench_noepi_$TAG.log 2>&1; echo rc=$?
python tools/profile_step.py > gpurun_out/pp.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:gemm_tf32_kernel<160, 0, 0>" -c 3 \
    -f -o gpurun_out/gemm_plain_$TAG python tools/profile_step.py > gpurun_out/ncu_gemm_plain_$TAG.log 2>&1
echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:gemm_tf32_kernel<160, 0, 4>" -c 2 \
    -f -o gpurun_out/gemm_atomic_$TAG python tools/profile_step.py > gpurun_out/ncu_gemm_atomic_$TAG.log 2>&1
echo "ncu rc=$?"
