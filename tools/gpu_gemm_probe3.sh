#!/bin/bash
# the MOLCLR_* tuning switches exist only in the debug-switch build of the library
python -m molclr_b200.build --debug-switches > /dev/null && export MOLCLR_B200_LIB=$PWD/molclr_b200/libmolclr_b200_dbg.so
mkdir -p gpurun_out
TAG=${1:-x}
MOLCLR_GEMM_DEBUG=1 timeout 120 python tools/bench_gemm.py > gpurun_out/gemm_bench_noepi_$TAG.log 2>&1; cat gpurun_out/gemm_bench_noepi_$TAG.log
MOLCLR_GEMM_PAIR=0 timeout 120 python tools/bench_gemm.py > gpurun_out/gemm_bench_1cta_$TAG.log 2>&1; cat gpurun_out/gemm_bench_1cta_$TAG.log
for c in "fwd1 x1   [M,300]" "fwd1 x3"; do
  n=$(echo "$c" | tr -d ' [],' )
  CASE="$c" ITERS=2 python tools/bench_gemm.py > gpurun_out/pp_$n.log 2>&1 &&
  CASE="$c" ITERS=2 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 3 -c 1 -f -o gpurun_out/g_${n}_$TAG python tools/bench_gemm.py > gpurun_out/ncu_$n.log 2>&1
  echo "$c rc=$?"
done
