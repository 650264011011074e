#!/bin/bash
# A/B of two builds of the library on the GEMM microbenchmark (same box, alternating)
set -u
mkdir -p gpurun_out
TAG=${1:-ab}
for rep in 1 2; do
for lib in libmolclr_prev.so libmolclr_b200.so; do
echo "== $lib (rep $rep)"
MOLCLR_B200_LIB=$PWD/molclr_b200/$lib CASE=d timeout 300 python tools/bench_gemm.py 2>&1 | tee -a gpurun_out/ab_$TAG.log
done; done
