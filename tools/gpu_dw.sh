#!/bin/bash
set -u
mkdir -p gpurun_out
for d in 0 8 16; do echo "== MOLCLR_GEMM_DEBUG=$d"; MOLCLR_GEMM_DEBUG=$d CASE=dW timeout 300 python tools/bench_gemm.py 2>&1 | tee -a gpurun_out/dw_$1.log; done
