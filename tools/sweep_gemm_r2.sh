#!/bin/bash
# Round-2 knock-out timings of the product-path GEMMs (debug-switch build): bit 1 = no epilogue, 2 = no conversion, 4 = no correction MMAs
python -m molclr_b200.build --debug-switches > /dev/null && export MOLCLR_B200_LIB=$PWD/molclr_b200/libmolclr_b200_dbg.so
for D in 0 1 2 4 6 7; do echo "== MOLCLR_GEMM_DEBUG=$D"; MOLCLR_GEMM_DEBUG=$D CASE="step" timeout 120 python tools/bench_gemm.py 2>&1 | grep "step"; done
for D in 0 1; do echo "== MOLCLR_GEMM_DEBUG=$D"; MOLCLR_GEMM_DEBUG=$D CASE="d" timeout 120 python tools/bench_gemm.py 2>&1 | grep -E "dU masked\+colsum|dX   x1   \[M,600\]|dW "; done
