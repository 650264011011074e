#!/bin/bash
# the MOLCLR_* tuning switches exist only in the debug-switch build of the library
python -m molclr_b200.build --debug-switches > /dev/null && export MOLCLR_B200_LIB=$PWD/molclr_b200/libmolclr_b200_dbg.so
# timing experiments on the mixed compensated GEMM: debug bit 1 = no epilogue, 2 = no conversion, 4 = no correction MMAs
for D in 0 1 2 4 6 7; do echo "== MOLCLR_GEMM_DEBUG=$D"; MOLCLR_GEMM_DEBUG=$D CASE=fwd timeout 120 python tools/bench_gemm.py 2>&1 | grep -E "mixed|fwd1 x1   "; done
