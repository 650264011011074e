#!/bin/bash
# the MOLCLR_* tuning switches exist only in the debug-switch build of the library
python -m molclr_b200.build --debug-switches > /dev/null && export MOLCLR_B200_LIB=$PWD/molclr_b200/libmolclr_b200_dbg.so
# First GPU call of the next session: what the last one could not run any more.
#  1. the motif-attention drop-in's parity test with its xfail marker ignored (shows the traceback if it still fails)
#  2. the general (running-maximum) NT-Xent forms, taken for tau < 0.045: microbenchmark with the bounded forms switched off
#  3. the whole GPU suite and a bench line
set -u
mkdir -p gpurun_out
TAG=${1:-next}
timeout 300 python -m pytest tests/test_gpu_zz_pending.py -m gpu -q --runxfail 2>&1 | tail -40 | tee gpurun_out/motif_$TAG.log
MOLCLR_NTX_NOBOUND=1 timeout 300 python tools/bench_ntxent.py 2>&1 | tee gpurun_out/ntx_nobound_$TAG.log
timeout 900 python -m pytest tests -m gpu -q -rxX > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python tools/bench_line.py
