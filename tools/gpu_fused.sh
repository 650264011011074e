#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-f}
for sw in 0 1; do
echo "== MOLCLR_NTX_SWAP_LBO=$sw"
MOLCLR_NTX_SWAP_LBO=$sw timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -k "ntxent" 2>&1 | tail -12
done
