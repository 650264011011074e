#!/bin/bash
# 2 GPUs: real-NCCL parity test and the weak-scaling bench at N = 2 (and N = 1 on the same box) with programmatic dependent launch
# and the fp16x3 default.
set -u
mkdir -p gpurun_out
TAG=${1:-r2f}
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q > gpurun_out/pytest_dist_$TAG.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/pytest_dist_$TAG.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/scale_${TAG}_2.log 2> gpurun_out/scale_${TAG}_2.err; echo "bench N=2 rc=$?"; tail -c 400 gpurun_out/scale_${TAG}_2.err
python tools/bench_line.py < gpurun_out/scale_${TAG}_2.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/scale_${TAG}_1.log 2> gpurun_out/scale_${TAG}_1.err; echo "bench N=1 rc=$?"
python tools/bench_line.py < gpurun_out/scale_${TAG}_1.log
