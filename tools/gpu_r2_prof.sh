#!/bin/bash
# tests touched this session + bench + ncu launch list of one steady-state step
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_config_sizes.py tests/test_gpu_tables.py -m gpu -x -q > gpurun_out/pytest_sub_$TAG.log 2>&1; echo "pytest subset rc=$?"
tail -4 gpurun_out/pytest_sub_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.log; tail -3 gpurun_out/bench_$TAG.err
python tools/profile_step.py > gpurun_out/prof_plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/launches_$TAG.csv | head -40
