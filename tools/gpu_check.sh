#!/bin/bash
# One GPU-box call: parity tests, bench line, ncu launch list of one steady-state step, full ncu capture of
# the aggregation kernel and of the masked backward GEMM.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.log
if [ "${NCU:-1}" = "1" ]; then
python tools/profile_step.py > gpurun_out/prof_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gine_aggregate_fwd -s 2 -c 2 \
    -f -o gpurun_out/agg_$TAG python tools/profile_step.py > gpurun_out/ncu_agg_$TAG.log 2>&1
echo "ncu agg rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tf32_kernel -s 4 -c 6 \
    -f -o gpurun_out/gemm_$TAG python tools/profile_step.py > gpurun_out/ncu_gemm_$TAG.log 2>&1
echo "ncu gemm rc=$?"
fi
