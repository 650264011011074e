#!/bin/bash
# End-of-session GPU call: parity tests, smoke, both bench arms, ncu launch list of one step, --set full captures of the
# aggregation kernel (roofline traffic) and of the NT-Xent kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-final}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; python tools/bench_line.py < gpurun_out/bench_$TAG.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; tail -c 400 gpurun_out/bench_ref_$TAG.log
timeout 600 python bench.py --precision tf32 --no-cpu-baseline 2>/dev/null | python tools/bench_line.py
timeout 600 python bench.py --model gcn --no-cpu-baseline 2>/dev/null | python tools/bench_line.py
timeout 300 python tools/profile_step.py > gpurun_out/prof_plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gine_aggregate_fwd -s 2 -c 1 \
    -f -o gpurun_out/agg_$TAG python tools/profile_step.py > gpurun_out/ncu_agg_$TAG.log 2>&1
echo "ncu agg rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ntx_bwd_fused|gemm_tf32_kernel<256, 0, 7" -c 2 \
    -f -o gpurun_out/ntx_$TAG python tools/profile_ntx.py > gpurun_out/ncu_ntx_$TAG.log 2>&1
echo "ncu ntx rc=$?"
