#!/bin/bash
# Round-2 profile captures: launch list of one steady-state step, ncu --set full of the aggregation kernel (roofline.traffic),
# and a metric subset (duration, tensor-pipe activity, DRAM / L2 bytes, occupancy, LSU shared-memory wavefronts) for every GEMM
# instance and the row-wise kernels of the step.  Reports are reduced to CSV on the box (gpurun_out/ is limited to 64 MiB).
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
rm -f gpurun_out/*.ncu-rep
python tools/profile_step.py > gpurun_out/prof_plain_$TAG.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gine_aggregate_fwd_tile -c 3 \
    -f -o gpurun_out/agg_$TAG python tools/profile_step.py > gpurun_out/ncu_agg_$TAG.log 2>&1; echo "ncu agg rc=$?"
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,\
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,\
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__inst_executed.sum,launch__registers_per_thread
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/kernels_$TAG.csv \
    python tools/profile_step.py > gpurun_out/ncu_kernels_$TAG.log 2>&1; echo "ncu kernels rc=$?"
ls -la gpurun_out/*_$TAG.* | head
