#!/bin/bash
# bench line (twice: run-to-run spread) + ncu launch list of one steady-state step
set -u
mkdir -p gpurun_out
TAG=${1:-x}
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_$i.log 2> gpurun_out/bench_${TAG}_$i.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_${TAG}_$i.log').read().strip().splitlines()[-1]);print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'clk',d['clocks'])"
done
timeout 300 python tools/profile_step.py > gpurun_out/prof_plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
