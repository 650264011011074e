#!/bin/bash
# End of round 2: parity tests, smoke, both bench arms, then the ncu captures of one steady-state step (tools/gpu_r2_ncu.sh:
# launch list, --set full of the aggregation kernel, metric subset per kernel).
set -u
mkdir -p gpurun_out
TAG=${1:-r2final2}
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; python tools/bench_line.py < gpurun_out/bench_$TAG.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; tail -c 300 gpurun_out/bench_ref_$TAG.log
timeout 300 python bench.py --precision tf32x3 --no-pdl --no-extra --no-cpu-baseline 2>/dev/null > gpurun_out/bench_${TAG}_tf32x3_nopdl.log; python tools/bench_line.py < gpurun_out/bench_${TAG}_tf32x3_nopdl.log
bash tools/gpu_r2_ncu.sh $TAG
python tools/launch_summary.py gpurun_out/launches_$TAG.csv | head -30
