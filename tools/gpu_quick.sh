#!/bin/bash
# tests + GEMM microbench + step bench (no ncu)
mkdir -p gpurun_out
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
python tools/bench_gemm.py > gpurun_out/gemm_bench_$TAG.log 2>&1; echo rc=$?; cat gpurun_out/gemm_bench_$TAG.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_$TAG.log').read().strip().splitlines()[-1]);print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'loss',d['config']['loss'])"
