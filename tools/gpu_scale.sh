#!/bin/bash
# 8-GPU box: real-NCCL parity test (4 ranks), then the weak-scaling bench at N = 8, 4, 2, 1 (+ 8 GPUs with local negatives).
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q > gpurun_out/pytest_dist_$TAG.log 2>&1; echo "dist rc=$?"; tail -2 gpurun_out/pytest_dist_$TAG.log
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 \
      > gpurun_out/scale_${TAG}_$N.log 2> gpurun_out/scale_${TAG}_$N.err; echo "bench N=$N rc=$?"
  python -c "import json,sys; l=json.loads(open('gpurun_out/scale_${TAG}_$N.log').read().strip().split('\n')[-1]); print($N, l['value'], l['ms_per_step'], l['e2e']['value'])"
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/scale_${TAG}_1.log 2> gpurun_out/scale_${TAG}_1.err; echo "bench N=1 rc=$?"
python -c "import json,sys; l=json.loads(open('gpurun_out/scale_${TAG}_1.log').read().strip().split('\n')[-1]); print(1, l['value'], l['ms_per_step'], l['e2e']['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 --local-negatives \
      > gpurun_out/scale_${TAG}_8local.log 2> gpurun_out/scale_${TAG}_8local.err; echo "bench N=8 local negatives rc=$?"
python -c "import json,sys; l=json.loads(open('gpurun_out/scale_${TAG}_8local.log').read().strip().split('\n')[-1]); print('8 local', l['value'], l['ms_per_step'])"
