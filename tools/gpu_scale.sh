#!/bin/bash
# N-GPU weak-scaling bench lines (launched as the driver does) -- usage: gpurun --gpus N -- 'bash tools/gpu_scale.sh TAG "1 2"'
set -u
mkdir -p gpurun_out
TAG=${1:-s}
for n in ${2:-1 2}; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${TAG}_$n.log 2> gpurun_out/scale_${TAG}_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline ${3:-} > gpurun_out/scale_${TAG}_$n.log 2> gpurun_out/scale_${TAG}_$n.err
  fi
  echo "n=$n rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/scale_${TAG}_$n.log').read().strip().splitlines()[-1]);print('n',d['n_gpus'],'ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'loss',d['config'].get('loss'))" || tail -20 gpurun_out/scale_${TAG}_$n.err
done
