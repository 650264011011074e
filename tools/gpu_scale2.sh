#!/bin/bash
set -u
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 15 --warmup 5 "$@" \
      > gpurun_out/s8_$tag.log 2> gpurun_out/s8_$tag.err; python -c "import json; l=json.loads(open('gpurun_out/s8_$tag.log').read().strip().split('\n')[-1]); print('$tag', l['value'], l['ms_per_step'])"; }
run default
run nooverlap --dp-bucket 0
run perlayer --dp-bucket 300000
run same --same-batches
run local_nooverlap --local-negatives --dp-bucket 0
run local_same --local-negatives --same-batches
