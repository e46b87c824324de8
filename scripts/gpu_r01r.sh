#!/bin/bash
# r01r: the default bench line (with head_step_f16) on N GPUs as the driver launches it.
TAG=${1:-r01r}
N=${2:-1}
OUT=gpurun_out
mkdir -p $OUT
if [ "$N" = "1" ]; then
  timeout 500 python bench.py 2>&1 | tail -1 | tee $OUT/bench_${TAG}.json | cut -c1-300
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N 2>&1 | tail -1 | tee $OUT/bench_${TAG}_n${N}.json | cut -c1-300
fi
