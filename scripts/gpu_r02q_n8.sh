#!/bin/bash
# r02q, 8 GPUs: the full-step bench with the copy-engine form of the gradient exchange (compare with r02n/r02o: ncclAllReduce form, 9.03 / 8.14 ms)
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 --full-steps 20 --head-steps -1 --no-heads-f16 --exchange gather > $OUT/bench_r02q_gather.log 2>&1
echo "exit $?"; tail -1 $OUT/bench_r02q_gather.log > $OUT/bench_r02q_gather.json
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02q_gather.json').read())
    for k in ('full_step','full_step_config5'):
        f=d[k]; print(k,'ovl',round(f['ms_per_step'],3),'seq',round(f['ms_per_step_exchange_after_backward'],3),'ar',round(f['allreduce_ms'],3),'exposed',round(f['allreduce_exposed_ms'],3),f.get('multi_gpu_check'),f.get('exchange_stats'))
except Exception as e:
    print('no json',e); print(open('gpurun_out/bench_r02q_gather.log').read()[-2500:])
PY
