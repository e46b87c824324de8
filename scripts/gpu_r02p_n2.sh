#!/bin/bash
# r02p, 2 GPUs: the copy-engine form of the gradient exchange (all-gather on the copy engines + local rank-ordered sum):
# correctness + what ran on the GPU (exchange_check), then the full-step bench in both forms.
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== slot sum tests (1 GPU)"
timeout 200 python -m pytest tests/test_exchange_gpu.py -x -q 2>&1 | tail -3
echo "== exchange_check, copy-engine form"
SAD_EXCHANGE_GATHER=1 NCCL_DEBUG=WARN timeout 150 $TR --master-port 29571 scripts/exchange_check.py > $OUT/exchange_check_r02p_gather.log 2>&1
echo "exit $?"; grep -v "^\[rank" $OUT/exchange_check_r02p_gather.log | tail -5 | cut -c1-3000
echo "== bench, copy-engine form"
timeout 300 $TR --master-port 29572 bench.py --gpus 2 --steps 20 --warmup 5 --full-steps 20 --head-steps -1 --no-heads-f16 --exchange gather > $OUT/bench_r02p_gather.log 2>&1
echo "exit $?"; tail -1 $OUT/bench_r02p_gather.log > $OUT/bench_r02p_gather.json
echo "== bench, ncclAllReduce form"
timeout 300 $TR --master-port 29573 bench.py --gpus 2 --steps 20 --warmup 5 --full-steps 20 --head-steps -1 --no-heads-f16 --exchange allreduce > $OUT/bench_r02p_allreduce.log 2>&1
echo "exit $?"; tail -1 $OUT/bench_r02p_allreduce.log > $OUT/bench_r02p_allreduce.json
python - <<'PY'
import json
for tag in ('gather','allreduce'):
    try:
        d=json.loads(open('gpurun_out/bench_r02p_%s.json'%tag).read())
        for k in ('full_step','full_step_config5'):
            f=d[k]; print(tag,k,'ovl',round(f['ms_per_step'],3),'seq',round(f['ms_per_step_exchange_after_backward'],3),'ar',round(f['allreduce_ms'],3),'exposed',round(f['allreduce_exposed_ms'],3),f.get('multi_gpu_check'),f.get('exchange_stats'))
    except Exception as e:
        print(tag,'no json',e); print(open('gpurun_out/bench_r02p_%s.log'%tag).read()[-2500:])
PY
