#!/bin/bash
# r02l, 2 GPUs: exchange check (all parts, clean exit) + short 2-GPU bench with the overlapped full step, hard timeouts
OUT=gpurun_out
mkdir -p $OUT
echo "== exchange_check x2"
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/exchange_check.py 2>&1 | grep "exchange_check" | tee $OUT/exchange_check_n2_r02l.json | cut -c1-500
echo "exit: $?"
echo "== bench --gpus 2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 50 --warmup 5 --full-steps 5 --head-steps 10 > $OUT/bench_r02l_n2.log 2>&1
echo "exit: $?"; tail -1 $OUT/bench_r02l_n2.log > $OUT/bench_r02l_n2.json; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02l_n2.json').read())
    print('value', d['value'], 'n', d['n_gpus'])
    print(json.dumps(d['config']['step_imgs_s'], indent=1))
except Exception as e:
    print('no json:', e); print(open('gpurun_out/bench_r02l_n2.log').read()[-3000:])
PY
