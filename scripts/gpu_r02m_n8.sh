#!/bin/bash
# r02m, 8 GPUs: short bench with the overlapped exchange + multi_gpu_check, hard timeouts (charged 8x)
OUT=gpurun_out
mkdir -p $OUT
echo "== exchange_check x8"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 scripts/exchange_check.py > $OUT/exchange_check_n8_r02m.log 2>&1
echo "exit: $?"; grep "exchange_check" $OUT/exchange_check_n8_r02m.log | tee $OUT/exchange_check_n8_r02m.json | cut -c1-700; grep -i "error\|assert" $OUT/exchange_check_n8_r02m.log | head -5
echo "== bench --gpus 8"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 10 --full-steps 10 --head-steps 20 > $OUT/bench_r02m_n8.log 2>&1
echo "exit: $?"; tail -1 $OUT/bench_r02m_n8.log > $OUT/bench_r02m_n8.json; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02m_n8.json').read())
    print('value', d['value'], 'n', d['n_gpus'], 'e2e', d['e2e']['value'], 'copy_only', d['e2e']['copy_only'])
    print(json.dumps(d['config']['step_imgs_s'], indent=1))
    for k in ('full_step','full_step_config5','full_step_config5_heads_f16'):
        f=d[k]; print(k,'seq',f['ms_per_step_exchange_after_backward'],'ar',f['allreduce_ms'],'busbw',f['allreduce_busbw_gbs'])
except Exception as e:
    print('no json:', e); print(open('gpurun_out/bench_r02m_n8.log').read()[-3000:])
PY
