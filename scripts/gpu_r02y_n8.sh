#!/bin/bash
# r02y, 8 GPUs: the full-step objects on the final tree, configs[4] with both heads in fp16 included (bs = 8 over 8 GPUs, 1 image per GPU)
OUT=gpurun_out
mkdir -p $OUT
timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29595 bench.py --gpus 8 --steps 10 --warmup 3 --full-steps 10 --head-steps -1 --e2e-steps 3 > $OUT/bench_r02y_n8.log 2>&1
echo "exit $?"; tail -1 $OUT/bench_r02y_n8.log > $OUT/bench_r02y_n8.json
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02y_n8.json').read())
    print('value',round(d['value'],1),'n_gpus',d['n_gpus'],'e2e',round(d['e2e']['value'],2))
    for k,v in d['config']['step_imgs_s'].items(): print(k,round(v['imgs_s'],1),'ms',round(v['ms_per_step'],3),'ar',v.get('allreduce_ms'),'exposed',v.get('allreduce_exposed_ms'),v.get('multi_gpu_check'))
except Exception as e:
    print('no json',e); print(open('gpurun_out/bench_r02y_n8.log').read()[-2000:])
PY
