#!/bin/bash
# r02x, 2 GPUs: both bench arms exactly as the driver launches them at N = 2 (short form), on the final tree
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29591 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/bench_r02x_ref_n2.log 2>&1; echo "ref exit $?"; tail -1 $OUT/bench_r02x_ref_n2.log | cut -c1-400
timeout 600 $TR --master-port 29592 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_r02x_n2.log 2>&1; echo "exit $?"; tail -1 $OUT/bench_r02x_n2.log > $OUT/bench_r02x_n2.json
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02x_n2.json').read())
    print('value',round(d['value'],1),'n_gpus',d['n_gpus'],'e2e',round(d['e2e']['value'],2),'clocks',d['clocks'])
    for k,v in d['config']['step_imgs_s'].items(): print(k,round(v['imgs_s'],1),'ms',round(v['ms_per_step'],3),'ar',v.get('allreduce_ms'),'exposed',v.get('allreduce_exposed_ms'),v.get('multi_gpu_check'))
except Exception as e:
    print('no json',e); print(open('gpurun_out/bench_r02x_n2.log').read()[-3000:])
PY
