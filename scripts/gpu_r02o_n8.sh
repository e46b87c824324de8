#!/bin/bash
# r02o, 8 GPUs: does the exchange overlap the backward pass when NCCL's CTAs are smaller / fewer?  Two short full-step benches.
OUT=gpurun_out
mkdir -p $OUT
run() {  # $1 tag, rest: env assignments
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 20 --warmup 5 --full-steps 20 --head-steps -1 --no-heads-f16 > $OUT/bench_r02o_$tag.log 2>&1
  echo "exit $?"; tail -1 $OUT/bench_r02o_$tag.log > $OUT/bench_r02o_$tag.json
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/bench_r02o_%s.json'%tag).read())
    for k in ('full_step','full_step_config5'):
        f=d[k]; print(tag,k,'ovl',round(f['ms_per_step'],3),'seq',round(f['ms_per_step_exchange_after_backward'],3),'ar',round(f['allreduce_ms'],3),'exposed',round(f['allreduce_exposed_ms'],3),'busbw',round(f['allreduce_busbw_gbs'],1))
except Exception as e:
    print(tag,'no json',e); print(open('gpurun_out/bench_r02o_%s.log'%tag).read()[-1500:])
PY
}
PORT=29561; run ctas12 SAD_EXCHANGE_MAX_CTAS=12
PORT=29562; run ctas16 SAD_EXCHANGE_MAX_CTAS=16
PORT=29563; run ctas20 SAD_EXCHANGE_MAX_CTAS=20
