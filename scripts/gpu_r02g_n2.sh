#!/bin/bash
# r02g, 2 GPUs: native exchange check (nccl_ops_test.py style) + a short 2-GPU bench with the overlapped full step
OUT=gpurun_out
mkdir -p $OUT
echo "== exchange_check x2"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/exchange_check.py 2>&1 | tail -5 | tee $OUT/exchange_check_n2_r02g.json
echo "== bench --gpus 2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 --full-steps 10 --head-steps 30 2>&1 | tail -3 | tee $OUT/bench_r02g_n2.json | cut -c1-600
