#!/bin/bash
# r02f: fused kernel v5 (barrier words carry the normaliser; static deals; exact loss sums)
OUT=gpurun_out
mkdir -p $OUT
echo "== distill tests"; timeout 900 python -m pytest tests/test_distill_gpu.py tests/test_operator_boundary_gpu.py -q -x 2>&1 | tail -15 | tee $OUT/pytest_distill_r02f.log
echo "== exchange + full step tests"; timeout 900 python -m pytest tests/test_exchange_gpu.py tests/test_sgd_gpu.py tests/test_full_step_gpu.py -q 2>&1 | tail -25 | tee $OUT/pytest_exchange_r02f.log
echo "== bench (short: headline + e2e only)"
timeout 600 python bench.py --steps 300 --warmup 20 --head-steps -1 --full-steps -1 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_r02f_short.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value',d['value'],'ms',d['ms_per_step'],'kernel_ms',r['kernel_ms'],'median',r['kernel_ms_median'],'min',r['kernel_ms_min'],'frac',r['frac'])"
echo "== stamps"
SAD_FUSED_DEBUG=8 timeout 300 python scripts/fused_stamps.py > $OUT/fused_stamps_r02f.txt 2>&1; tail -6 $OUT/fused_stamps_r02f.txt
echo "== ncu full fused kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:distill_fused_kernel -s 5 -c 2 -o $OUT/prof_fused_r02f -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --head-steps -1 --full-steps -1 --e2e-steps 1 > $OUT/ncu_full_r02f.log 2>&1
ls -la $OUT | tail -4
