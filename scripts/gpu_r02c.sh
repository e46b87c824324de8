#!/bin/bash
# r02c: fused kernel with dynamic phase-2 hand-out + 96 KB phase-1 ring; f32x3 gates; exchange library on one GPU
OUT=gpurun_out
mkdir -p $OUT
echo "== distill tests"; timeout 900 python -m pytest tests/test_distill_gpu.py tests/test_operator_boundary_gpu.py -q -x 2>&1 | tail -30 | tee $OUT/pytest_distill_r02c.log
echo "== f32x3 tests"; timeout 900 python -m pytest tests/test_conv_f32x3_gpu.py -q 2>&1 | tail -8 | tee $OUT/pytest_f32x3_r02c.log
echo "== bench (short: headline + e2e only)"
timeout 600 python bench.py --steps 300 --warmup 20 --head-steps -1 --full-steps -1 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_r02c_short.json | cut -c1-1800
echo "== stamps"
SAD_FUSED_DEBUG=8 timeout 300 python scripts/fused_stamps.py > $OUT/fused_stamps_r02c.txt 2>&1; tail -12 $OUT/fused_stamps_r02c.txt
echo "== ncu full fused kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:distill_fused_kernel -s 5 -c 2 -o $OUT/prof_fused_r02c -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --head-steps -1 --full-steps -1 --e2e-steps 1 > $OUT/ncu_full_r02c.log 2>&1
ls -la $OUT | tail -5
