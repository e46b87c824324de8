#!/bin/bash
# r02b: rewritten fused kernel (packed math, prefetch across the grid barrier, tail levels) + f32x3 two-chain + new tests, then bench + ncu
OUT=gpurun_out
mkdir -p $OUT
echo "== distill tests"; timeout 900 python -m pytest tests/test_distill_gpu.py -q -x 2>&1 | tail -30 | tee $OUT/pytest_distill_r02b.log
echo "== f32x3 tests"; timeout 900 python -m pytest tests/test_conv_f32x3_gpu.py -q 2>&1 | tail -15 | tee $OUT/pytest_f32x3_r02b.log
cp $OUT/f32x3_errors.txt $OUT/f32x3_errors_r02b.txt
echo "== pytest -m gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_r02b.log
echo "== bench (short: headline + e2e only)"
timeout 600 python bench.py --steps 300 --warmup 20 --head-steps -1 --full-steps -1 --no-cpu-baseline 2>&1 | tail -2 | tee $OUT/bench_r02b_short.json
echo "== stamps"
SAD_FUSED_DEBUG=8 timeout 300 python scripts/fused_stamps.py > $OUT/fused_stamps_r02b.txt 2>&1; tail -12 $OUT/fused_stamps_r02b.txt
echo "== ncu full fused kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:distill_fused_kernel -s 5 -c 2 -o $OUT/prof_fused_r02b -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --head-steps -1 --full-steps -1 --e2e-steps 1 > $OUT/ncu_full_r02b.log 2>&1
ls -la $OUT | tail -8
