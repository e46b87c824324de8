#!/bin/bash
# r02a: new 3xTF32 tests first (fast feedback), then the whole GPU suite, then a fresh ncu capture of the fused loss kernel.
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_r02a.txt 2>&1
nproc >> $OUT/gpu_r02a.txt
echo "== f32x3 tests"; timeout 900 python -m pytest tests/test_conv_f32x3_gpu.py -q 2>&1 | tail -40 | tee $OUT/pytest_f32x3_r02a.log
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee $OUT/pytest_gpu_r02a.log
echo "== ncu full fused kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:distill_fused_kernel -s 5 -c 2 -o $OUT/prof_fused_r02a -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --head-steps -1 --full-steps -1 --e2e-steps 1 > $OUT/ncu_full_r02a.log 2>&1
ls -la $OUT | tail -12
