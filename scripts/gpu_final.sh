#!/bin/bash
# The round's last GPU-box visit (tight budget): new parity tests first, then the whole GPU suite, microbenchmarks, bench.
TAG=${1:-r01k}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/gpu_${TAG}.txt 2>&1
echo "== new tests"; timeout 240 python -m pytest tests/test_body_ops_gpu.py tests/test_conv_f16_gpu.py -m gpu -q 2>&1 | tail -40 | tee $OUT/pytest_new_${TAG}.log
echo "== f16 bench"; timeout 150 python scripts/f16_bench.py 2>&1 | tail -3 | tee $OUT/f16_bench_${TAG}.json
echo "== full gpu suite"; timeout 420 python -m pytest tests -m gpu -q --deselect tests/test_body_ops_gpu.py --deselect tests/test_conv_f16_gpu.py 2>&1 | tail -15 | tee $OUT/pytest_gpu_${TAG}.log
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/smoke_${TAG}.log
echo "== bench"; timeout 400 python bench.py --teacher-f16 2>&1 | tail -2 | tee $OUT/bench_${TAG}.json
ls -la $OUT | tail -12
