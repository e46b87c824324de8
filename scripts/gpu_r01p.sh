#!/bin/bash
# r01p: fp16 head against the staged fp64 reference, whole GPU suite, bench with both heads of configs[4] in fp16.
TAG=${1:-r01p}
OUT=gpurun_out
mkdir -p $OUT
echo "== fp16 tests"; timeout 240 python -m pytest tests/test_conv_f16_gpu.py -m gpu -q 2>&1 | tail -60 | tee $OUT/pytest_f16_${TAG}.log
echo "== whole gpu suite"; timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_${TAG}.log
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/smoke_${TAG}.log
echo "== bench"; timeout 500 python bench.py --heads-f16 2>&1 | tail -2 | tee $OUT/bench_${TAG}.json
