#!/bin/bash
# r01o: 64-pixel stages in the fp16 weight gradient, statistical gate of the fp16 head backward, microbenchmark, whole GPU suite.
TAG=${1:-r01o}
OUT=gpurun_out
mkdir -p $OUT
echo "== fp16 tests"; timeout 240 python -m pytest tests/test_conv_f16_gpu.py -m gpu -q 2>&1 | tail -60 | tee $OUT/pytest_f16_${TAG}.log
echo "== f16 bench"; timeout 200 python scripts/f16_bench.py 2>&1 | tail -3 | tee $OUT/f16_bench_${TAG}.json
echo "== whole gpu suite"; timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_${TAG}.log
