#!/bin/bash
# r01n: fp16 data gradient + fp16 head backward (new), microbenchmark incl. the student head in both precisions, whole GPU suite.
TAG=${1:-r01n}
OUT=gpurun_out
mkdir -p $OUT
echo "== fp16 tests"; timeout 240 python -m pytest tests/test_conv_f16_gpu.py -m gpu -q 2>&1 | tail -40 | tee $OUT/pytest_f16_${TAG}.log
echo "== f16 bench"; timeout 200 python scripts/f16_bench.py 2>&1 | tail -3 | tee $OUT/f16_bench_${TAG}.json
echo "== whole gpu suite"; timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_${TAG}.log
ls -la $OUT | tail -5
