#!/bin/bash
# Quick GPU visit: parity tests + one bench line (no profiler).  Usage: bash scripts/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
if [ -n "$2" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "$2" 2>&1 | tail -30 | tee $OUT/pytest_gpu_${TAG}.log
else
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee $OUT/pytest_gpu_${TAG}.log
fi
echo "== bench"
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -3 | tee $OUT/bench_${TAG}.json
