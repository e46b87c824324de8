#!/bin/bash
# r01s: last check of the round: the whole GPU suite and smoke on the final libraries.
TAG=${1:-r01s}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee $OUT/pytest_gpu_${TAG}.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/smoke_${TAG}.log
