#!/bin/bash
# ncu full capture of one kernel regex while running bench.py briefly.  Usage: bash scripts/gpu_prof.sh <tag> <kernel-regex> [skip] [count]
TAG=$1; RX=$2; SKIP=${3:-3}; CNT=${4:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $CNT -o $OUT/prof_${TAG} -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_${TAG}.log 2>&1
tail -3 $OUT/ncu_${TAG}.log
