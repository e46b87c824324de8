#!/bin/bash
# r01q: ncu of the final fp16 weight gradient (64-pixel stages), then the two bench arms exactly as the driver runs them.
TAG=${1:-r01q}
OUT=gpurun_out
mkdir -p $OUT
echo "== ncu full: fp16 weight gradient"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_wgrad_tf32_kernel' -c 3 \
    -o $OUT/prof_wgrad_f16_${TAG} -f python scripts/ncu_target_f16.py > $OUT/ncu_wgrad_f16_${TAG}.log 2>&1
tail -2 $OUT/ncu_wgrad_f16_${TAG}.log
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_${TAG}.json | cut -c1-400
echo "== bench (default)"; timeout 500 python bench.py 2>&1 | tail -1 | tee $OUT/bench_${TAG}.json | cut -c1-300
