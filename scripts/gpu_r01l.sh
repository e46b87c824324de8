#!/bin/bash
# r01l: anchor-run chunking of the host pipeline (tests + sweep) and ncu evidence for the fp16 convolution / body operators.
TAG=${1:-r01l}
OUT=gpurun_out
mkdir -p $OUT
echo "== host step tests"; timeout 200 python -m pytest tests/test_distill_gpu.py -m gpu -q -k "host_step" 2>&1 | tail -15 | tee $OUT/pytest_host_${TAG}.log
echo "== e2e sweep"; timeout 200 python scripts/e2e_sweep.py 2>&1 | tail -2 | tee $OUT/e2e_sweep_${TAG}.json
echo "== ncu launch list of the fp16 / body-operator microbenchmark"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_f16_${TAG}.csv \
    python scripts/f16_bench.py --iters 3 > $OUT/ncu_list_f16_${TAG}.log 2>&1
echo "== ncu full: fp16 convolution (pair kernel), AffineChannel, UpsampleNearest"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tf32_kernel|affine_channel_vec4|upsample2_vec4_kernel' -s 20 -c 6 \
    -o $OUT/prof_f16_${TAG} -f python scripts/f16_bench.py --iters 3 > $OUT/ncu_full_f16_${TAG}.log 2>&1
tail -3 $OUT/ncu_full_f16_${TAG}.log
ls -la $OUT | tail -12
