#!/bin/bash
# r01m: fp16 weight gradient (new), host pipeline with hoisted label copies (sweep), whole GPU suite, ncu full of the fp16 kernels.
TAG=${1:-r01m}
OUT=gpurun_out
mkdir -p $OUT
echo "== fp16 wgrad tests"; timeout 200 python -m pytest tests/test_conv_f16_gpu.py -m gpu -q -k "wgrad" 2>&1 | tail -30 | tee $OUT/pytest_wgrad_f16_${TAG}.log
echo "== whole gpu suite"; timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_${TAG}.log
echo "== e2e sweep"; timeout 200 python scripts/e2e_sweep.py 2>&1 | tail -2 | tee $OUT/e2e_sweep_${TAG}.json
echo "== ncu full: fp16 convolution / wgrad, AffineChannel, UpsampleNearest"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tf32_kernel|conv3x3_wgrad_tf32_kernel|affine_channel_vec4|upsample2' -c 24 \
    -o $OUT/prof_f16_${TAG} -f python scripts/ncu_target_f16.py > $OUT/ncu_full_f16_${TAG}.log 2>&1
tail -3 $OUT/ncu_full_f16_${TAG}.log
ls -la $OUT | tail -8
