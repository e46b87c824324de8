#!/bin/bash
# Build the variant library first (it is not kept in the tree):
#   K=semi-supervised-adaptive-distillation_b200/csrc/kernels; mkdir -p variants; nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
#     -Xcompiler -fPIC,-fvisibility=hidden,-fno-gnu-unique -w -DSAD_CONV_ROWS=4 -Iinclude -I$K -shared -o variants/libsad_b200_rows4.so $K/*.cu
# r02s: A/B of the convolution's pixel tile: 8 rows x 32 (default build) against 4 rows x 32 (variants/libsad_b200_rows4.so, -DSAD_CONV_ROWS=4):
# head forward / backward / whole head step at bs = 2 and bs = 16, then the convolution + head tests on the 4-row build.
OUT=gpurun_out
PKG=semi-supervised-adaptive-distillation_b200
mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json,sys
tag=sys.argv[1]
for bs in (2,16):
    try:
        d=json.load(open('gpurun_out/head_bench_bs%d.json'%bs))
        print(tag,'bs',bs,'fwd %.3f ms %.0f TF/s'%(d['head_forward_eager']['ms'],d['head_forward_eager']['tflops']),'bwd %.3f ms %.0f TF/s'%(d['head_backward_eager']['ms'],d['head_backward_eager']['tflops']),'step(graph) %.3f ms %.0f TF/s'%(d['step_graph']['ms'],d['step_graph']['tflops']))
        import shutil; shutil.copy('gpurun_out/head_bench_bs%d.json'%bs,'gpurun_out/head_bench_r02s_%s_bs%d.json'%(tag,bs))
    except Exception as e: print(tag,bs,'failed',e)
PY
}
for bs in 2 16; do timeout 200 python scripts/head_bench.py --bs $bs --iters 30 > $OUT/head_bench_r02s_rows8_bs$bs.log 2>&1; done
summ rows8
cp $PKG/libsad_b200.so /tmp/libsad_b200_rows8.so
cp variants/libsad_b200_rows4.so $PKG/libsad_b200.so; touch $PKG/libcaffe2_detectron_ops_gpu.so $PKG/libsad_exchange.so
for bs in 2 16; do timeout 200 python scripts/head_bench.py --bs $bs --iters 30 > $OUT/head_bench_r02s_rows4_bs$bs.log 2>&1; done
summ rows4
echo "== tests on the 4-row build"
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_conv_f16_gpu.py tests/test_conv_f32x3_gpu.py tests/test_head_gpu.py tests/test_operator_boundary_gpu.py -x -q 2>&1 | tail -6
