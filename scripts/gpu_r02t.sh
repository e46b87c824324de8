#!/bin/bash
# r02t: A/B of the CTA-pair convolution with one tile per weight stage (SAD_CONV_DUAL=0) against two tiles per weight stage (default):
# head forward / backward / whole head step at bs = 2 and bs = 16, then the convolution / head / operator / full-step tests.
OUT=gpurun_out
mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json,sys,shutil
tag=sys.argv[1]
for bs in (2,16):
    try:
        d=json.load(open('gpurun_out/head_bench_bs%d.json'%bs))
        print(tag,'bs',bs,'fwd %.3f ms %.0f TF/s'%(d['head_forward_eager']['ms'],d['head_forward_eager']['tflops']),'bwd %.3f ms %.0f TF/s'%(d['head_backward_eager']['ms'],d['head_backward_eager']['tflops']),'step(graph) %.3f ms %.0f TF/s'%(d['step_graph']['ms'],d['step_graph']['tflops']))
        shutil.copy('gpurun_out/head_bench_bs%d.json'%bs,'gpurun_out/head_bench_r02t_%s_bs%d.json'%(tag,bs))
    except Exception as e: print(tag,bs,'failed',e)
PY
}
for bs in 2 16; do SAD_CONV_DUAL=0 timeout 200 python scripts/head_bench.py --bs $bs --iters 30 > $OUT/head_bench_r02t_single_bs$bs.log 2>&1 || tail -5 $OUT/head_bench_r02t_single_bs$bs.log; done
summ single
rm -f $OUT/head_bench_bs2.json $OUT/head_bench_bs16.json
for bs in 2 16; do timeout 200 python scripts/head_bench.py --bs $bs --iters 30 > $OUT/head_bench_r02t_dual_bs$bs.log 2>&1 || tail -5 $OUT/head_bench_r02t_dual_bs$bs.log; done
summ dual
echo "== tests on the two-tile form"
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_conv_f16_gpu.py tests/test_conv_f32x3_gpu.py tests/test_head_gpu.py tests/test_operator_boundary_gpu.py tests/test_full_step_gpu.py -x -q 2>&1 | tail -8
