#!/bin/bash
# r02v: A/B of the CTA-pair convolution: one pixel box per tap (SAD_CONV_HALO=0) against halo tiles (pixel rows + the rows above / below loaded once
# per channel block and dx, the three dy taps as row offsets): head forward / backward / whole head step at bs = 2 and 16, then the tests.
OUT=gpurun_out
mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json,sys,shutil
tag=sys.argv[1]
for bs in (2,16):
    try:
        d=json.load(open('gpurun_out/head_bench_bs%d.json'%bs))
        print(tag,'bs',bs,'fwd %.3f ms %.0f TF/s'%(d['head_forward_eager']['ms'],d['head_forward_eager']['tflops']),'bwd %.3f ms %.0f TF/s'%(d['head_backward_eager']['ms'],d['head_backward_eager']['tflops']),'step(graph) %.3f ms %.0f TF/s'%(d['step_graph']['ms'],d['step_graph']['tflops']))
        shutil.copy('gpurun_out/head_bench_bs%d.json'%bs,'gpurun_out/head_bench_r02v_%s_bs%d.json'%(tag,bs))
    except Exception as e: print(tag,bs,'failed',e)
PY
}
for bs in 2 16; do SAD_CONV_HALO=0 timeout 200 python scripts/head_bench.py --bs $bs --iters 30 > $OUT/head_bench_r02v_pertap_bs$bs.log 2>&1 || tail -5 $OUT/head_bench_r02v_pertap_bs$bs.log; done
summ pertap
rm -f $OUT/head_bench_bs2.json $OUT/head_bench_bs16.json
for bs in 2 16; do timeout 200 python scripts/head_bench.py --bs $bs --iters 30 > $OUT/head_bench_r02v_halo_bs$bs.log 2>&1 || tail -5 $OUT/head_bench_r02v_halo_bs$bs.log; done
summ halo
echo "== tests on the halo form"
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_conv_f16_gpu.py tests/test_conv_f32x3_gpu.py tests/test_head_gpu.py tests/test_operator_boundary_gpu.py tests/test_full_step_gpu.py -x -q 2>&1 | tail -8
