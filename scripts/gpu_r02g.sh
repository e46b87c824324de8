#!/bin/bash
# r02g: fused kernel v5b (fire-and-forget loss sums), whole GPU suite, full default bench on one GPU
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_r02g.log
echo "== stamps"
SAD_FUSED_DEBUG=8 timeout 300 python scripts/fused_stamps.py > $OUT/fused_stamps_r02g.txt 2>&1; tail -6 $OUT/fused_stamps_r02g.txt
echo "== bench (default flags)"
timeout 1500 python bench.py > $OUT/bench_r02g.json 2> $OUT/bench_r02g.err; tail -c 600 $OUT/bench_r02g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02g.json').read().strip().split('\n')[-1])
r=d['roofline']; print('value',d['value'],'kernel_ms',r['kernel_ms'],'frac',r['frac'],'e2e',d['e2e']['value'], 'copy_only', d['e2e']['copy_only']['ms_per_step'])
print(json.dumps(d['config']['step_imgs_s'], indent=1))
PY
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400 | tee $OUT/bench_ref_r02g.json
