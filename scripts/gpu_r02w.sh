#!/bin/bash
# r02w: final validation of the round-2 tree: whole GPU suite, smoke, default bench line, reference arm (short).
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_r02w.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== default bench"
timeout 900 python bench.py > $OUT/bench_r02w.log 2>&1; echo "exit $?"; tail -1 $OUT/bench_r02w.log > $OUT/bench_r02w.json
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_r02w_ref.log 2>&1; echo "exit $?"; tail -1 $OUT/bench_r02w_ref.log > $OUT/bench_r02w_reference_arm.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02w.json').read())
print('value',round(d['value'],1),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value'],2),'launches',d['gpu_launches'],'clocks',d['clocks'])
print({k:round(v['imgs_s'],1) for k,v in d['config']['step_imgs_s'].items()})
r=json.loads(open('gpurun_out/bench_r02w_reference_arm.json').read())
print('reference arm',r.get('value'),r.get('unit'),r.get('cpu_baseline',{}).get('cores'))
PY
