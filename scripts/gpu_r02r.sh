#!/bin/bash
# r02r: validation of the tree + evidence: whole GPU suite, smoke, ncu launch list of ONE full step, default bench line.
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_r02r.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== launch list: one full step"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/full_step_launches_r02r.csv \
    python scripts/full_step_launches.py > $OUT/ncu_full_step_r02r.log 2>&1
tail -2 $OUT/ncu_full_step_r02r.log
python scripts/launch_summary.py $OUT/full_step_launches_r02r.csv $OUT/full_step_launches_r02r.txt --native-share | tail -6
echo "== default bench"
timeout 900 python bench.py > $OUT/bench_r02r.log 2>&1; echo "exit $?"; tail -1 $OUT/bench_r02r.log > $OUT/bench_r02r.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02r.json').read())
print('value',round(d['value'],1),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value'],2),'clocks',d['clocks'])
print(json.dumps(d['config'].get('step_imgs_s'))[:900])
PY
