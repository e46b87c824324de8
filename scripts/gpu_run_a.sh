set -x
python -m pytest tests -m gpu -x -q > gpurun_out/j_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/j_tests.log
python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; tail -c 600 gpurun_out/j_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/j_bench_ref.json 2>> gpurun_out/j_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/j_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
print('head_step', d['head_step']['value'], d['head_step']['ms_per_step'], d['head_step']['roofline']['frac'])
for k in ('full_step','full_step_bs16','full_step_config5'):
    print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('cuda_graph'))
print(d['cpu_baseline']); print(d.get('operator_net'))
P
