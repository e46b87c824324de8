set -x
python -m pytest tests/test_full_step_gpu.py -x -q > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/a_tests.log
python bench.py --steps 50 --warmup 5 --head-steps -1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/a_bench.json').read().strip().splitlines()[-1])
for k in ('full_step','full_step_bs16','full_step_config5'):
    print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('cuda_graph'), d[k].get('cuda_graph_error'), d[k]['params'])
P
tail -5 gpurun_out/a_bench.err
