#!/bin/bash
# r02j: whole GPU suite on the current tree, then the evidence captures for profiles/: fused kernel (--set full), conv instantiations
# (--set full, one launch each), launch lists of the headline bench, the head step and the full step.
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $OUT/pytest_gpu_r02j.log
echo "== ncu full: fused kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:distill_fused_kernel -s 5 -c 2 -o $OUT/prof_fused_r02j -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --head-steps -1 --full-steps -1 --e2e-steps 1 > $OUT/ncu_full_fused_r02j.log 2>&1
echo "== launch list: headline + head step + full step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_bench_r02j.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --head-steps 2 --full-steps 2 --no-heads-f16 > $OUT/ncu_list_r02j.log 2>&1
python scripts/launch_summary.py $OUT/launches_bench_r02j.csv $OUT/launches_bench_r02j.txt | head -45
for spec in "conv:conv3x3_tf32_kernel:12:3" "wgrad:conv3x3_wgrad_tf32_kernel:3:2"; do
  IFS=: read name rx skip cnt <<< "$spec"
  echo "== ncu full $name"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $OUT/prof_${name}_r02j -f \
      python scripts/head_bench.py --bs 2 --iters 2 --quick > $OUT/ncu_full_${name}_r02j.log 2>&1
  tail -2 $OUT/ncu_full_${name}_r02j.log
done
ls -la $OUT | tail -8
