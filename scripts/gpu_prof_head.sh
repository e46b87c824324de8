#!/bin/bash
# ncu evidence for the head kernels + the launch list of the bench command.  Usage: bash scripts/gpu_prof_head.sh <tag>
TAG=${1:-r01f}
OUT=gpurun_out; mkdir -p $OUT
echo "== launch list of bench.py"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_bench_${TAG}.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --head-steps 2 --full-steps -1 > $OUT/ncu_list_${TAG}.log 2>&1
python scripts/launch_summary.py $OUT/launches_bench_${TAG}.csv $OUT/launches_bench_${TAG}.txt | head -20
for spec in "conv:conv3x3_tf32_kernel:12:3" "wgrad:conv3x3_wgrad_tf32_kernel:3:2" "distill:distill_fused_kernel:1:2" "focal:focal_kernel:0:1"; do
  IFS=: read name rx skip cnt <<< "$spec"
  echo "== ncu full $name"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $OUT/prof_${name}_${TAG} -f \
      python scripts/head_bench.py --bs 2 --iters 2 --quick > $OUT/ncu_full_${name}_${TAG}.log 2>&1
  tail -2 $OUT/ncu_full_${name}_${TAG}.log
done
ls -la $OUT | tail -12
