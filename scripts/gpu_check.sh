#!/bin/bash
# One GPU-box visit: parity tests, reference-GPU golden vectors, smoke, bench, ncu launch list + full capture.
# Usage (from the repo root on the box):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_${TAG}.txt 2>&1
nproc >> $OUT/gpu_${TAG}.txt
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_${TAG}.log
echo "== ref gpu golden" ; timeout 300 python tests/golden/make_ref_gpu_golden.py 2>&1 | tail -5
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/smoke_${TAG}.log
echo "== bench" ; timeout 900 python bench.py 2>&1 | tail -3 | tee $OUT/bench_${TAG}.json
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee $OUT/bench_ref_${TAG}.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --head-steps 2 --full-steps -1 > $OUT/ncu_list_${TAG}.log 2>&1
echo "== ncu full (the dominant kernel: PowSum + loss + gradient in one cooperative launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:distill_fused_kernel -s 5 -c 2 -o $OUT/prof_fused_${TAG} -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --head-steps -1 --full-steps -1 --e2e-steps 1 > $OUT/ncu_full_${TAG}.log 2>&1
echo "== head kernels: see scripts/gpu_prof_head.sh"
ls -la $OUT
