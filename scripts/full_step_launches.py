"""One full distillation step (configs[3] geometry: R-50-FPN student <- R-101-FPN teacher, bs = 2, 640x1024) inside a
cudaProfilerStart / Stop range, for an ncu launch list of exactly that step:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/full_step_launches.csv \
        python scripts/full_step_launches.py
    python scripts/launch_summary.py gpurun_out/full_step_launches.csv profiles/rNN_full_step_launches.txt --native-share

The step runs eagerly here (the same launches the captured graph replays): forward_backward + exchange + optimiser launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sad_b200.full_step import FullDistillStep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
step = FullDistillStep(n_images=n)
for _ in range(3):
    step.forward_backward()
    step.allreduce()
    step.sgd()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step.forward_backward()
step.allreduce()
step.sgd()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("losses", step.losses())
step.close()
