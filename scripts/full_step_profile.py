"""Where the full distillation step's GPU time goes, by kernel (eager step under torch.profiler; kernel durations are
device-side, so eager launch overhead does not distort them).  Usage: python scripts/full_step_profile.py [n_images]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from sad_b200.full_step import FullDistillStep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
step = FullDistillStep(n_images=n)
for _ in range(3):
    step.forward_backward()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step.forward_backward()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
total = sum(r[2] for r in rows)
print("total device time %.3f ms over %d kernels (bs=%d)" % (total / 1e3, sum(r[1] for r in rows), n))
for k, c, t in rows[:45]:
    print("%8.1f us %5.1f%% %5d  %s" % (t, 100.0 * t / total, c, k[:110]))
