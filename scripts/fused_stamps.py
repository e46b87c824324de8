"""Debug: phase timestamps of distill_fused_kernel (SAD_FUSED_DEBUG=8 [+4 static])."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from sad_b200 import ops, synthetic
host = synthetic.make_pyramid(1234, 2, 600)
HEAD = dict(gamma=2.0, alpha=0.5, beta=0.0, scale=1.0, num_classes=80, ignored_label=-1)
dev = [tuple(torch.from_numpy(a).cuda() for a in l) for l in host]
plan = ops.DistillPlan(dev, power=1.8, **HEAD)
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
for it in range(4):
    flush.fill_(it)
    plan.run(); torch.cuda.synchronize()
ws = plan.ws_fused.buf
off = ws.numel() - 1024 * 5 * 8          # the debug stamps are the last block of the workspace (distill_fused.cu)
st = ws[off:off + 1024 * 5 * 8].view(torch.int64).cpu().numpy().reshape(1024, 5)[:296]
t0 = st[:, 0].min()
r = (st - t0) / 1e3
print("start  min/med/max us", r[:, 0].min(), np.median(r[:, 0]), r[:, 0].max())
print("p1 end min/med/max us", r[:, 1].min(), np.median(r[:, 1]), r[:, 1].max())
print("p2 beg min/med/max us", r[:, 2].min(), np.median(r[:, 2]), r[:, 2].max())
print("p2 end min/med/max us", r[:, 3].min(), np.median(r[:, 3]), r[:, 3].max())
print("final  us", r[:, 4].max())
