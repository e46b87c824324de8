"""Bring-up harness: one small convolution, prints max error vs torch fp32.  Env SAD_CONV_DEBUG bisects the kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
from sad_b200 import ops
torch.backends.cudnn.allow_tf32 = False
shape = tuple(int(v) for v in (sys.argv[1:6] if len(sys.argv) >= 6 else (1, 32, 128, 8, 32)))
N, Cin, Cout, H, W = shape
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(N, Cin, H, W, device="cuda", generator=g).clamp_(min=0)
w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / np.sqrt(9 * Cin)
b = torch.randn(Cout, device="cuda", generator=g)
ref = torch.nn.functional.conv2d(x, w, b, padding=1)
torch.cuda.synchronize()
y = ops.conv3x3_forward([x], w, b)[0][0]
torch.cuda.synchronize()
d = (y - ref).abs()
print("shape", shape, "debug", os.environ.get("SAD_CONV_DEBUG"), "rows/box", os.environ.get("SAD_CONV_ROWS_PER_BOX"),
      "max|d| %.3g max|ref| %.3g rel %.3g" % (d.max().item(), ref.abs().max().item(), (d.max() / ref.abs().max()).item()))
if d.max() > 3e-3 * ref.abs().max():
    bad = (d > 3e-3 * ref.abs().max())
    idx = bad.nonzero()
    print("bad elements:", int(bad.sum()), "of", bad.numel(), "first", idx[:5].tolist())
    print("per-channel bad counts (first 16):", bad.sum(dim=(0, 2, 3))[:16].tolist())
    print("per-row bad counts:", bad.sum(dim=(0, 1, 3)).tolist())
    print("per-col bad counts:", bad.sum(dim=(0, 1, 2)).tolist())
