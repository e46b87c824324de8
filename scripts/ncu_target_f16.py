#!/usr/bin/env python
"""ncu target: a handful of launches of the fp16 convolution (CTA-pair kernel, 256 -> 256 with ReLU + channels-last output and
256 -> 720 with NCHW output), the fp16 weight gradient and the body operators at BASELINE.json configs[1] geometry.
    ncu --set full -k regex:'conv3x3|affine_channel|upsample2' ... python scripts/ncu_target_f16.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.build()
from sad_b200 import ops  # noqa: E402

SHAPES = [(80, 128), (40, 64), (20, 32), (10, 16), (5, 8)]
g = torch.Generator(device="cuda").manual_seed(1)
xs = [torch.randn(2, 256, h, w, device="cuda", generator=g).clamp_(min=0) for h, w in SHAPES]
x16 = ops.to_nhwc_f16(xs)
for cout in (256, 720):
    w = torch.randn(cout, 256, 3, 3, device="cuda", generator=g) * 0.02
    b = torch.randn(cout, device="cuda", generator=g)
    p16 = ops.conv3x3_pack_f16(w)
    for _ in range(3):
        ops.conv3x3_forward_f16(x16, p16, cout, b, relu=1 if cout == 256 else 0, want_nchw=cout != 256, want_nhwc=cout == 256)
if os.environ.get("SAD_NCU_WGRAD", "1") == "1":
    dys = [torch.randn(2, h, w, 256, device="cuda", generator=g).half() for h, w in SHAPES]
    for _ in range(3):
        ops.conv3x3_wgrad_f16(x16, dys)
x = torch.randn(2, 256, 160, 256, device="cuda", generator=g)
s, b = torch.randn(256, device="cuda"), torch.randn(256, device="cuda")
y = torch.empty_like(x)
u = torch.randn(2, 256, 40, 64, device="cuda", generator=g)
du = torch.randn(2, 256, 80, 128, device="cuda", generator=g)
for _ in range(3):
    ops.affine_channel(x, s, b, out=y)
    ops.upsample_nearest(u, 2)
    ops.upsample_nearest_grad(u, du, 2)
torch.cuda.synchronize()
print("ncu target done")
