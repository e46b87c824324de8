#!/usr/bin/env python
"""Times the fp16-operand forward convolution against the tf32 one on the head's geometry (all 5 levels in one launch), the
teacher head forward in both precisions, and the new body operators (AffineChannel, UpsampleNearest, Scale) with CUDA events.
    python scripts/f16_bench.py [--bs 2] [--iters 50]  ->  one JSON object
Not a bench line (bench.py is); produces profiles/ summaries."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.build()
from sad_b200 import head, ops  # noqa: E402


def timeit(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bs", type=int, default=2)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--px", type=int, default=600)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    shapes = [(80, 128), (40, 64), (20, 32), (10, 16), (5, 8)] if a.px == 600 else [(64, 112), (32, 56), (16, 28), (8, 14), (4, 7)]
    g = torch.Generator(device="cuda").manual_seed(1)
    pixels = a.bs * sum(h * w for h, w in shapes)
    res = {"bs": a.bs, "px": a.px, "pixels": pixels}

    def rnd(*s):
        return torch.randn(*s, device="cuda", generator=g)

    xs = [rnd(a.bs, 256, h, w).clamp_(min=0) for h, w in shapes]
    x32, x16 = ops.to_nhwc(xs), ops.to_nhwc_f16(xs)
    res["to_nhwc_tf32_ms"] = timeit(lambda: ops.to_nhwc(xs), a.iters)
    res["to_nhwc_f16_ms"] = timeit(lambda: ops.to_nhwc_f16(xs), a.iters)
    for cout in (256, 720, 36):
        w, b = rnd(cout, 256, 3, 3) * 0.02, rnd(cout)
        p32, p16 = ops.conv3x3_pack(w, 0), ops.conv3x3_pack_f16(w)
        flops = 2.0 * pixels * cout * 2304
        nhwc = cout == 256
        ms32 = timeit(lambda: ops.conv3x3_forward(None, w, b, relu=nhwc, packed=p32, xs_nhwc=x32, want_nchw=not nhwc, want_nhwc=nhwc), a.iters)
        ms16 = timeit(lambda: ops.conv3x3_forward_f16(x16, p16, cout, b, relu=1 if nhwc else 0, want_nchw=not nhwc, want_nhwc=nhwc), a.iters)
        y32 = ops.conv3x3_forward(None, w, b, packed=p32, xs_nhwc=x32)[0]
        y16 = ops.conv3x3_forward_f16(x16, p16, cout, b)[0]
        dmax = max((p - q).abs().max().item() / p.abs().max().item() for p, q in zip(y32, y16))
        res["fwd_256_%d" % cout] = {"tf32_ms": ms32, "tf32_tflops": flops / ms32 / 1e9, "f16_ms": ms16, "f16_tflops": flops / ms16 / 1e9,
                                    "speedup": ms32 / ms16, "max_rel_diff_f16_vs_tf32": dmax}
    # the teacher head (forward only, Sigmoid fused) in both precisions
    fpn = [x * 0.5 for x in xs]
    for name, f16 in (("tf32", False), ("f16", True)):
        h = head.RetinaNetHead(a.bs, shapes, cls_output_sigmoid=True, seed=5, compute_f16=f16)
        out = h.alloc_outputs()
        ms = timeit(lambda: h.forward(fpn, training=False, out=out), a.iters)
        flops = 2.0 * pixels * 2304 * (8 * 256 + 720 + 36)
        res["teacher_head_forward_" + name] = {"ms": ms, "tflops": flops / ms / 1e9}
        h.close()
    # the student head: forward (training) + backward (data + weight gradients of all 10 convolutions) in both precisions
    for name, f16 in (("tf32", False), ("f16", True)):
        h = head.RetinaNetHead(a.bs, shapes, seed=5, compute_f16=f16)
        cls, box = h.alloc_outputs()
        d_cls = [torch.randn(c.shape, device="cuda", generator=g) * 1e-5 for c in cls]
        d_box = [torch.randn(c.shape, device="cuda", generator=g) * 1e-5 for c in box]
        d_fpn = [torch.empty_like(x) for x in fpn]

        def step():
            h.forward(fpn, training=True, out=(cls, box))
            h.backward(d_cls, d_box, d_fpn=d_fpn)

        ms = timeit(step, a.iters)
        flops = 3 * 2.0 * pixels * 2304 * (8 * 256 + 720 + 36)
        res["student_head_fwd_bwd_" + name] = {"ms": ms, "tflops": flops / ms / 1e9}
        h.close()
    # body operators (HBM streams): res2-sized AffineChannel, the FPN top-down upsamples, momentum-sized Scale
    x = rnd(a.bs, 256, 160, 256)
    s, b = rnd(256), rnd(256)
    y = torch.empty_like(x)
    ms = timeit(lambda: ops.affine_channel(x, s, b, out=y), a.iters)
    res["affine_channel_%dx256x160x256" % a.bs] = {"ms": ms, "gbs": 8.0 * x.numel() / ms / 1e6}
    u = rnd(a.bs, 256, 40, 64)
    ms = timeit(lambda: ops.upsample_nearest(u, 2), a.iters)
    res["upsample_nearest_%dx256x40x64" % a.bs] = {"ms": ms, "gbs": 20.0 * u.numel() / ms / 1e6}
    du = rnd(a.bs, 256, 80, 128)
    ms = timeit(lambda: ops.upsample_nearest_grad(u, du, 2), a.iters)
    res["upsample_nearest_grad_%dx256x40x64" % a.bs] = {"ms": ms, "gbs": 20.0 * u.numel() / ms / 1e6}
    m = rnd(37_700_000)
    ms = timeit(lambda: ops.scale_(m, 1.0), a.iters)
    res["scale_37.7M"] = {"ms": ms, "gbs": 8.0 * m.numel() / ms / 1e6}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
