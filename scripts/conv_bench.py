#!/usr/bin/env python
"""Times the head convolution kernels on BASELINE.json configs[1] geometry (bs=2, 600 px, 5 levels in
one launch) with CUDA events and prints TFLOP/s per kernel.  Run on the GPU box:
    python scripts/conv_bench.py [--bs 2] [--iters 50]
Not a bench line (bench.py is); used to steer kernel work and to produce profiles/ summaries."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.build()
from sad_b200 import ops  # noqa: E402

SHAPES = [(80, 128), (40, 64), (20, 32), (10, 16), (5, 8)]


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
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    pixels = a.bs * sum(h * w for h, w in SHAPES)
    res = {"bs": a.bs, "pixels": pixels}

    def rnd(*s):
        return torch.randn(*s, device="cuda", generator=g)

    xs = [rnd(a.bs, 256, h, w).clamp_(min=0) for h, w in SHAPES]
    res["to_nhwc_256_ms"] = timeit(lambda: ops.to_nhwc(xs), a.iters)
    xs_cl = ops.to_nhwc(xs)
    couts = [int(c) for c in a.only.split(',')] if a.only else [256, 720, 36]
    for cout in couts:
        w = rnd(cout, 256, 3, 3) * 0.02
        b = rnd(cout)
        packed = ops.conv3x3_pack(w, 0)
        flops = 2.0 * pixels * cout * 2304
        if cout == 256:
            ms = timeit(lambda: ops.conv3x3_forward(None, w, b, relu=True, packed=packed, xs_nhwc=xs_cl, want_nchw=False, want_nhwc=True), a.iters)
            res["fwd_256_256_relu_nhwc"] = {"ms": ms, "tflops": flops / ms / 1e9}
        ms = timeit(lambda: ops.conv3x3_forward(None, w, b, packed=packed, xs_nhwc=xs_cl, want_nchw=True, want_nhwc=False), a.iters)
        res["fwd_256_%d_nchw" % cout] = {"ms": ms, "tflops": flops / ms / 1e9}
        # data gradient: K = cout, M = 256
        dys = [rnd(a.bs, h, w, cout) for h, w in SHAPES]
        pk1 = ops.conv3x3_pack(w, 1)
        ms = timeit(lambda: ops.conv3x3_dgrad(None, w, packed=pk1, dys_nhwc=dys, want_nchw=False, want_nhwc=True), a.iters)
        res["dgrad_%d_256_nhwc" % cout] = {"ms": ms, "tflops": flops / ms / 1e9}
        if hasattr(ops, "conv3x3_wgrad"):
            ms = timeit(lambda: ops.conv3x3_wgrad(xs_cl, dys), a.iters)
            res["wgrad_256_%d" % cout] = {"ms": ms, "tflops": flops / ms / 1e9}
        res["pack_%d_ms" % cout] = timeit(lambda: ops.conv3x3_pack(w, 0), a.iters)
    if a.only:
        print(json.dumps(res, indent=1))
        return
    # torch / cuDNN fp32 (TF32 allowed and not) for context: library call, not the product
    import torch.nn.functional as F
    w = rnd(256, 256, 3, 3) * 0.02
    b = rnd(256)
    for allow in (False, True):
        torch.backends.cudnn.allow_tf32 = allow
        ms = timeit(lambda: [F.conv2d(x, w, b, padding=1) for x in xs], max(5, a.iters // 5))
        res["cudnn_fwd_256_256_%s" % ("tf32" if allow else "fp32")] = {"ms": ms, "tflops": 2.0 * pixels * 256 * 2304 / ms / 1e9}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "conv_bench_bs%d.json" % a.bs), "w"), indent=1)


if __name__ == "__main__":
    main()
