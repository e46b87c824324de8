#!/usr/bin/env python
"""Times the head forward / backward / whole distillation step (eager launches and CUDA-graph replay) on one GPU.
    python scripts/head_bench.py [--bs 2] [--iters 30] [--quick]
Not a bench line (bench.py is)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.build()
from sad_b200 import native  # noqa: E402
from sad_b200.step import DistillHeadStep  # noqa: E402


def timeit(fn, iters, warm=3):
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
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--quick", action="store_true", help="only the eager step (for ncu launch lists)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    st = DistillHeadStep(n_images=a.bs)
    fwd_f, bwd_f = st.flops()
    res = {"bs": a.bs, "fwd_gflop": fwd_f / 1e9, "bwd_gflop": bwd_f / 1e9, "head_device_MB": st.head.device_bytes() / 1e6}
    if a.quick:
        for _ in range(a.iters):
            st.forward_backward()
        torch.cuda.synchronize()
        return
    n0 = native.lib().sad_launch_count()
    st.forward_backward()
    res["launches_per_step"] = int(native.lib().sad_launch_count() - n0)
    ms = timeit(lambda: st.head.forward(st.fpn, training=True, out=(st.cls, st.box)), a.iters)
    res["head_forward_eager"] = {"ms": ms, "tflops": fwd_f / ms / 1e9}
    ms = timeit(lambda: st.head.backward(st.plan.grads, st.d_box, want_d_fpn=True, d_fpn=st.d_fpn), a.iters)
    res["head_backward_eager"] = {"ms": ms, "tflops": bwd_f / ms / 1e9}
    ms = timeit(st.plan.run, a.iters)
    res["powsum_distill_eager_ms"] = ms
    ms = timeit(st.forward_backward, a.iters)
    res["step_eager"] = {"ms": ms, "tflops": (fwd_f + bwd_f) / ms / 1e9, "imgs_per_s": a.bs / ms * 1e3}
    st.capture()
    ms = timeit(st.run, a.iters)
    res["step_graph"] = {"ms": ms, "tflops": (fwd_f + bwd_f) / ms / 1e9, "imgs_per_s": a.bs / ms * 1e3}
    res["losses"] = st.losses()
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "head_bench_bs%d.json" % a.bs), "w"), indent=1)


if __name__ == "__main__":
    main()
