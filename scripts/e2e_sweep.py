#!/usr/bin/env python
"""Times sad_distill_step_host at BASELINE.json configs[1] (bs=2, 600 px, 5 levels; pinned host buffers, gradients returned) for
several pipeline granularities (sad_ctx_set_host_chunk_bytes).  One JSON object; steers the default chunk size.
    python scripts/e2e_sweep.py [--steps 15]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.build()
from sad_b200 import ops, synthetic  # noqa: E402

HEAD = dict(gamma=2.0, alpha=0.5, beta=0.0, scale=1.0, num_classes=80, ignored_label=-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=15)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    host = synthetic.make_pyramid(1234, 2, 600)
    anchors = sum(l[2].size for l in host)
    cpu = [tuple(torch.from_numpy(x).pin_memory() for x in l) for l in host]
    outs = [torch.empty_like(l[0]).pin_memory() for l in cpu]
    res = {"anchors": anchors, "steps": a.steps, "h2d_bytes": int(sum(l[0].nbytes + l[1].nbytes + l[2].nbytes for l in host)),
           "d2h_bytes": int(sum(l[0].nbytes for l in host)), "sweep": {}}
    ref = None
    for cap in (1 << 40, 16 << 20, 8 << 20, 4 << 20, 2 << 20, 1 << 20):
        step = ops.HostStep(0)
        step.set_chunk_bytes(cap)
        step.bind(cpu, outs, power=1.8, **HEAD)
        for _ in range(3):
            losses, norm = step.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            losses, norm = step.run()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / a.steps * 1e3
        step.close()
        if ref is None:
            ref = [o.clone() for o in outs]
        same = all(torch.equal(x, y) for x, y in zip(ref, outs))
        res["sweep"]["whole images" if cap == 1 << 40 else "%d MB" % (cap >> 20)] = {
            "ms_per_step": ms, "m_anchors_per_s": anchors / ms / 1e3, "gradients_bit_equal_to_whole_image_run": same,
            "losses": [float(x) for x in losses]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
