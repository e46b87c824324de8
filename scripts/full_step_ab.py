"""A/B of FullDistillStep options on one box: ms per captured step.  Usage: python scripts/full_step_ab.py [n_images]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sad_b200.full_step import FullDistillStep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
variants = [("overlap_teacher=False", dict(overlap_teacher=False)), ("overlap_teacher=True", dict(overlap_teacher=True)),
            ("fused_body=False", dict(fused_body=False))]
for name, kw in variants:
    st = FullDistillStep(n_images=n, **kw)
    ok = st.capture()
    for _ in range(3):
        st.run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        a.record()
        for _ in range(10):
            st.run()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 10)
    print("bs=%d %-24s graph=%s  %.3f ms/step  %.1f imgs/s" % (n, name, ok, best, n / best * 1e3), flush=True)
    st.head.close(); st.teacher_head.close()
    del st
    torch.cuda.empty_cache()
