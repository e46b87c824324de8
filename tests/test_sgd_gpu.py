"""GPU parity of the one-launch momentum-SGD update against the CPU oracle restating
detectron/lib/modeling/optimizer.py:115-130 (Scale 2x for biases, WeightedSum weight decay) + MomentumSGDKernel
(caffe2/caffe2/sgd/momentum_sgd_op_gpu.cu:23-54).  Tolerance: 2 ulp-level (rtol 1e-6): the same handful of fp32
operations, possibly contracted into FMAs on the device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nesterov", [False, True])
@pytest.mark.parametrize("sizes", [[(1000, 1.0, 1e-4), (40, 2.0, 0.0)], [(7, 1.0, 1e-4), (5, 2.0, 0.0), (1026, 1.0, 5e-4), (3, 2.0, 0.0)]])
def test_momentum_sgd_matches_oracle(oracle, sizes, nesterov):
    from sad_b200 import ops
    rng = np.random.default_rng(len(sizes) + int(nesterov))
    n = sum(c for c, _, _ in sizes)
    p, g, m = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
    pd, gd, md = (torch.from_numpy(a.copy()).cuda() for a in (p, g, m))
    lr = torch.tensor(0.0137, device="cuda")
    for step in range(2):     # two steps: the momentum written by the first feeds the second
        ops.momentum_sgd(pd, gd, md, sizes, lr, momentum=0.9, nesterov=nesterov)
        off = 0
        for cnt, mult, wd in sizes:
            sl = slice(off, off + cnt)
            p[sl], g[sl], m[sl] = oracle.momentum_sgd(p[sl], g[sl], m[sl], 0.0137, 0.9, nesterov, mult, wd)
            off += cnt
        torch.cuda.synchronize()
        for got, ref, what in ((pd, p, "param"), (gd, g, "grad"), (md, m, "momentum")):
            np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=2e-6, atol=2e-6, err_msg="%s step %d" % (what, step))  # atol: (1 + mu) * m_new - mu * m cancels (Nesterov)
        g = rng.standard_normal(n).astype(np.float32)
        gd.copy_(torch.from_numpy(g))


def test_momentum_sgd_rejects_bad_segments():
    from sad_b200 import ops
    t = torch.zeros(16, device="cuda")
    with pytest.raises(ValueError):
        ops.momentum_sgd(t, t.clone(), t.clone(), [(8, 1.0, 0.0)], torch.tensor(0.1, device="cuda"))
