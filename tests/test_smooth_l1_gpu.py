"""GPU parity of SelectSmoothL1Loss / SelectSmoothL1LossGradient through the C ABI and through the operator
registry, against the CPU oracle (oracle/focal_oracle.c restating select_smooth_l1_loss_op.cu:23-86, 90-181) and the UNMODIFIED
reference CUDA operators (oracle/_ref).  Loss: 1e-4 relative; gradient: which elements are touched is bit-exact (location
indexing), values within 1e-6 relative (a handful of fp32 operations per element)."""
import os

import numpy as np
import pytest
import torch

from parity import assert_loss_close

pytestmark = pytest.mark.gpu


def _case(seed, n, a, h, w, m):
    rng = np.random.default_rng(seed)
    y_hat = rng.normal(0, 0.3, size=(n, a * 4, h, w)).astype(np.float32)
    # M distinct foreground anchors: rows {image, first channel = 4 * anchor, y, x} as roi_data/retinanet.py builds them
    flat = rng.choice(n * a * h * w, size=m, replace=False)
    ni, rem = np.divmod(flat, a * h * w)
    ai, rem = np.divmod(rem, h * w)
    yi, xi = np.divmod(rem, w)
    locs = np.stack([ni, ai * 4, yi, xi], axis=1).astype(np.float32)
    y = rng.normal(0, 0.3, size=(m, 4)).astype(np.float32)
    return y_hat, y, locs


@pytest.mark.parametrize("shape,m,beta,scale", [((2, 9, 10, 16), 57, 0.11, 1.0), ((1, 3, 5, 7), 1, 1.0, 0.125),
                                                 ((2, 9, 20, 32), 700, 0.11, 0.5), ((1, 9, 4, 4), 0, 0.11, 1.0)])
def test_select_smooth_l1_matches_oracle(oracle, shape, m, beta, scale):
    from sad_b200 import ops
    n, a, h, w = shape
    y_hat, y, locs = _case(m + 1, n, a, h, w, m)
    fg = float(max(m, 1) * 1.7)
    ref_loss, ref_grad = oracle.select_smooth_l1(y_hat, y, locs, fg, beta=beta, scale=scale, d_loss=0.6)
    yd = torch.from_numpy(y).cuda() if m else torch.empty(0, 4, device="cuda")
    ld = torch.from_numpy(locs).cuda() if m else torch.empty(0, 4, device="cuda")
    loss, grad = ops.select_smooth_l1_loss(torch.from_numpy(y_hat).cuda(), yd, ld, torch.tensor([fg], device="cuda"), beta=beta, scale=scale,
                                           d_loss=torch.tensor(0.6, device="cuda"))
    torch.cuda.synchronize()
    assert_loss_close(loss.item(), ref_loss) if m else None
    if m == 0:
        assert loss.item() == 0.0
    g = grad.cpu().numpy()
    assert np.array_equal(g != 0, ref_grad != 0), "the set of touched elements must follow the location rows exactly"
    np.testing.assert_allclose(g, ref_grad, rtol=2e-6, atol=0)


def test_select_smooth_l1_operators_against_unmodified_reference(oracle):
    from oracle import cpu_oracle
    from sad_b200 import c2
    if not os.path.exists(cpu_oracle.REF_GPU_LIB):
        pytest.skip("oracle/_ref/libref_ops.so not built")
    y_hat, y, locs = _case(5, 2, 9, 10, 16, 123)
    dev = c2.DeviceOption(c2.CUDA, 0)
    out = {}
    for name, lib in (("product", c2.OperatorLibrary()), ("reference", c2.OperatorLibrary(cpu_oracle.REF_GPU_LIB))):
        if not lib.HasOperator("SelectSmoothL1Loss", c2.CUDA):
            pytest.skip("oracle/_ref predates the smooth-L1 sources")
        ws = lib.Workspace()
        for blob, arr in (("Yh", y_hat), ("Y", y), ("L", locs)):
            ws.FeedBlob(blob, torch.from_numpy(arr).cuda())
        ws.FeedBlob("S", torch.tensor([200.0], device="cuda"))
        ws.FeedBlob("loss_grad", torch.tensor(1.0, device="cuda"))
        ws.RunOperatorOnce(c2.CreateOperator("SelectSmoothL1Loss", ["Yh", "Y", "L", "S"], ["loss"], device_option=dev, beta=0.11, scale=0.25))
        ws.RunOperatorOnce(c2.CreateOperator("SelectSmoothL1LossGradient", ["Yh", "Y", "L", "S", "loss_grad"], ["dYh"], device_option=dev,
                                             beta=0.11, scale=0.25))
        out[name] = (ws.FetchBlob("loss"), ws.FetchBlob("dYh"))
    assert out["product"][0].shape == ()
    assert_loss_close(out["product"][0], out["reference"][0], "loss product vs reference")
    np.testing.assert_allclose(out["product"][1], out["reference"][1], rtol=2e-6, atol=0)
    ref_loss, ref_grad = oracle.select_smooth_l1(y_hat, y, locs, 200.0, beta=0.11, scale=0.25)
    assert_loss_close(ref_loss, out["reference"][0], "oracle vs reference loss")
    np.testing.assert_allclose(ref_grad, out["reference"][1], rtol=2e-6, atol=0)
