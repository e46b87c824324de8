"""GPU parity of SigmoidFocalLoss / SigmoidFocalLossGradient through the C ABI and through the
operator registry, against the CPU oracle (oracle/focal_oracle.c restating sigmoid_focal_loss_op.cu:26-109) and — when
oracle/_ref was built — against the UNMODIFIED reference CUDA operators run side by side on the same inputs.
Gates: tests/parity.py (loss 1e-4 relative; gradient max 1e-4 of max|ref| + elementwise with floor); label/class
indexing bit-exact through the zero pattern of the gradient."""
import os

import numpy as np
import pytest
import torch

from parity import assert_grad_close, assert_loss_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from sad_b200 import ops as o
    return o


def _level(seed, n, a, c, h, w, fg_frac=0.02, stress=False):
    rng = np.random.default_rng(seed)
    x = (rng.uniform(-12, 12, size=(n, a * c, h, w)) if stress else rng.normal(-4.595, 2.0, size=(n, a * c, h, w))).astype(np.float32)
    u = rng.random(size=(n, a, h, w))
    g = np.zeros((n, a, h, w), np.int32)
    g[u < 0.05] = -1
    fg = (u >= 0.05) & (u < 0.05 + fg_frac)
    g[fg] = rng.integers(1, c + 1, size=int(fg.sum()), dtype=np.int32)
    return x, g, float(max(1, fg.sum()))


CASES = [
    # (n, a, c, h, w), gamma, alpha, scale, stress
    ((2, 9, 80, 8, 16), 2.0, 0.25, 1.0, False),      # RetinaNet defaults (RETINANET.LOSS_GAMMA / LOSS_ALPHA)
    ((1, 3, 5, 5, 7), 2.0, 0.25, 0.125, True),       # H*W % 4 != 0: scalar path; scale = 1 / NUM_GPUS
    ((2, 9, 80, 4, 8), 1.0, 0.5, 1.0, True),         # operator defaults gamma = 1
    ((1, 2, 4, 6, 8), 0.0, 0.75, 2.0, False),        # gamma = 0: plain weighted cross-entropy
    ((1, 2, 4, 6, 8), 3.5, 0.25, 1.0, True),
]


@pytest.mark.parametrize("shape,gamma,alpha,scale,stress", CASES)
def test_focal_matches_oracle(ops, oracle, shape, gamma, alpha, scale, stress):
    n, a, c, h, w = shape
    x, g, fg = _level(abs(hash((shape, gamma))) % 2 ** 31, n, a, c, h, w, stress=stress)
    kw = dict(gamma=gamma, alpha=alpha, scale=scale, num_classes=c)
    ref_loss = oracle.focal_loss(x, g, fg, **kw)
    ref_grad = oracle.focal_grad(x, g, fg, d_loss=0.7, **kw)
    xd, gd = torch.from_numpy(x).cuda(), torch.from_numpy(g).cuda()
    fgd, dl = torch.tensor([fg], device="cuda"), torch.tensor(0.7, device="cuda")
    loss, grad = ops.sigmoid_focal_loss(xd, gd, fgd, d_loss=dl, **kw)
    torch.cuda.synchronize()
    assert_loss_close(loss.item(), ref_loss)
    assert_grad_close(grad.cpu().numpy(), ref_grad)
    # the three entry modes agree, and accumulate adds
    only_l, _ = ops.sigmoid_focal_loss(xd, gd, fgd, want_grad=False, **kw)
    _, only_g = ops.sigmoid_focal_loss(xd, gd, fgd, want_loss=False, d_loss=dl, **kw)
    assert only_l.item() == loss.item() and torch.equal(only_g, grad)
    base = torch.full_like(xd, 0.25)
    ops.sigmoid_focal_loss(xd, gd, fgd, want_loss=False, d_loss=dl, accumulate_into=base, **kw)
    assert torch.allclose(base, grad + 0.25, rtol=0, atol=1e-7)
    # run-to-run bit-identical loss
    assert ops.sigmoid_focal_loss(xd, gd, fgd, want_grad=False, **kw)[0].item() == loss.item()


def test_focal_class_and_label_indexing_is_bit_exact(ops):
    # which (anchor, class, y, x) elements are positive / negative / ignored follows sigmoid_focal_loss_op.cu:32-43 exactly:
    # with x = 0 everywhere the gradient takes exactly three values (positive class, negative class, ignored = 0)
    n, a, c, h, w = 2, 3, 6, 4, 8
    rng = np.random.default_rng(4)
    g = rng.integers(-1, c + 1, size=(n, a, h, w)).astype(np.int32)
    x = np.zeros((n, a * c, h, w), np.float32)
    _, grad = ops.sigmoid_focal_loss(torch.from_numpy(x).cuda(), torch.from_numpy(g).cuda(), torch.tensor([1.0], device="cuda"),
                                     want_loss=False, gamma=2.0, alpha=0.25, num_classes=c)
    grad = grad.cpu().numpy().reshape(n, a, c, h, w)
    t = g[:, :, None, :, :]
    cls = np.arange(1, c + 1).reshape(1, 1, c, 1, 1)
    pos, ign = (t == cls), np.broadcast_to(t == -1, grad.shape)
    assert np.all(grad[ign] == 0)
    assert np.all(grad[pos] < 0) and np.all(grad[~pos & ~ign] > 0)
    assert len(np.unique(grad[pos])) == 1 and len(np.unique(grad[~pos & ~ign])) == 1


def test_focal_config2_size_against_exact_sum(ops, oracle):
    # BASELINE.json configs[1] geometry, finest level: the reference's single-block sum is itself ~1e-4 off the exact sum
    # at this size, so the gate is: within 2e-5 of the fp64 sum of the oracle's per-element terms
    x, g, fg = _level(99, 2, 9, 80, 80, 128, fg_frac=0.001)
    kw = dict(gamma=2.0, alpha=0.25, scale=1.0, num_classes=80)
    _, elems = oracle.focal_loss(x, g, fg, return_elements=True, **kw)
    exact = float(elems.astype(np.float64).sum())
    loss, grad = ops.sigmoid_focal_loss(torch.from_numpy(x).cuda(), torch.from_numpy(g).cuda(), torch.tensor([fg], device="cuda"), **kw)
    assert abs(loss.item() - exact) <= 2e-5 * abs(exact)
    assert_grad_close(grad.cpu().numpy(), oracle.focal_grad(x, g, fg, **kw))


def test_focal_operators_by_name_against_unmodified_reference(oracle):
    """Same OperatorDefs, same inputs, two operator libraries: the product and the reference's own .cu (oracle/_ref)."""
    from oracle import cpu_oracle
    from sad_b200 import c2
    if not os.path.exists(cpu_oracle.REF_GPU_LIB):
        pytest.skip("oracle/_ref/libref_ops.so not built")
    x, g, fg = _level(7, 2, 9, 80, 10, 12, stress=True)
    dev = c2.DeviceOption(c2.CUDA, 0)
    kw = dict(gamma=2.0, alpha=0.25, scale=0.5, num_classes=80)
    out = {}
    for name, lib in (("product", c2.OperatorLibrary()), ("reference", c2.OperatorLibrary(cpu_oracle.REF_GPU_LIB))):
        if not lib.HasOperator("SigmoidFocalLoss", c2.CUDA):
            pytest.skip("oracle/_ref predates the focal-loss sources")
        ws = lib.Workspace()
        ws.FeedBlob("X", torch.from_numpy(x).cuda())
        ws.FeedBlob("G", torch.from_numpy(g).cuda())
        ws.FeedBlob("fg", torch.tensor([fg], device="cuda"))
        ws.FeedBlob("loss_grad", torch.tensor(1.0, device="cuda"))
        ws.RunOperatorOnce(c2.CreateOperator("SigmoidFocalLoss", ["X", "G", "fg"], ["loss"], device_option=dev, **kw))
        ws.RunOperatorOnce(c2.CreateOperator("SigmoidFocalLossGradient", ["X", "G", "fg", "loss_grad"], ["dX"], device_option=dev, **kw))
        out[name] = (ws.FetchBlob("loss"), ws.FetchBlob("dX"))
    assert out["product"][0].shape == ()
    assert_loss_close(out["product"][0], out["reference"][0], "loss product vs reference")
    assert_grad_close(out["product"][1], out["reference"][1], "grad product vs reference")
    assert_loss_close(oracle.focal_loss(x, g, fg, **kw), out["reference"][0], "oracle vs reference loss")
    assert_grad_close(oracle.focal_grad(x, g, fg, **kw), out["reference"][1], "oracle vs reference grad")
