"""GPU parity of the tcgen05 3x3 convolution (forward and data gradient) against the CPU oracle
(oracle/conv_oracle.c restating caffe2/caffe2/operators/conv_op_impl.h:31-180) and, at the head's full
size, against torch's fp32 convolution used as an independent implementation.

Tolerance: the tensor cores take tf32 operands (10-bit mantissa: activations truncated, weights rounded
to nearest) and accumulate in fp32, the oracle is fp32 throughout.  With K = 9*Cin products per
output the observed deviation is ~5e-4 of max|ref|; the gate is  max|d| <= 3e-3 * max|ref|  and
rms(d) <= 1e-3 * rms(ref).  The SIMT fallback path (C % 4 != 0) meets the same gate.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TF32_MAX_TOL = 3e-3
TF32_RMS_TOL = 1e-3


def assert_conv_close(got, ref, what):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert np.isfinite(got).all(), what + ": non-finite output"
    m = np.abs(ref).max()
    d = np.abs(got - ref)
    assert d.max() <= TF32_MAX_TOL * m, "%s: max|d| %.3g > %.1e * max|ref| %.3g (at %s)" % (
        what, d.max(), TF32_MAX_TOL, m, np.unravel_index(d.argmax(), d.shape))
    rms = np.sqrt((d ** 2).mean()) / max(np.sqrt((ref ** 2).mean()), 1e-30)
    assert rms <= TF32_RMS_TOL, "%s: relative rms %.3g" % (what, rms)


@pytest.fixture(scope="module")
def ops():
    from sad_b200 import ops as o
    assert torch.cuda.is_available()
    return o


def _rand(rng, shape, relu_like=False, scale=1.0):
    a = (rng.standard_normal(shape) * scale).astype(np.float32)
    return np.maximum(a, 0).astype(np.float32) if relu_like else a


CASES = [
    # (N, Cin, Cout, H, W)  name
    ((1, 32, 128, 8, 32), "one tile, one k-block per tap"),
    ((2, 64, 128, 16, 64), "2x2 pixel tiles"),
    ((1, 256, 256, 20, 32), "head tower shape, ragged rows (P5)"),
    ((2, 256, 256, 10, 16), "P6: half-width tile"),
    ((2, 256, 256, 5, 8), "P7"),
    ((1, 48, 36, 12, 20), "K tail (Cin=48), M tail (Cout=36), ragged columns"),
    ((1, 64, 720, 8, 40), "cls_pred Cout=720: 6 M tiles, last partial"),
    ((1, 32, 64, 7, 14), "W % 4 != 0: scalar NCHW stores"),
    ((1, 30, 20, 6, 10), "C % 4 != 0: SIMT fallback"),
]


@pytest.mark.parametrize("shape,name", CASES, ids=[c[1] for c in CASES])
def test_forward_matches_oracle(ops, oracle, shape, name):
    N, Cin, Cout, H, W = shape
    rng = np.random.default_rng(abs(hash(shape)) % (2 ** 31))
    x = _rand(rng, (N, Cin, H, W), relu_like=True)
    w = _rand(rng, (Cout, Cin, 3, 3), scale=1.0 / np.sqrt(9 * Cin))
    b = _rand(rng, (Cout,))
    ref = oracle.conv2d_fwd(x, w, b)
    xd, wd, bd = (torch.from_numpy(a).cuda() for a in (x, w, b))
    got = ops.conv3x3_forward([xd], wd, bd)[0][0]
    torch.cuda.synchronize()
    assert_conv_close(got.cpu().numpy(), ref, "conv fwd " + name)
    got_relu, got_cl = (r[0] for r in ops.conv3x3_forward([xd], wd, bd, relu=True, want_nhwc=True))
    assert_conv_close(got_relu.cpu().numpy(), oracle.relu(ref), "conv+relu fwd " + name)
    assert_conv_close(got_cl.permute(0, 3, 1, 2).cpu().numpy(), oracle.relu(ref), "conv+relu fwd channels-last " + name)
    got_nobias = ops.conv3x3_forward([xd], wd, None)[0][0]
    assert_conv_close(got_nobias.cpu().numpy(), oracle.conv2d_fwd(x, w, None), "conv fwd (no bias) " + name)


@pytest.mark.parametrize("shape,name", CASES[:6], ids=[c[1] for c in CASES[:6]])
def test_dgrad_matches_oracle(ops, oracle, shape, name):
    N, Cin, Cout, H, W = shape
    rng = np.random.default_rng(7 + abs(hash(shape)) % (2 ** 31))
    x = _rand(rng, (N, Cin, H, W), relu_like=True)
    w = _rand(rng, (Cout, Cin, 3, 3), scale=1.0 / np.sqrt(9 * Cout))
    dy = _rand(rng, (N, Cout, H, W))
    _, _, ref_dx = oracle.conv2d_bwd(x, w, dy)
    got = ops.conv3x3_dgrad([torch.from_numpy(dy).cuda()], torch.from_numpy(w).cuda())[0][0]
    torch.cuda.synchronize()
    assert_conv_close(got.cpu().numpy(), ref_dx, "conv dgrad " + name)


def test_impulse_response_is_exact(ops):
    # indexing check without rounding: x = one-hot pixels, weights = small integers (exact in tf32);
    # the output must equal the fp64 convolution exactly
    N, Cin, Cout, H, W = 1, 32, 128, 9, 36
    rng = np.random.default_rng(3)
    x = np.zeros((N, Cin, H, W), np.float32)
    for _ in range(40):
        x[0, rng.integers(Cin), rng.integers(H), rng.integers(W)] = float(rng.integers(1, 4))
    w = rng.integers(-3, 4, size=(Cout, Cin, 3, 3)).astype(np.float32)
    ref = torch.nn.functional.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), padding=1).numpy()
    got = ops.conv3x3_forward([torch.from_numpy(x).cuda()], torch.from_numpy(w).cuda())[0][0].cpu().numpy()
    assert np.array_equal(got.astype(np.float64), ref)


def test_all_levels_one_launch_head_shapes(ops):
    # BASELINE.json configs[1] geometry: bs = 2, 600 px pyramid, 256 -> 256 tower conv, all 5 levels in
    # one launch; reference = torch fp32 convolution (independent implementation, TF32 disabled)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(5)
    shapes = [(80, 128), (40, 64), (20, 32), (10, 16), (5, 8)]
    xs = [torch.randn(2, 256, h, w, device="cuda", generator=g).clamp_(min=0) for (h, w) in shapes]
    wt = torch.randn(256, 256, 3, 3, device="cuda", generator=g) / np.sqrt(9 * 256)
    b = torch.randn(256, device="cuda", generator=g)
    got = ops.conv3x3_forward(xs, wt, b, relu=True)[0]
    torch.cuda.synchronize()
    for x, y in zip(xs, got):
        ref = torch.relu(torch.nn.functional.conv2d(x, wt, b, padding=1))
        assert_conv_close(y.cpu().numpy(), ref.cpu().numpy(), "level %dx%d" % tuple(x.shape[2:]))


# ---------------------------------------------------------------------------------------------
# weight / bias gradient (tcgen05, pixels as the reduction axis) and the fused ReluGradient mask
# ---------------------------------------------------------------------------------------------
WGRAD_CASES = [
    # (N, Cin, Cout, H, W)
    ((1, 32, 128, 8, 32), "one tile, one segment per row"),
    ((2, 64, 128, 16, 64), "two segments per row"),
    ((1, 256, 256, 20, 32), "head tower shape (P5)"),
    ((2, 256, 256, 10, 16), "P6: half-filled pixel block"),
    ((2, 256, 256, 5, 8), "P7"),
    ((1, 48, 36, 12, 20), "Cin=48 / Cout=36 tails, ragged columns"),
    ((1, 64, 720, 8, 40), "cls_pred Cout=720: 6 M tiles, last partial"),
    ((1, 30, 20, 6, 10), "C % 4 != 0: SIMT path"),
]


@pytest.mark.parametrize("shape,name", WGRAD_CASES, ids=[c[1] for c in WGRAD_CASES])
def test_wgrad_matches_oracle(ops, oracle, shape, name):
    N, Cin, Cout, H, W = shape
    rng = np.random.default_rng(11 + abs(hash(shape)) % (2 ** 31))
    x = _rand(rng, (N, Cin, H, W), relu_like=True)
    w = _rand(rng, (Cout, Cin, 3, 3), scale=0.05)
    dy = _rand(rng, (N, Cout, H, W))
    ref_dw, ref_db, _ = oracle.conv2d_bwd(x, w, dy, need_dx=False)
    xt = ops.to_nhwc([torch.from_numpy(x).cuda()])
    dyt = ops.to_nhwc([torch.from_numpy(dy).cuda()])
    dw, db = ops.conv3x3_wgrad(xt, dyt)
    torch.cuda.synchronize()
    assert_conv_close(dw.cpu().numpy(), ref_dw, "conv wgrad " + name)
    assert_conv_close(db.cpu().numpy(), ref_db, "conv bias grad " + name)
    # run to run bit-identical (fixed-order split reduction, no atomics)
    dw2, db2 = ops.conv3x3_wgrad(xt, dyt)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
    # accumulate: adding the same gradient again doubles it
    ops.conv3x3_wgrad(xt, dyt, accumulate_into=(dw2, db2))
    assert_conv_close(dw2.cpu().numpy(), 2 * ref_dw, "conv wgrad accumulate " + name)
    assert_conv_close(db2.cpu().numpy(), 2 * ref_db, "conv bias grad accumulate " + name)


def test_wgrad_impulse_is_exact(ops):
    # indexing check without rounding: sparse integer x and dy (exact in tf32): dW must equal the fp64 result
    N, Cin, Cout, H, W = 2, 64, 160, 9, 36
    rng = np.random.default_rng(5)
    x = np.zeros((N, Cin, H, W), np.float32)
    dy = np.zeros((N, Cout, H, W), np.float32)
    for _ in range(300):
        x[rng.integers(N), rng.integers(Cin), rng.integers(H), rng.integers(W)] = float(rng.integers(1, 4))
        dy[rng.integers(N), rng.integers(Cout), rng.integers(H), rng.integers(W)] = float(rng.integers(-3, 4))
    xt64 = torch.from_numpy(x).double().requires_grad_(False)
    w64 = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv2d(xt64, w64, padding=1).backward(torch.from_numpy(dy).double())
    dw, db = ops.conv3x3_wgrad(ops.to_nhwc([torch.from_numpy(x).cuda()]), ops.to_nhwc([torch.from_numpy(dy).cuda()]))
    assert np.array_equal(dw.cpu().numpy().astype(np.float64), w64.grad.numpy())
    assert np.array_equal(db.cpu().numpy().astype(np.float64), dy.astype(np.float64).sum(axis=(0, 2, 3)))


def test_wgrad_sums_levels_in_one_launch(ops):
    # the head's weights are shared by the 5 FPN levels: one launch must equal the sum of the per-level
    # gradients (what Caffe2 autograd's Sum op computes); reference = torch fp32 autograd, TF32 disabled
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(9)
    shapes = [(80, 128), (40, 64), (20, 32), (10, 16), (5, 8)]
    xs = [torch.randn(2, 256, h, w, device="cuda", generator=g).clamp_(min=0) for (h, w) in shapes]
    dys = [torch.randn(2, 256, h, w, device="cuda", generator=g) for (h, w) in shapes]
    wt = torch.zeros(256, 256, 3, 3, device="cuda", requires_grad=True)
    b = torch.zeros(256, device="cuda", requires_grad=True)
    for x, dy in zip(xs, dys):
        torch.nn.functional.conv2d(x, wt, b, padding=1).backward(dy)
    dw, db = ops.conv3x3_wgrad(ops.to_nhwc(xs), ops.to_nhwc(dys))
    torch.cuda.synchronize()
    assert_conv_close(dw.cpu().numpy(), wt.grad.cpu().numpy(), "wgrad 5 levels")
    assert_conv_close(db.cpu().numpy(), b.grad.cpu().numpy(), "bias grad 5 levels")


@pytest.mark.parametrize("shape,name", CASES[:6] + CASES[8:], ids=[c[1] for c in CASES[:6] + CASES[8:]])
def test_dgrad_with_fused_relu_gradient(ops, oracle, shape, name):
    # tower backward: dX = ReluGradient(Y_prev, conv_dgrad(dY)) in one pass (relu_op.cu:29-35)
    N, Cin, Cout, H, W = shape
    rng = np.random.default_rng(23 + abs(hash(shape)) % (2 ** 31))
    y_prev = _rand(rng, (N, Cin, H, W), relu_like=True)   # forward output of the layer below (post-ReLU)
    w = _rand(rng, (Cout, Cin, 3, 3), scale=1.0 / np.sqrt(9 * Cout))
    dy = _rand(rng, (N, Cout, H, W))
    _, _, ref_dx = oracle.conv2d_bwd(y_prev, w, dy)
    ref = oracle.relu_grad(y_prev, ref_dx)
    mask = ops.to_nhwc([torch.from_numpy(y_prev).cuda()])
    got_nchw, got_cl = ops.conv3x3_dgrad([torch.from_numpy(dy).cuda()], torch.from_numpy(w).cuda(), want_nhwc=True,
                                         relu_masks_nhwc=mask)
    torch.cuda.synchronize()
    assert_conv_close(got_nchw[0].cpu().numpy(), ref, "dgrad+relu-grad " + name)
    assert_conv_close(got_cl[0].permute(0, 3, 1, 2).cpu().numpy(), ref, "dgrad+relu-grad channels-last " + name)
    # masked positions are exactly zero
    assert np.all(got_nchw[0].cpu().numpy()[y_prev <= 0] == 0)


@pytest.mark.parametrize("shape,name", [CASES[1], CASES[2], CASES[5], CASES[7], CASES[8]], ids=[c[1] for c in [CASES[1], CASES[2], CASES[5], CASES[7], CASES[8]]])
def test_sign_bits_roundtrip(ops, oracle, shape, name):
    # forward conv + ReLU leaves 1 bit per output element; the data-gradient pass of the layer above applies them
    # (ReluGradient) and must equal the float-mask form exactly
    N, Cin, Cout, H, W = shape
    rng = np.random.default_rng(31 + abs(hash(shape)) % (2 ** 31))
    x = _rand(rng, (N, Cin, H, W), relu_like=True)
    w = _rand(rng, (Cout, Cin, 3, 3), scale=1.0 / np.sqrt(9 * Cin))
    b = _rand(rng, (Cout,))
    xd, wd, bd = (torch.from_numpy(a).cuda() for a in (x, w, b))
    ys, yts, bits = ops.conv3x3_forward([xd], wd, bd, relu=True, want_nhwc=True, want_bits=True)
    # bit (n, y, seg, co) i  <->  y_nchw[n, co, y, 32 * seg + i] > 0
    yb = (ys[0] > 0).cpu().numpy()
    segs = (W + 31) // 32
    words = bits[0].cpu().numpy().view(np.uint32).reshape(N, H, segs, Cout)
    for i in range(min(W, 32 * segs)):
        got = (words[:, :, i // 32, :] >> np.uint32(i % 32)) & 1
        assert np.array_equal(got.astype(bool), yb[:, :, :, i].transpose(0, 2, 1)), "bit plane column %d" % i
    # next layer's data gradient (Cout2 -> Cout channels) masked by bits == masked by the float tensor
    w2 = _rand(rng, (40, Cout, 3, 3), scale=0.05)
    dy2 = _rand(rng, (N, 40, H, W))
    a_, _ = ops.conv3x3_dgrad([torch.from_numpy(dy2).cuda()], torch.from_numpy(w2).cuda(), relu_bits=bits)
    b_, _ = ops.conv3x3_dgrad([torch.from_numpy(dy2).cuda()], torch.from_numpy(w2).cuda(), relu_masks_nhwc=yts)
    torch.cuda.synchronize()
    assert torch.equal(a_[0], b_[0])
