"""GPU parity of the 3xTF32 ("f32x3") convolution mode — the fp32-accurate mode that matches the reference's fp32 head
convolution (cuDNN without tensor-op math, caffe2/caffe2/operators/conv_op_cudnn.cc:494-498) — against

  * oracle/conv_oracle.c, which tests/test_conv_reference_pin.py pins bit for bit to the reference's own CPU `Conv` /
    `ConvGradient` operators (conv_op_impl.h:31-180, 357-560), and
  * those operators themselves, run live from oracle/_ref/libref_ops.so on the same inputs.

Gate: BASELINE.json north_star's 1e-4: max|d| <= 1e-4 * max|ref| and relative rms <= 1e-4 for every output of every kernel
(forward, data gradient with the fused ReluGradient, weight gradient, bias gradient) and for the whole head.
Measured on a B200 (profiles/r02_f32x3_errors.txt): 2e-6 .. 2.4e-5 per kernel (the tensor core's accumulator add rounds toward
zero: the deviation grows with the number of MMAs in an accumulation chain), < 7e-5 for every tensor of the whole head.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MAX_TOL = 1e-4
RMS_TOL = 1e-4
MEASURED = []   # (what, max|d| / max|ref|, relative rms) of every comparison, dumped by the last test


def assert_close(got, ref, what, max_tol=MAX_TOL, rms_tol=RMS_TOL):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert np.isfinite(got).all(), what + ": non-finite output"
    m = np.abs(ref).max()
    d = np.abs(got - ref)
    rms = np.sqrt((d ** 2).mean()) / max(np.sqrt((ref ** 2).mean()), 1e-30)
    MEASURED.append((what, d.max() / max(m, 1e-30), rms))
    assert d.max() <= max_tol * m, "%s: max|d| %.3g > %.1e * max|ref| %.3g (at %s)" % (
        what, d.max(), max_tol, m, np.unravel_index(d.argmax(), d.shape))
    assert rms <= rms_tol, "%s: relative rms %.3g" % (what, rms)


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
    ((1, 32, 128, 8, 32), "one tile, one k-block per part"),
    ((2, 64, 128, 16, 64), "2x2 pixel tiles"),
    ((1, 256, 256, 20, 32), "head tower shape, ragged rows (P5)"),
    ((2, 256, 256, 5, 8), "P7"),
    ((1, 48, 36, 12, 20), "Cin=48 (split stride 64), Cout=36, ragged columns"),
    ((1, 64, 720, 8, 40), "cls_pred Cout=720: CTA pairs, last M tile partial"),
    ((1, 36, 64, 7, 14), "Cin=36 (the box data gradient's K), W % 4 != 0"),
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
    xs = ops.to_nhwc_f32x3([xd])
    # the split tensor carries the input to 2^-22
    assert_close(ops.join_split(xs[0], Cin).cpu().numpy(), x, "split layout " + name, 1e-6, 1e-6)
    pk = ops.conv3x3_pack_f32x3(wd, 0)
    got = ops.conv3x3_forward_f32x3(xs, pk, Cin, Cout, bd)[0][0]
    torch.cuda.synchronize()
    assert_close(got.cpu().numpy(), ref, "x3 conv fwd " + name)
    ys, yts, bits = ops.conv3x3_forward_f32x3(xs, pk, Cin, Cout, bd, relu=1, want_nhwc=True, want_bits=True)
    assert_close(ys[0].cpu().numpy(), oracle.relu(ref), "x3 conv+relu fwd " + name)
    # the split channels-last output is the NCHW output to 2^-22, and its pad channels stay zero
    assert_close(ops.join_split(yts[0], Cout).cpu().numpy(), ys[0].cpu().numpy(), "x3 split output " + name, 1e-6, 1e-6)
    cs = ops.split_channels(Cout)
    assert float(yts[0][..., Cout:cs].abs().max().item() if cs > Cout else 0.0) == 0.0
    got_nobias = ops.conv3x3_forward_f32x3(xs, pk, Cin, Cout, None)[0][0]
    assert_close(got_nobias.cpu().numpy(), oracle.conv2d_fwd(x, w, None), "x3 conv fwd (no bias) " + name)


@pytest.mark.parametrize("shape,name", CASES, ids=[c[1] for c in CASES])
def test_dgrad_with_fused_relu_gradient_matches_oracle(ops, oracle, shape, name):
    N, Cin, Cout, H, W = shape
    rng = np.random.default_rng(7 + abs(hash(shape)) % (2 ** 31))
    y_prev = _rand(rng, (N, Cin, H, W), relu_like=True)   # forward output of the layer below (post-ReLU)
    w = _rand(rng, (Cout, Cin, 3, 3), scale=1.0 / np.sqrt(9 * Cout))
    dy = _rand(rng, (N, Cout, H, W))
    _, _, ref_dx = oracle.conv2d_bwd(y_prev, w, dy)
    wd = torch.from_numpy(w).cuda()
    dys = ops.to_nhwc_f32x3([torch.from_numpy(dy).cuda()])
    pk1 = ops.conv3x3_pack_f32x3(wd, 1)
    got = ops.conv3x3_forward_f32x3(dys, pk1, Cout, Cin)[0][0]
    torch.cuda.synchronize()
    assert_close(got.cpu().numpy(), ref_dx, "x3 conv dgrad " + name)
    # ReluGradient from the sign bits of a forward pass that produced y_prev-shaped output: make them with an identity-free
    # route: bits of (y_prev > 0) written by a forward conv whose output IS y_prev is not available, so build them on the host
    segs = (W + 31) // 32
    words = np.zeros((N, H, segs, Cin), np.uint32)
    pos = (y_prev > 0)
    for i in range(W):
        words[:, :, i // 32, :] |= (pos[:, :, :, i].transpose(0, 2, 1).astype(np.uint32) << np.uint32(i % 32))
    bits = [torch.from_numpy(words.view(np.int32)).cuda()]
    ys, yts = ops.conv3x3_forward_f32x3(dys, pk1, Cout, Cin, want_nhwc=True, relu_bits=bits)
    ref = oracle.relu_grad(y_prev, ref_dx)
    assert_close(ys[0].cpu().numpy(), ref, "x3 dgrad+relu-grad " + name)
    assert_close(ops.join_split(yts[0], Cin).cpu().numpy(), ref, "x3 dgrad+relu-grad split output " + name)
    assert np.all(ys[0].cpu().numpy()[y_prev <= 0] == 0)


@pytest.mark.parametrize("shape,name", CASES, ids=[c[1] for c in CASES])
def test_wgrad_matches_oracle(ops, oracle, shape, name):
    N, Cin, Cout, H, W = shape
    rng = np.random.default_rng(11 + abs(hash(shape)) % (2 ** 31))
    x = _rand(rng, (N, Cin, H, W), relu_like=True)
    w = _rand(rng, (Cout, Cin, 3, 3), scale=0.05)
    dy = _rand(rng, (N, Cout, H, W))
    ref_dw, ref_db, _ = oracle.conv2d_bwd(x, w, dy, need_dx=False)
    xt = ops.to_nhwc_f32x3([torch.from_numpy(x).cuda()])
    dyt = ops.to_nhwc_f32x3([torch.from_numpy(dy).cuda()])
    dw, db = ops.conv3x3_wgrad_f32x3(xt, dyt, Cin, Cout)
    torch.cuda.synchronize()
    assert_close(dw.cpu().numpy(), ref_dw, "x3 conv wgrad " + name)
    assert_close(db.cpu().numpy(), ref_db, "x3 conv bias grad " + name)
    dw2, db2 = ops.conv3x3_wgrad_f32x3(xt, dyt, Cin, Cout)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)   # deterministic


def test_against_the_reference_cpu_operators_live(ops):
    """The same NetDef arguments through the reference's own CPU Conv / ConvGradient (oracle/_ref) and the 3xTF32 kernels."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_ref_cpu_conv_golden as gen
    from oracle import cpu_oracle
    from sad_b200 import c2
    if not os.path.exists(cpu_oracle.REF_GPU_LIB):
        pytest.skip("oracle/_ref/libref_ops.so not present")
    reflib = c2.OperatorLibrary(cpu_oracle.REF_GPU_LIB)
    N, Cin, Cout, H, W = 2, 256, 256, 10, 16
    rng = np.random.default_rng(41)
    x = _rand(rng, (N, Cin, H, W), relu_like=True, scale=0.5)
    w = _rand(rng, (Cout, Cin, 3, 3), scale=0.01)
    b = _rand(rng, (Cout,), scale=0.1)
    dy = _rand(rng, (N, Cout, H, W))
    ref = gen.run_reference(reflib, x, w, b, dy)
    xd, wd, bd, dyd = (torch.from_numpy(a).cuda() for a in (x, w, b, dy))
    xs, dys = ops.to_nhwc_f32x3([xd]), ops.to_nhwc_f32x3([dyd])
    y = ops.conv3x3_forward_f32x3(xs, ops.conv3x3_pack_f32x3(wd, 0), Cin, Cout, bd)[0][0]
    dx = ops.conv3x3_forward_f32x3(dys, ops.conv3x3_pack_f32x3(wd, 1), Cout, Cin)[0][0]
    dw, db = ops.conv3x3_wgrad_f32x3(xs, dys, Cin, Cout)
    torch.cuda.synchronize()
    assert_close(y.cpu().numpy(), ref["y"], "x3 vs reference CPU Conv: Y")
    assert_close(dx.cpu().numpy(), ref["dx"], "x3 vs reference CPU ConvGradient: dX")
    assert_close(dw.cpu().numpy(), ref["dw"], "x3 vs reference CPU ConvGradient: dW")
    assert_close(db.cpu().numpy(), ref["db"], "x3 vs reference CPU ConvGradient: db")


def test_head_f32x3_matches_fp32_references_config2_geometry():
    """Whole head, BASELINE.json configs[1] geometry (bs = 2, 600 px, 5 levels), compute_f32x3 = 1."""
    import test_head_gpu as th
    from sad_b200.head import RetinaNetHead
    shapes = [(80, 128), (40, 64), (20, 32), (10, 16), (5, 8)]
    head = RetinaNetHead(2, shapes, seed=7, compute_f32x3=True)
    g = torch.Generator(device="cuda").manual_seed(8)
    for name, p in head.params.items():
        if name.endswith("_w"):
            p.normal_(0.0, 1.0 / np.sqrt(9 * 256) * 1.4, generator=g)
        else:
            p.normal_(0.0, 0.1, generator=g)
    fpn = [torch.randn(2, 256, h, w, device="cuda", generator=g) for h, w in shapes]
    d_cls = [torch.randn(2, head.cls_out, h, w, device="cuda", generator=g) for h, w in shapes]
    d_box = [torch.randn(2, head.bbox_out, h, w, device="cuda", generator=g) for h, w in shapes]
    cls, box = head.forward(fpn)
    d_fpn = head.backward(d_cls, d_box)
    torch.cuda.synchronize()

    # (1) fp64 reference of the same graph, NO rounding emulation, ReLU masks taken from the product's kept activations
    # (an element whose pre-activation is within fp32 round-off of zero may legitimately fall on either side: with masks
    # fixed, every tensor must agree to 1e-4)
    class Exact(th.TorchF64Backend):
        def rna(self, t):
            return t.double()
    rcls, rbox, rg, rdx = th.staged_reference(head, fpn, d_cls, d_box, Exact(), product_acts=True)
    for l in range(len(shapes)):
        assert_close(th.to_np(cls[l]), th.to_np(rcls[l]), "x3 head cls logits level %d (fp64)" % l)
        assert_close(th.to_np(box[l]), th.to_np(rbox[l]), "x3 head bbox pred level %d (fp64)" % l)
        assert_close(th.to_np(d_fpn[l]), th.to_np(rdx[l]), "x3 head d_fpn level %d (fp64)" % l)
    for n in head.names:
        assert_close(th.to_np(head.grads[n]), th.to_np(rg[n]), "x3 head grad %s (fp64)" % n)

    # (2) plain fp32 autograd (cuDNN fp32, TF32 disabled: the reference's arithmetic class), nothing taken from the product.
    # Predictions: 1e-4.  Gradients: ANY two fp32 implementations of this graph disagree on the few ReLU masks whose pre-activation
    # is within round-off of zero, and one flipped mask moves the 9 * 256 input-gradient elements it feeds by a whole product
    # term (measured here: d_fpn max 3.8e-2 / rms 4.4e-3 of max|ref| while every tensor is < 7e-5 with the masks held fixed,
    # part (1)) — so this part is a statistical gate, not the parity gate.
    tcls, tbox, tg, tdx = th.torch_head(head, fpn, d_cls, d_box)
    for l in range(len(shapes)):
        assert_close(th.to_np(cls[l]), th.to_np(tcls[l]), "x3 head cls logits level %d (cuDNN fp32)" % l)
        assert_close(th.to_np(box[l]), th.to_np(tbox[l]), "x3 head bbox pred level %d (cuDNN fp32)" % l)
        assert_close(th.to_np(d_fpn[l]), th.to_np(tdx[l]), "x3 head d_fpn level %d (cuDNN fp32, free masks)" % l, max_tol=0.25, rms_tol=2e-2)
    for n in head.names:
        loose = "_conv_" in n
        assert_close(th.to_np(head.grads[n]), th.to_np(tg[n]), "x3 head grad %s (cuDNN fp32, free masks)" % n,
                     max_tol=0.25 if loose else 1e-4, rms_tol=2e-2 if loose else 1e-4)


def test_zz_dump_measured_errors():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "f32x3_errors.txt"), "w") as f:
        f.write("# 3xTF32 convolution mode: measured deviation from the fp32 references (tests/test_conv_f32x3_gpu.py)\n")
        f.write("# what | max|d| / max|ref| | relative rms\n")
        for what, mx, rms in MEASURED:
            f.write("%-72s %.3e %.3e\n" % (what, mx, rms))
    assert MEASURED
