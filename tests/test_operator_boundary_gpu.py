"""GPU: the path driven the way the reference drives it — operators created by NAME from
OperatorDef / NetDef text through the registry, blobs in a Workspace — checked against the oracle
and, when oracle/_ref was built, against the UNMODIFIED reference CUDA operators run side by side."""
import os

import numpy as np
import pytest
import torch

from parity import assert_grad_close, assert_loss_close

pytestmark = pytest.mark.gpu
HEAD = dict(gamma=2.0, alpha=0.5, beta=0.0, scale=1.0, num_classes=80, ignored_label=-1)


@pytest.fixture(scope="module")
def oplib():
    from sad_b200 import c2
    return c2.OperatorLibrary()


@pytest.fixture(scope="module")
def reflib():
    from oracle import cpu_oracle
    from sad_b200 import c2
    if not os.path.exists(cpu_oracle.REF_GPU_LIB):
        pytest.skip("oracle/_ref/libref_ops.so not built (needs /root/reference at build time)")
    return c2.OperatorLibrary(cpu_oracle.REF_GPU_LIB)


def _pyramid(n=2):
    from sad_b200 import synthetic
    shapes = [(20, 32), (10, 16), (5, 8), (3, 4), (2, 2)]
    return [synthetic.make_level(np.random.default_rng(300 + i), n, h, w) for i, (h, w) in enumerate(shapes)]


def _run_net(lib, host, fuse):
    from sad_b200 import c2, retinanet_heads
    net, losses, grads = retinanet_heads.add_distill_loss(gpu_id=0, num_gpus=1)
    text = net.to_text()
    if fuse:
        text, n = lib.FuseAdaptiveDistillOps(text)
        assert n == 1
    ws = lib.Workspace()
    dev = [tuple(torch.from_numpy(a).cuda() for a in l) for l in host]
    retinanet_heads.feed_level_blobs(ws, 0, dev)
    ws.CreateNet(text)
    ws.RunNet(net.name)
    ws.RunNet(net.name)  # operators keep their scratch across runs
    return ([ws.FetchBlob(b) for b in losses], [ws.FetchBlob(b) for b in grads], ws.FetchBlob("gpu_0/distill_normalizer"))


def test_single_operators_by_name(oplib, oracle):
    from sad_b200 import c2
    host = _pyramid(1)[:2]
    dev = c2.DeviceOption(c2.CUDA, 0)
    ws = oplib.Workspace()
    for i, (x, t, g) in enumerate(host):
        ws.FeedBlob("X%d" % i, torch.from_numpy(x).cuda())
        ws.FeedBlob("T%d" % i, torch.from_numpy(t).cuda())
        ws.FeedBlob("G%d" % i, torch.from_numpy(g).cuda())
    ws.RunOperatorOnce(c2.CreateOperator("PowSum", ["T0", "T1"], ["wp"], device_option=dev, power=1.8))
    wp = oracle.pow_sum([host[0][1], host[1][1]], 1.8)
    got = ws.FetchBlob("wp")
    assert got.shape == () and got.dtype == np.float32   # Resize(vector<TIndex>()): a scalar
    assert_loss_close(got, wp)
    ws.RunOperatorOnce(c2.CreateOperator("SigmoidAdaptiveDistillLoss", ["X0", "T0", "G0", "wp"], ["loss"], device_option=dev, **HEAD))
    ws.FeedBlob("loss_grad", torch.tensor(1.0, device="cuda"))
    ws.RunOperatorOnce(c2.CreateOperator("SigmoidAdaptiveDistillLossGradient", ["X0", "T0", "G0", "wp", "loss_grad"], ["dX"],
                                         device_option=dev, **HEAD))
    assert ws.FetchBlob("loss").shape == ()
    assert_loss_close(ws.FetchBlob("loss"), oracle.distill_loss(*host[0], wp, **HEAD))
    dX = ws.FetchBlob("dX")
    assert dX.shape == host[0][0].shape
    assert_grad_close(dX, oracle.distill_grad(*host[0], wp, **HEAD))
    # default arguments are the reference's (gamma 1, alpha 0.25, scale 1, num_classes 80, ignored -1)
    ws.RunOperatorOnce(c2.CreateOperator("SigmoidAdaptiveDistillLoss", ["X0", "T0", "G0", "wp"], ["loss_d"], device_option=dev))
    assert_loss_close(ws.FetchBlob("loss_d"), oracle.distill_loss(*host[0], wp))
    # wrong label dtype is a type-mismatch enforce naming the blob (tensor.h:500-506)
    ws.FeedBlob("Gf", torch.zeros(host[0][2].shape, device="cuda"))
    with pytest.raises(c2.EnforceNotMet, match="Gf"):
        ws.RunOperatorOnce(c2.CreateOperator("SigmoidAdaptiveDistillLoss", ["X0", "T0", "Gf", "wp"], ["l2"], device_option=dev, **HEAD))


def test_reference_graph_unfused_and_fused(oplib, oracle):
    host = _pyramid(2)
    wp = oracle.pow_sum([l[1] for l in host], 1.8)
    lu, gu, nu = _run_net(oplib, host, fuse=False)
    lf, gf, nf = _run_net(oplib, host, fuse=True)
    assert_loss_close(nu, wp)
    # the pass folds PowSum + 5 losses + 5 gradients into ONE SigmoidAdaptiveDistillStep op (one cooperative launch).
    # It does not change results beyond summation order: the normaliser and the losses are sums whose partition over
    # CTAs differs (last-bit differences); the gradients use the same element arithmetic and inherit the normaliser's
    # last-bit difference as a common factor
    assert abs(float(nu) - float(nf)) <= 2e-6 * abs(float(nu))
    for i, l in enumerate(host):
        assert_loss_close(lu[i], oracle.distill_loss(*l, wp, **HEAD), "level %d" % i)
        assert_grad_close(gu[i], oracle.distill_grad(*l, wp, **HEAD), "level %d" % i)
        assert_loss_close(lf[i], oracle.distill_loss(*l, wp, **HEAD), "fused level %d" % i)
        assert_grad_close(gf[i], oracle.distill_grad(*l, wp, **HEAD), "fused level %d" % i)
        np.testing.assert_allclose(gf[i], gu[i], rtol=1e-5, atol=0)
        assert abs(float(lu[i]) - float(lf[i])) <= 1e-5 * abs(float(lu[i]))


def test_against_unmodified_reference_cuda_ops(oplib, reflib, oracle):
    """Same NetDef text, same inputs, two operator libraries: the product and the reference's own
    .cu files (oracle/_ref).  Also pins the CPU oracle against the reference itself."""
    host = _pyramid(2)
    lp, gp, np_ = _run_net(oplib, host, fuse=False)
    lr, gr, nr = _run_net(reflib, host, fuse=False)
    assert_loss_close(np_, nr, "PowSum product vs reference")
    wp = float(nr)
    for i, l in enumerate(host):
        assert_loss_close(lp[i], lr[i], "loss level %d product vs reference" % i)
        assert_grad_close(gp[i], gr[i], "grad level %d product vs reference" % i)
        assert_loss_close(oracle.distill_loss(*l, wp, **HEAD), lr[i], "oracle vs reference loss %d" % i)
        assert_grad_close(oracle.distill_grad(*l, wp, **HEAD), gr[i], "oracle vs reference grad %d" % i)


def test_reference_ops_on_golden_inputs_match_committed_vectors(reflib, oracle):
    """The committed tests/golden/ref_gpu_kat.npz must be what the reference ops produce here."""
    from sad_b200 import c2
    path = os.path.join(os.path.dirname(__file__), "golden", "ref_gpu_kat.npz")
    if not os.path.exists(path):
        pytest.skip("ref_gpu_kat.npz not generated yet")
    import golden.make_ref_gpu_golden as gen
    fresh = gen.run_reference_ops(reflib)
    stored = np.load(path)
    for k in stored.files:
        np.testing.assert_allclose(fresh[k], stored[k], rtol=1e-6, atol=1e-12, err_msg=k)


# ---------------------------------------------------------------------------------------------
# head convolutions through the registry: Conv (engine CUDNN, as Detectron requests) -> in-place Relu -> Conv,
# then the gradient ops the makers emit, executed in reverse order — one FPN level of the classification
# branch as Detectron builds it (retinanet_heads.py:101-152), against the CPU oracle.
# ---------------------------------------------------------------------------------------------
def _conv_close(got, ref, what, max_tol=3e-3, rms_tol=1e-3):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    d = np.abs(got - ref)
    assert d.max() <= max_tol * np.abs(ref).max(), "%s: max|d| %.3g vs max|ref| %.3g" % (what, d.max(), np.abs(ref).max())
    assert np.sqrt((d ** 2).mean()) <= rms_tol * np.sqrt((ref ** 2).mean()), what


@pytest.mark.parametrize("tensor_core", [None, 1], ids=["default: fp32-accurate (3xTF32)", "enable_tensor_core=1: tf32"])
def test_head_conv_ops_by_name_forward_and_gradient(oplib, oracle, tensor_core):
    """Conv / ConvGradient / Relu / ReluGradient through the registry.  Without enable_tensor_core the operators compute like the
    reference's fp32 convolution (gate 1e-4 max / 1e-4 rms against the reference-pinned oracle); with enable_tensor_core = 1
    (conv_op_cudnn.cc:77-86) in single-pass tf32 (gate 3e-3 / 1e-3)."""
    from sad_b200 import c2
    tol = dict(max_tol=1e-4, rms_tol=1e-4) if tensor_core is None else {}
    rng = np.random.default_rng(77)
    N, C, M, H, W = 2, 64, 36, 10, 24
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    w0 = (rng.standard_normal((C, C, 3, 3)) / np.sqrt(9 * C)).astype(np.float32)
    b0 = (0.1 * rng.standard_normal(C)).astype(np.float32)
    w1 = (rng.standard_normal((M, C, 3, 3)) / np.sqrt(9 * C)).astype(np.float32)
    b1 = (0.1 * rng.standard_normal(M)).astype(np.float32)
    dy = rng.standard_normal((N, M, H, W)).astype(np.float32)
    dev = c2.DeviceOption(c2.CUDA, 0)
    conv_args = dict(kernel=3, pad=1, stride=1, order="NCHW")
    if tensor_core is not None:
        conv_args["enable_tensor_core"] = tensor_core
    fwd = [c2.CreateOperator("Conv", ["fpn", "w0", "b0"], ["t"], device_option=dev, engine="CUDNN", **conv_args),
           c2.CreateOperator("Relu", ["t"], ["t"], device_option=dev),
           c2.CreateOperator("Conv", ["t", "w1", "b1"], ["pred"], device_option=dev, engine="CUDNN", **conv_args)]
    ws = oplib.Workspace()
    for name, a in (("fpn", x), ("w0", w0), ("b0", b0), ("w1", w1), ("b1", b1), ("pred_grad", dy)):
        ws.FeedBlob(name, torch.from_numpy(a).cuda())
    for op in fwd:
        ws.RunOperatorOnce(op)
    t_ref = oracle.relu(oracle.conv2d_fwd(x, w0, b0))
    pred_ref = oracle.conv2d_fwd(t_ref, w1, b1)
    _conv_close(ws.FetchBlob("t"), t_ref, "tower activation", **tol)
    _conv_close(ws.FetchBlob("pred"), pred_ref, "prediction", **tol)
    # backward: gradient defs from the registered makers, run through the same registry
    g_out = {"pred": "pred_grad"}
    for op in reversed(fwd):
        text = oplib.GetGradientDefs(op, [g_out.get(o, "") for o in op.output])
        ws.CreateNet('name: "g"\n' + "\n".join(l for l in text.splitlines() if not l.startswith("external_output")), overwrite=True)
        ws.RunNet("g")
        gin = [l.split('"')[1] for l in text.splitlines() if l.startswith("external_output")]
        for name, g in zip(op.input, gin):
            if g:
                g_out[name] = g
    dw1, db1, dt = oracle.conv2d_bwd(ws.FetchBlob("t"), w1, dy)        # masks from the product's own activation
    dt = oracle.relu_grad(ws.FetchBlob("t"), dt)
    dw0, db0, dx = oracle.conv2d_bwd(x, w0, dt)
    for name, ref in (("w1_grad", dw1), ("b1_grad", db1), ("w0_grad", dw0), ("b0_grad", db0), ("fpn_grad", dx)):
        _conv_close(ws.FetchBlob(name), ref, name, **tol)


def test_head_conv_op_rejects_shapes_outside_its_class(oplib):
    from sad_b200 import c2
    dev = c2.DeviceOption(c2.CUDA, 0)
    ws = oplib.Workspace()
    ws.FeedBlob("x", torch.zeros(1, 8, 4, 4, device="cuda"))
    ws.FeedBlob("w", torch.zeros(8, 8, 3, 3, device="cuda"))
    for bad in (dict(kernel=1, pad=0, stride=1), dict(kernel=3, pad=1, stride=2), dict(kernel=3, pad=1, stride=1, group=2),
                dict(kernel=3, pad=1, stride=1, order="NHWC")):
        with pytest.raises(c2.EnforceNotMet, match="B200 head convolution"):
            ws.RunOperatorOnce(c2.CreateOperator("Conv", ["x", "w"], ["y"], device_option=dev, **bad))
    ws.FeedBlob("w5", torch.zeros(8, 4, 3, 3, device="cuda"))
    with pytest.raises(c2.EnforceNotMet, match="channels"):
        ws.RunOperatorOnce(c2.CreateOperator("Conv", ["x", "w5"], ["y"], device_option=dev, kernel=3, pad=1, stride=1))


@pytest.mark.parametrize("tensor_core", [1, None], ids=["enable_tensor_core=1 vs the tf32 head", "default vs the 3xTF32 head"])
def test_head_netdef_through_the_operators_equals_the_fused_head_object(oplib, tensor_core):
    # retinanet_heads.py:63-245 emitted op by op (90 Conv / Relu operators, ConvShared levels reading level 3's blobs) against
    # sad_head_forward, which runs the same graph as 10 launches.  Same kernels, same packed tf32 weights, same rounding points
    # (activations are rounded when their channels-last copy is written), so the results agree to fp32 round-off.
    from sad_b200 import c2, retinanet_heads
    from sad_b200.head import RetinaNetHead
    shapes, n, dim = [(16, 24), (8, 12), (4, 6), (2, 3), (1, 2)], 2, 32
    x3 = tensor_core is None
    student = RetinaNetHead(n, shapes, dim=dim, seed=9, compute_f32x3=x3)
    teacher = RetinaNetHead(n, shapes, dim=dim, seed=9, cls_output_sigmoid=True, compute_f32x3=x3)
    g = torch.Generator(device="cuda").manual_seed(4)
    for name, p in student.params.items():
        p.normal_(0.0, 0.08 if name.endswith("_w") else 0.1, generator=g)
    teacher.flat_params.copy_(student.flat_params)
    fpn = [torch.randn(n, dim, h, w, device="cuda", generator=g).clamp_(min=0) for h, w in shapes]
    cls, box = student.forward(fpn, training=False)
    prob, _ = teacher.forward(fpn, training=False)
    blobs_in = ["gpu_0/fpn_%d" % l for l in (7, 6, 5, 4, 3)]
    for train, scope, ref_cls in ((True, "", cls), (False, "teacher/", prob)):
        net, params, cls_out, box_out = retinanet_heads.add_fpn_retinanet_outputs(
            [b.replace("gpu_0/", "gpu_0/" + scope) for b in blobs_in], train=train, dim_in=dim, scope=scope, enable_tensor_core=tensor_core)
        ws = oplib.Workspace()
        for l, f in zip((3, 4, 5, 6, 7), fpn):
            ws.FeedBlob("gpu_0/%sfpn_%d" % (scope, l), f)
        for name, shape, _ in params:
            p = student.params[name.split("/")[-1]]
            assert tuple(p.shape) == shape
            ws.FeedBlob(name, p)
        ws.CreateNet(net.to_text())
        ws.RunNet(net.name)
        for l in range(5):
            got_cls, got_box = ws.FetchBlob(cls_out[l]), ws.FetchBlob(box_out[l])
            assert got_cls.shape == tuple(ref_cls[l].shape) and got_box.shape == tuple(box[l].shape)
            assert np.abs(got_cls - ref_cls[l].cpu().numpy()).max() <= 2e-5 * max(1.0, float(ref_cls[l].abs().max()))
            assert np.abs(got_box - box[l].cpu().numpy()).max() <= 2e-5 * max(1.0, float(box[l].abs().max()))
        if not train:
            assert float(ws.FetchBlob(cls_out[0]).min()) > 0.0 and float(ws.FetchBlob(cls_out[0]).max()) < 1.0


def test_sigmoid_operator(oplib):
    from sad_b200 import c2
    dev = c2.DeviceOption(c2.CUDA, 0)
    ws = oplib.Workspace()
    x = torch.linspace(-30.0, 30.0, 4099, device="cuda")          # odd length: the scalar tail runs too
    ws.FeedBlob("x", x)
    ws.RunOperatorOnce(c2.CreateOperator("Sigmoid", ["x"], ["y"], device_option=dev))
    ref = 1.0 / (1.0 + np.exp(-x.double().cpu().numpy()))
    assert np.abs(ws.FetchBlob("y") - ref).max() <= 1e-6
    ws.RunOperatorOnce(c2.CreateOperator("Sigmoid", ["x"], ["x"], device_option=dev))   # in place (AllowInplace {0, 0})
    assert np.abs(ws.FetchBlob("x") - ref).max() <= 1e-6


def test_head_loss_netdef_runs_through_the_operators(oplib, oracle):
    from sad_b200 import c2, retinanet_heads
    rng = np.random.default_rng(12)
    net, losses = retinanet_heads.add_fpn_retinanet_losses(gpu_id=0, num_gpus=2)
    ws = oplib.Workspace()
    shapes = [(8, 12), (4, 6), (2, 3), (1, 2), (1, 1)]
    host, fg_total = [], 0
    for l, (h, w) in zip(range(3, 8), shapes):
        logits = rng.normal(-2.0, 2.0, (2, 720, h, w)).astype(np.float32)
        labels = rng.integers(-1, 81, (2, 9, h, w)).astype(np.int32)
        labels[rng.random(labels.shape) < 0.9] = 0
        box = rng.normal(0, 1, (2, 36, h, w)).astype(np.float32)
        fg = np.argwhere(labels > 0)
        locs = np.stack([fg[:, 0], fg[:, 1] * 4, fg[:, 2], fg[:, 3]], axis=1).astype(np.float32).reshape(-1, 4)
        tgt = rng.normal(0, 0.3, (locs.shape[0], 4)).astype(np.float32)
        fg_total += locs.shape[0]
        host.append((logits, labels, box, locs, tgt))
    fg_num = np.array([float(max(fg_total, 1))], dtype=np.float32)
    ws.FeedBlob("gpu_0/retnet_fg_num", torch.from_numpy(fg_num).cuda())
    for l, (logits, labels, box, locs, tgt) in zip(range(3, 8), host):
        for name, a in (("retnet_cls_pred", logits), ("retnet_cls_labels", labels), ("retnet_bbox_pred", box),
                        ("retnet_roi_fg_bbox_locs", locs), ("retnet_roi_bbox_targets", tgt)):
            ws.FeedBlob("gpu_0/%s_fpn%d" % (name, l), torch.from_numpy(a).cuda())
    ws.CreateNet(net.to_text())
    ws.RunNet(net.name)
    for i, (logits, labels, box, locs, tgt) in enumerate(host):
        ref_fl = oracle.focal_loss(logits, labels, float(fg_num[0]), gamma=2.0, alpha=0.25, scale=0.5, num_classes=80)
        assert_loss_close(ws.FetchBlob(losses[5 + i]), ref_fl, "focal level %d" % i)
        ref_box = oracle.select_smooth_l1(box, tgt, locs, float(fg_num[0]), beta=0.11, scale=0.5)[0]
        assert_loss_close(ws.FetchBlob(losses[i]), ref_box, "box level %d" % i)
