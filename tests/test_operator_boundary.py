"""CPU: host logic of the operator boundary — registration under the reference's names, schema
arity, argument typing, gradient maker, NetDef text round trip, graph wiring and the fusion pass.
No kernels run here."""
import numpy as np
import pytest

from sad_b200 import c2, retinanet_heads


@pytest.fixture(scope="module")
def oplib():
    return c2.OperatorLibrary()


def test_operators_registered_under_reference_names(oplib):
    # reference: pow_sum_op.cc:22-24, pow_sum_op.cu:45-46, ...loss_op.cc:21-24, ...loss_op.cu:174-177
    for dev in (c2.CPU, c2.CUDA):
        for name in ("PowSum", "SigmoidAdaptiveDistillLoss", "SigmoidAdaptiveDistillLossGradient"):
            assert oplib.HasOperator(name, dev), (name, dev)
    assert oplib.HasOperator("SigmoidAdaptiveDistillLossMultiLevel", c2.CUDA)
    assert oplib.HasOperator("SigmoidAdaptiveDistillStep", c2.CUDA)


def test_schema_arity_matches_reference(oplib):
    assert oplib.SchemaArity("PowSum")[0] == 1 and oplib.SchemaArity("PowSum")[2:] == (1, 1)
    assert oplib.SchemaArity("SigmoidAdaptiveDistillLoss") == (4, 4, 1, 1)
    assert oplib.SchemaArity("SigmoidAdaptiveDistillLossGradient") == (5, 5, 1, 1)


def test_schema_violation_is_an_enforce_failure(oplib):
    ws = oplib.Workspace()
    ws.FeedBlob("a", np.zeros((1, 4, 2, 2), np.float32))
    op = c2.CreateOperator("SigmoidAdaptiveDistillLoss", ["a", "a", "a"], ["l"])  # 3 inputs, needs 4
    with pytest.raises(c2.EnforceNotMet, match="schema"):
        ws.RunOperatorOnce(op)


def test_missing_input_blob_names_the_blob(oplib):
    ws = oplib.Workspace()
    op = c2.CreateOperator("PowSum", ["nope"], ["s"])
    with pytest.raises(c2.EnforceNotMet, match="nope"):
        ws.RunOperatorOnce(op)


def test_cpu_operators_are_not_implemented_like_the_reference(oplib):
    # pow_sum_op.h:33-36, ...loss_op.h:42-45: REGISTER_CPU_OPERATOR exists, RunOnDevice throws
    ws = oplib.Workspace()
    ws.FeedBlob("x", np.ones((1, 4, 2, 2), np.float32))
    with pytest.raises(c2.EnforceNotMet, match="Not Implemented"):
        ws.RunOperatorOnce(c2.CreateOperator("PowSum", ["x"], ["s"], power=1.8))
    ws.FeedBlob("g", np.zeros((1, 1, 2, 2), np.int32))
    ws.FeedBlob("wp", np.ones((1,), np.float32))
    with pytest.raises(c2.EnforceNotMet, match="Not Implemented"):
        ws.RunOperatorOnce(c2.CreateOperator("SigmoidAdaptiveDistillLoss", ["x", "x", "g", "wp"], ["l"], num_classes=4))


def test_argument_fields_are_type_strict(oplib):
    # proto_utils.cc:263-267: a float argument stored in `i` fails
    ws = oplib.Workspace()
    ws.FeedBlob("x", np.ones((4,), np.float32))
    with pytest.raises(c2.EnforceNotMet, match="power"):
        ws.RunOperatorOnce('input: "x" output: "s" type: "PowSum" arg { name: "power" i: 2 }')


def test_negative_scale_rejected_at_construction(oplib):
    ws = oplib.Workspace()
    ws.FeedBlob("x", np.ones((1, 4, 2, 2), np.float32))
    ws.FeedBlob("g", np.zeros((1, 1, 2, 2), np.int32))
    ws.FeedBlob("wp", np.ones((1,), np.float32))
    with pytest.raises(c2.EnforceNotMet, match="scale"):
        ws.RunOperatorOnce(c2.CreateOperator("SigmoidAdaptiveDistillLoss", ["x", "x", "g", "wp"], ["l"], scale=-1.0))


def test_gradient_maker_matches_reference(oplib):
    # ...loss_op.cc:99-110: {I0, I1, I2, I3, GO0} -> {GI0}, arguments copied from the forward def
    dev = c2.DeviceOption(c2.CUDA, 3)
    op = c2.CreateOperator("SigmoidAdaptiveDistillLoss", ["X", "T", "G", "wp"], ["loss"], device_option=dev,
                           gamma=2.0, alpha=0.5, scale=0.125, beta=0.0, num_classes=80, ignored_label=-1)
    text = oplib.GetGradientDefs(op, ["loss_grad"])
    assert 'type: "SigmoidAdaptiveDistillLossGradient"' in text
    ins = [l.split('"')[1] for l in text.splitlines() if l.strip().startswith("input:")]
    outs = [l.split('"')[1] for l in text.splitlines() if l.strip().startswith("output:")]
    assert ins == ["X", "T", "G", "wp", "loss_grad"] and outs == ["X_grad"]
    assert "cuda_gpu_id: 3" in text and "is_gradient_op: true" in text
    for frag in ('name: "gamma"', "f: 2", 'name: "scale"', "f: 0.125", 'name: "num_classes"', "i: 80"):
        assert frag in text
    # only the student logits get a gradient (teacher probs, labels, normaliser do not)
    assert [l.split('"')[1] for l in text.splitlines() if l.startswith("external_output")] == ["X_grad", "", "", ""]
    with pytest.raises(c2.EnforceNotMet, match="PowSum"):
        oplib.GetGradientDefs(c2.CreateOperator("PowSum", ["a"], ["s"]), ["s_grad"])  # no gradient registered


def test_netdef_text_round_trip(oplib):
    net, losses, grads = retinanet_heads.add_distill_loss(gpu_id=2, num_gpus=8)
    text = net.to_text()
    norm = oplib.NormalizeNetText(text)
    assert oplib.NormalizeNetText(norm) == norm
    assert norm.count("op {") == 1 + 5 + 5 + 5
    assert 'f: 0.125' in norm  # scale = T^2 / NUM_GPUS (retinanet_heads.py:342)
    assert 'f: 1.79999995' in norm or 'f: 1.8' in norm
    with pytest.raises(c2.EnforceNotMet):
        oplib.NormalizeNetText("op { type: \"X\" ")


def test_graph_wiring_matches_reference_names():
    net, losses, grads = retinanet_heads.add_distill_loss(gpu_id=0, num_gpus=1)
    ps = net.op[0]
    assert ps.type == "PowSum" and ps.output == ["gpu_0/distill_normalizer"]
    assert ps.input == ["gpu_0/teacher/retnet_cls_prob_fpn%d" % l for l in range(3, 8)]
    assert abs(ps.arg["power"] - 1.8) < 1e-12
    l3 = net.op[1]
    assert l3.type == "SigmoidAdaptiveDistillLoss"
    assert l3.input == ["gpu_0/retnet_cls_pred_fpn3", "gpu_0/teacher/retnet_cls_prob_fpn3",
                        "gpu_0/retnet_cls_labels_fpn3", "gpu_0/distill_normalizer"]
    assert l3.arg == dict(gamma=2.0, alpha=0.5, scale=1.0, beta=0.0, num_classes=80, ignored_label=-1)
    assert losses == ["gpu_0/fl_distill_fpn%d" % l for l in range(3, 8)]
    assert grads == ["gpu_0/retnet_cls_pred_fpn%d_grad" % l for l in range(3, 8)]
    # without the adaptive normaliser the foreground count is the normaliser (retinanet_heads.py:337)
    net2, _, _ = retinanet_heads.add_distill_loss(cfg=dict(ADAPTIVE_NORMALIZER=False))
    assert net2.op[0].type == "SigmoidAdaptiveDistillLoss" and net2.op[0].input[3] == "gpu_0/retnet_fg_num"


def test_fusion_pass_groups_levels(oplib):
    net, losses, grads = retinanet_heads.add_distill_loss(gpu_id=1, num_gpus=8)
    fused, n = oplib.FuseAdaptiveDistillOps(net.to_text())
    assert n == 1
    # the adaptive normaliser is a PowSum over exactly the group's teacher blobs (retinanet_heads.py:320-328): it is
    # folded in too -> ONE op (one cooperative launch) for PowSum + 5 losses + 5 gradients
    assert fused.count('type: "SigmoidAdaptiveDistillStep"') == 1
    assert fused.count('type: "SigmoidAdaptiveDistillLoss"') == 0
    assert fused.count('type: "SigmoidAdaptiveDistillLossGradient"') == 0
    assert fused.count('type: "PowSum"') == 0 and fused.count('type: "ConstantFill"') == 5
    for b in losses + grads + ["gpu_1/distill_normalizer"]:  # blob names preserved, the normaliser is still produced
        assert 'output: "%s"' % b in fused
    assert 'name: "power"' in fused and ('f: 1.8' in fused or 'f: 1.79999995' in fused)
    assert fused.count("input:") == 15 + 5          # 3 per level + the ConstantFills' inputs
    # different arguments -> not fusable
    net.op[2].arg["alpha"] = 0.25
    fused2, n2 = oplib.FuseAdaptiveDistillOps(net.to_text())
    assert fused2.count('type: "SigmoidAdaptiveDistillLoss"') >= 1
    # forward-only nets are left alone
    net3, _, _ = retinanet_heads.add_distill_loss(with_gradients=False)
    fused3, n3 = oplib.FuseAdaptiveDistillOps(net3.to_text())
    assert n3 == 0 and fused3.count('type: "SigmoidAdaptiveDistillLoss"') == 5


def test_fusion_pass_keeps_a_foreign_normaliser(oplib):
    # ADAPTIVE_NORMALIZER off: the normaliser is retnet_fg_num, produced elsewhere -> the multi-level op with a normaliser input
    net, losses, grads = retinanet_heads.add_distill_loss(cfg=dict(ADAPTIVE_NORMALIZER=False))
    fused, n = oplib.FuseAdaptiveDistillOps(net.to_text())
    assert n == 1 and fused.count('type: "SigmoidAdaptiveDistillLossMultiLevel"') == 1
    assert 'input: "gpu_0/retnet_fg_num"' in fused
    # a PowSum over OTHER blobs (or in another order) is not folded
    net, _, _ = retinanet_heads.add_distill_loss()
    net.op[0].input = list(reversed(net.op[0].input))
    fused, n = oplib.FuseAdaptiveDistillOps(net.to_text())
    assert n == 1 and fused.count('type: "PowSum"') == 1 and fused.count('type: "SigmoidAdaptiveDistillLossMultiLevel"') == 1
    assert fused.index("PowSum") < fused.index("MultiLevel")


# ---------------------------------------------------------------------------------------------
# head convolution operators (Conv / ConvGradient with and without engine CUDNN, Relu / ReluGradient)
# ---------------------------------------------------------------------------------------------
def test_head_conv_operators_registered(oplib):
    # conv_op_cudnn.cc:1130-1131 registers the CUDNN engine; detector.py:56-60 makes Detectron ask for it
    for name in ("Conv", "ConvGradient", "Relu", "ReluGradient"):
        assert oplib.HasOperator(name, c2.CUDA), name
    assert oplib.SchemaArity("Conv") == (2, 3, 1, 1)
    assert oplib.SchemaArity("ConvGradient") == (2, 3, 1, 3)
    assert oplib.SchemaArity("Relu") == (1, 1, 1, 1)
    assert oplib.SchemaArity("ReluGradient") == (2, 2, 1, 1)


def _grad_fields(text):
    line = lambda key: [l.split('"')[1] for l in text.splitlines() if l.strip().startswith(key + ":")]
    return line("type"), line("input"), line("output"), [l.split('"')[1] for l in text.splitlines() if l.startswith("external_output")]


def test_conv_gradient_maker_follows_reference(oplib):
    # conv_gradient_op.cc:35-77: inputs {X, W, dY}; outputs {dW, db, dX}; engine, device and arguments copied
    dev = c2.DeviceOption(c2.CUDA, 0)
    op = c2.CreateOperator("Conv", ["x", "w", "b"], ["y"], device_option=dev, engine="CUDNN", kernel=3, pad=1, stride=1, order="NCHW")
    text = oplib.GetGradientDefs(op, ["y_grad"])
    types, ins, outs, gin = _grad_fields(text)
    assert types == ["ConvGradient"] and 'engine: "CUDNN"' in text and "is_gradient_op: true" in text
    assert ins == ["x", "w", "y_grad"] and outs == ["w_grad", "b_grad", "x_grad"]
    for frag in ('name: "kernel"', "i: 3", 'name: "pad"', 'name: "order"', 's: "NCHW"'):
        assert frag in text
    assert gin == ["x_grad", "w_grad", "b_grad"]
    # without bias: no_bias = 1 is appended and only {dW, dX} are produced
    op = c2.CreateOperator("Conv", ["x", "w"], ["y"], device_option=dev, kernel=3, pad=1, stride=1)
    text = oplib.GetGradientDefs(op, ["y_grad"])
    assert _grad_fields(text)[2] == ["w_grad", "x_grad"] and 'name: "no_bias"' in text
    # no_gradient_to_input (the first conv of a frozen body) drops dX
    op = c2.CreateOperator("Conv", ["x", "w", "b"], ["y"], device_option=dev, kernel=3, pad=1, stride=1, no_gradient_to_input=1)
    assert _grad_fields(oplib.GetGradientDefs(op, ["y_grad"]))[2] == ["w_grad", "b_grad"]


def test_relu_gradient_maker_uses_the_output(oplib):
    # relu_op.cc:98-108: ReluGradient(Y, dY) -> dX, which is what makes the in-place Relu of the towers legal
    op = c2.CreateOperator("Relu", ["t"], ["t"], device_option=c2.DeviceOption(c2.CUDA, 0))
    types, ins, outs, gin = _grad_fields(oplib.GetGradientDefs(op, ["t_grad"]))
    assert types == ["ReluGradient"] and ins == ["t", "t_grad"] and outs == ["t_grad"]


# ---------------------------------------------------------------------------------------------
# head graph (retinanet_heads.py:63-245) and its losses (:248-311) as NetDefs
# ---------------------------------------------------------------------------------------------
def test_head_graph_wiring_matches_reference(oplib):
    from sad_b200 import head
    blobs_in = ["gpu_0/fpn_%d" % l for l in (7, 6, 5, 4, 3)]                 # coarsest first, like FPN.add_fpn returns them
    net, params, cls_out, box_out = retinanet_heads.add_fpn_retinanet_outputs(blobs_in, gpu_id=0, train=True)
    # per level: 2 towers x 4 x (Conv + in-place Relu) + 2 prediction convolutions
    assert len(net.op) == 5 * (2 * 4 * 2 + 2)
    assert [o.type for o in net.op[:9]] == ["Conv", "Relu"] * 4 + ["Conv"]
    first = net.op[0]
    assert first.input == ["gpu_0/fpn_3", "gpu_0/retnet_cls_conv_n0_fpn3_w", "gpu_0/retnet_cls_conv_n0_fpn3_b"]
    assert first.output == ["gpu_0/retnet_cls_conv_n0_fpn3"] and first.engine == "CUDNN"
    assert first.arg == dict(kernel=3, pad=1, stride=1, order="NCHW")
    assert net.op[1].input == net.op[1].output == first.output               # model.Relu(bl_out, bl_out)
    # levels 4..7 are ConvShared: their convolutions read level 3's parameter blobs
    lvl5 = [o for o in net.op if o.type == "Conv" and o.output[0].endswith("_fpn5")]
    assert len(lvl5) == 10 and all(o.input[1].endswith("_fpn3_w") and o.input[2].endswith("_fpn3_b") for o in lvl5)
    assert [o.input[0] for o in net.op if o.output == ["gpu_0/retnet_cls_conv_n0_fpn7"] and o.type == "Conv"] == ["gpu_0/fpn_7"]
    assert cls_out == ["gpu_0/retnet_cls_pred_fpn%d" % l for l in range(3, 8)]
    assert box_out == ["gpu_0/retnet_bbox_pred_fpn%d" % l for l in range(3, 8)]
    # parameters: created once (level 3), same names as the fused head object's flat buffer
    names = [n for n, _, _ in params]
    assert len(names) == 20 and sorted(names) == sorted("gpu_0/" + n for n in head.param_names())
    shapes = {n: s for n, s, _ in params}
    assert shapes["gpu_0/retnet_cls_pred_fpn3_w"] == (720, 256, 3, 3) and shapes["gpu_0/retnet_bbox_pred_fpn3_w"] == (36, 256, 3, 3)
    assert shapes["gpu_0/retnet_bbox_conv_n3_fpn3_b"] == (256,)
    fills = {n: f for n, _, f in params}
    assert fills["gpu_0/retnet_cls_conv_n0_fpn3_w"] == ("GaussianFill", {"std": 0.01})
    kind, args = fills["gpu_0/retnet_cls_pred_fpn3_b"]
    assert kind == "ConstantFill" and abs(args["value"] - (-4.59511985013459)) < 1e-12       # -log((1 - pi) / pi), pi = 0.01
    assert sum(int(np.prod(s)) for s in shapes.values()) == 6463220
    # every op resolves in the registry and the text survives the parser
    assert all(oplib.HasOperator(o.type, c2.CUDA) for o in net.op)
    assert oplib.NormalizeNetText(net.to_text()).count("op {") == len(net.op)
    # the teacher's graph (model.train = False, name scope teacher/): Sigmoid -> retnet_cls_prob_fpnL
    tnet, tparams, tcls, _ = retinanet_heads.add_fpn_retinanet_outputs(["gpu_1/teacher/fpn_%d" % l for l in (7, 6, 5, 4, 3)],
                                                                       gpu_id=1, train=False, scope="teacher/")
    assert len(tnet.op) == len(net.op) + 5 and oplib.HasOperator("Sigmoid", c2.CUDA)
    sig = [o for o in tnet.op if o.type == "Sigmoid"]
    assert [o.input[0] for o in sig] == ["gpu_1/teacher/retnet_cls_pred_fpn%d" % l for l in range(3, 8)]
    assert tcls == ["gpu_1/teacher/retnet_cls_prob_fpn%d" % l for l in range(3, 8)]
    # ... which are exactly PowSum's inputs in add_distill_loss
    assert retinanet_heads.add_distill_loss(gpu_id=1, num_gpus=8)[0].op[0].input == tcls
    # a shared tower feeds the box predictions from the class tower
    snet, sparams, _, _ = retinanet_heads.add_fpn_retinanet_outputs(blobs_in, cfg=dict(SHARE_CLS_BBOX_TOWER=True))
    assert len(snet.op) == 5 * (4 * 2 + 2) and len(sparams) == 12
    assert [o.input[0] for o in snet.op if o.output == ["gpu_0/retnet_bbox_pred_fpn4"]] == ["gpu_0/retnet_cls_conv_n3_fpn4"]


def test_head_loss_wiring_matches_reference(oplib):
    net, losses = retinanet_heads.add_fpn_retinanet_losses(gpu_id=2, num_gpus=8)
    assert [o.type for o in net.op] == ["SelectSmoothL1Loss"] * 5 + ["SigmoidFocalLoss"] * 5
    box = net.op[0]
    assert box.input == ["gpu_2/retnet_bbox_pred_fpn3", "gpu_2/retnet_roi_bbox_targets_fpn3", "gpu_2/retnet_roi_fg_bbox_locs_fpn3",
                         "gpu_2/retnet_fg_num"]
    assert box.arg == dict(beta=0.11, scale=0.125)
    fl = net.op[5]
    assert fl.input == ["gpu_2/retnet_cls_pred_fpn3", "gpu_2/retnet_cls_labels_fpn3", "gpu_2/retnet_fg_num"]
    assert fl.arg == dict(gamma=2.0, alpha=0.25, scale=0.125, num_classes=80)
    assert losses == ["gpu_2/retnet_loss_bbox_fpn%d" % l for l in range(3, 8)] + ["gpu_2/fl_fpn%d" % l for l in range(3, 8)]
    assert all(oplib.HasOperator(o.type, c2.CUDA) for o in net.op)
