"""Host mirror of the graph wiring on the path: detectron/lib/modeling/retinanet_heads.py:63-245
(`add_fpn_retinanet_outputs`, the head), :248-311 (`add_fpn_retinanet_losses`) and :313-352
(`add_distill_loss`) plus what Caffe2 autograd appends for the latter (loss-gradient ConstantFill from
detectron/lib/utils/blob.py:166-172 and the gradient ops from the C++ gradient maker).

The reference emits OperatorDefs by name into the NetDef; this module emits the same ops with the
same blob names and argument values, so the resulting NetDef text is what the drop-in library
sees when it is loaded into the reference's training graph.
"""
from . import c2

# config defaults: detectron/lib/core/config.py:989-1016 and
# configs/focal_distillation/retinanet_R-50-FPN_distillation.yaml:45-54
DISTILLATION = dict(LOSS_ALPHA=0.5, LOSS_GAMMA=2.0, LOSS_BETA=0.0, IGNORED_LABEL=-1, TEMPERATURE=1.0,
                    ADAPTIVE_NORMALIZER=True, LOGITS_POWER=1.8)
RPN_MIN_LEVEL, RPN_MAX_LEVEL = 3, 7
NUM_CLASSES = 81  # cfg.MODEL.NUM_CLASSES incl. background; the op gets NUM_CLASSES - 1
# detectron/lib/core/config.py RETINANET.*: 3 aspect ratios x 3 scales per octave, 4 tower convolutions, separate towers,
# per-class sigmoid, prior 0.01, focal gamma 2 / alpha 0.25, box beta 0.11 / weight 1
RETINANET = dict(NUM_CONVS=4, ASPECT_RATIOS=3, SCALES_PER_OCTAVE=3, SHARE_CLS_BBOX_TOWER=False, PRIOR_PROB=0.01,
                 LOSS_GAMMA=2.0, LOSS_ALPHA=0.25, BBOX_REG_BETA=0.11, BBOX_REG_WEIGHT=1.0, CLASS_SPECIFIC_BBOX=False)
FPN_DIM = 256


def get_retinanet_bias_init(cfg=None):
    """retinanet_heads.py:29-60, per-class sigmoid case: ('ConstantFill', value = -log((1 - pi) / pi))."""
    import math
    pi = dict(RETINANET, **(cfg or {}))["PRIOR_PROB"]
    return ("ConstantFill", {"value": -math.log((1.0 - pi) / pi)})


def add_fpn_retinanet_outputs(blobs_in, gpu_id=0, train=True, dim_in=FPN_DIM, cfg=None, scope="",
                              k_min=RPN_MIN_LEVEL, k_max=RPN_MAX_LEVEL, enable_tensor_core=None):
    """The RetinaNet head as the reference emits it (retinanet_heads.py:63-245): per level two towers of NUM_CONVS
    Conv3x3 + in-place Relu, then the class-logit and box convolutions; level k_min creates the parameters, the other levels
    reuse them (ConvShared = a Conv reading level k_min's `_w` / `_b` blobs); `train=False` (the teacher,
    model_builder.py:379-393) appends Sigmoid -> retnet_cls_prob_fpnL.  `blobs_in` is in the reference's reversed order
    (coarsest level first), `scope` is e.g. "teacher/".

    Returns (NetDef, params, cls_outputs, bbox_outputs): params = [(blob, shape, (fill_op, fill_args))] in creation order —
    the same blob names as head.param_names(), which lays them out weights first, biases second."""
    cfg = dict(RETINANET, **(cfg or {}))
    assert len(blobs_in) == k_max - k_min + 1
    A = cfg["ASPECT_RATIOS"] * cfg["SCALES_PER_OCTAVE"]
    cls_dim = (NUM_CLASSES - 1) * A
    box_dim = (4 * (NUM_CLASSES - 1) if cfg["CLASS_SPECIFIC_BBOX"] else 4) * A
    pre = "gpu_%d/%s" % (gpu_id, scope)
    dev = c2.DeviceOption(c2.CUDA, gpu_id)
    gauss, zero = ("GaussianFill", {"std": 0.01}), ("ConstantFill", {"value": 0.0})
    ops, params = [], []

    def conv(bl_in, name, cout, lvl, bias_init=zero):
        # model.Conv / model.ConvShared (cnn.py): Conv with engine CUDNN, order NCHW; the parameter blobs belong to level k_min
        owner = name.replace("fpn%d" % lvl, "fpn%d" % k_min)
        if lvl == k_min:
            params.append((pre + name + "_w", (cout, dim_in, 3, 3), gauss))
            params.append((pre + name + "_b", (cout,), bias_init))
        # enable_tensor_core: the reference's own Conv argument (conv_op_cudnn.cc:77-86), absent from the graphs Detectron emits
        # (= 0: fp32 arithmetic; this library then computes in 3xTF32).  1 selects single-pass tf32.
        extra = {} if enable_tensor_core is None else {"enable_tensor_core": int(enable_tensor_core)}
        ops.append(c2.CreateOperator("Conv", [bl_in, pre + owner + "_w", pre + owner + "_b"], [pre + name], device_option=dev,
                                     engine="CUDNN", kernel=3, pad=1, stride=1, order="NCHW", **extra))
        return pre + name

    def tower(kind, lvl):
        bl = blobs_in[k_max - lvl]
        for n in range(cfg["NUM_CONVS"]):
            bl = conv(bl, "retnet_%s_conv_n%d_fpn%d" % (kind, n, lvl), dim_in, lvl)
            ops.append(c2.CreateOperator("Relu", [bl], [bl], device_option=dev))      # model.Relu(bl_out, bl_out)
        return bl

    cls_out, bbox_feat = [], []
    for lvl in range(k_min, k_max + 1):
        feat = tower("cls", lvl)
        pred = conv(feat, "retnet_cls_pred_fpn%d" % lvl, cls_dim, lvl, bias_init=get_retinanet_bias_init(cfg))
        if not train:
            prob = pre + "retnet_cls_prob_fpn%d" % lvl
            ops.append(c2.CreateOperator("Sigmoid", [pred], [prob], device_option=dev))
            pred = prob
        cls_out.append(pred)
        if cfg["SHARE_CLS_BBOX_TOWER"]:
            bbox_feat.append(feat)
    if not cfg["SHARE_CLS_BBOX_TOWER"]:
        for lvl in range(k_min, k_max + 1):
            bbox_feat.append(tower("bbox", lvl))
    box_out = [conv(bbox_feat[i], "retnet_bbox_pred_fpn%d" % lvl, box_dim, lvl) for i, lvl in enumerate(range(k_min, k_max + 1))]
    return c2.NetDef("retinanet_head_gpu%d%s" % (gpu_id, "_" + scope.strip("/") if scope else ""), ops), params, cls_out, box_out


def add_fpn_retinanet_losses(gpu_id=0, num_gpus=1, cfg=None, k_min=RPN_MIN_LEVEL, k_max=RPN_MAX_LEVEL):
    """retinanet_heads.py:248-311: SelectSmoothL1Loss per level, then SigmoidFocalLoss per level, loss scale 1 / NUM_GPUS
    (model.GetLossScale, detector.py:650-655).  Returns (NetDef, loss blob names)."""
    cfg = dict(RETINANET, **(cfg or {}))
    pre = "gpu_%d/" % gpu_id
    dev = c2.DeviceOption(c2.CUDA, gpu_id)
    scale = 1.0 / num_gpus
    ops, losses = [], []
    for lvl in range(k_min, k_max + 1):
        sfx = "fpn%d" % lvl
        ops.append(c2.CreateOperator(
            "SelectSmoothL1Loss",
            [pre + "retnet_bbox_pred_" + sfx, pre + "retnet_roi_bbox_targets_" + sfx, pre + "retnet_roi_fg_bbox_locs_" + sfx,
             pre + "retnet_fg_num"],
            [pre + "retnet_loss_bbox_" + sfx], device_option=dev, beta=float(cfg["BBOX_REG_BETA"]),
            scale=scale * float(cfg["BBOX_REG_WEIGHT"])))
        losses.append(pre + "retnet_loss_bbox_" + sfx)
    for lvl in range(k_min, k_max + 1):
        sfx = "fpn%d" % lvl
        ops.append(c2.CreateOperator(
            "SigmoidFocalLoss", [pre + "retnet_cls_pred_" + sfx, pre + "retnet_cls_labels_" + sfx, pre + "retnet_fg_num"],
            [pre + "fl_" + sfx], device_option=dev, gamma=float(cfg["LOSS_GAMMA"]), alpha=float(cfg["LOSS_ALPHA"]), scale=scale,
            num_classes=NUM_CLASSES - 1))
        losses.append(pre + "fl_" + sfx)
    return c2.NetDef("retinanet_losses_gpu%d" % gpu_id, ops), losses


def add_distill_loss(gpu_id=0, num_gpus=1, cfg=None, with_gradients=True, k_min=RPN_MIN_LEVEL, k_max=RPN_MAX_LEVEL):
    """Returns (NetDef, loss_blob_names, gradient_blob_names) for one GPU's name scope."""
    cfg = dict(DISTILLATION, **(cfg or {}))
    scope = "gpu_%d/" % gpu_id
    dev = c2.DeviceOption(c2.CUDA, gpu_id)
    ops, losses, grads = [], [], []
    levels = list(range(k_min, k_max + 1))
    if cfg["ADAPTIVE_NORMALIZER"]:
        # retinanet_heads.py:320-328: normaliser = sum over levels of teacher_prob ** LOGITS_POWER
        normalizer = scope + "distill_normalizer"
        ops.append(c2.CreateOperator("PowSum", [scope + "teacher/retnet_cls_prob_fpn%d" % l for l in levels],
                                     [normalizer], device_option=dev, power=float(cfg["LOGITS_POWER"])))
    else:
        normalizer = scope + "retnet_fg_num"
    for l in levels:
        suffix = "fpn%d" % l
        loss = scope + "fl_distill_" + suffix
        # retinanet_heads.py:331-345
        ops.append(c2.CreateOperator(
            "SigmoidAdaptiveDistillLoss",
            [scope + "retnet_cls_pred_" + suffix, scope + "teacher/retnet_cls_prob_" + suffix,
             scope + "retnet_cls_labels_" + suffix, normalizer],
            [loss], device_option=dev,
            gamma=float(cfg["LOSS_GAMMA"]), alpha=float(cfg["LOSS_ALPHA"]),
            scale=float(cfg["TEMPERATURE"]) ** 2 / num_gpus, beta=float(cfg["LOSS_BETA"]),
            num_classes=NUM_CLASSES - 1, ignored_label=int(cfg["IGNORED_LABEL"])))
        losses.append(loss)
    if with_gradients:
        fwd = [op for op in ops if op.type == "SigmoidAdaptiveDistillLoss"]
        for op in fwd:  # blob.py:166-172: loss gradient seeded with ones
            ops.append(c2.CreateOperator("ConstantFill", [op.output[0]], [op.output[0] + "_grad"], device_option=dev, value=1.0))
        for op in reversed(fwd):  # autograd walks the net backwards
            g = c2.CreateOperator("SigmoidAdaptiveDistillLossGradient",
                                  op.input + [op.output[0] + "_grad"], [op.input[0] + "_grad"],
                                  device_option=dev, **op.arg)
            g.is_gradient_op = True
            ops.append(g)
            grads.append(op.input[0] + "_grad")
        grads.reverse()
    return c2.NetDef("distill_loss_gpu%d" % gpu_id, ops), losses, grads


def feed_level_blobs(ws, gpu_id, levels, k_min=RPN_MIN_LEVEL):
    """Feed (logits, teacher_prob, labels) device tensors under the reference's blob names."""
    scope = "gpu_%d/" % gpu_id
    for i, (x, t, g) in enumerate(levels):
        suffix = "fpn%d" % (k_min + i)
        ws.FeedBlob(scope + "retnet_cls_pred_" + suffix, x)
        ws.FeedBlob(scope + "teacher/retnet_cls_prob_" + suffix, t)
        ws.FeedBlob(scope + "retnet_cls_labels_" + suffix, g)
