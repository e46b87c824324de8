"""Host mirror of the graph wiring on the path: detectron/lib/modeling/retinanet_heads.py:313-352
(`add_distill_loss`) plus what Caffe2 autograd appends for it (loss-gradient ConstantFill from
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
