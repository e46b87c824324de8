"""The full RetinaNet distillation training step of BASELINE.json configs[2..4] on one GPU's image shard:

    teacher  (frozen, forward only): ResNet-101 + FPN body -> RetinaNet head -> Sigmoid           model_builder.py:379-393
    student  (trained)             : ResNet-50  + FPN body -> RetinaNet head                      model_builder.py:395-400
    losses   : SigmoidFocalLoss + SelectSmoothL1Loss (retinanet_heads.py:248-311) and the adaptive distillation loss
               PowSum + SigmoidAdaptiveDistillLoss (retinanet_heads.py:313-352)                   model_builder.py:402-406
    backward : head (ConvGradient / ReluGradient / Sum) -> FPN -> ResNet body (res2 and below frozen, ResNet.py:88-104)
    exchange : ONE allreduce over the flat gradient buffer [head | body]                          optimizer.py:72-92
    update   : momentum SGD, bias 2x / weight decay preamble, ONE launch over the flat buffers    optimizer.py:95-130

What runs where.  On this repository's kernels: both RetinaNet heads (tcgen05 convolutions,
`sad_head_*`), PowSum + distillation loss + gradient (one cooperative launch), SigmoidFocalLoss + gradient (accumulated into
the same d(logits)), SelectSmoothL1Loss + gradient, and the gradient exchange — every loss and the whole head, forward and
backward.  SCAFFOLDING in plain PyTorch / cuDNN (§8f ranks 2-4), there only so that the step is complete and its imgs/s can
be measured: the ResNet + FPN bodies (random init, AffineChannel = frozen per-channel scale/bias,
affine_channel_op.cc:70-78; autograd carries d(fpn_L) from the head's backward into them).  The teacher's Sigmoid is fused
into its prediction convolution (§8f rank 2) and the optimiser step is one native launch over the flat buffers (§8f rank 4).  Synthetic images, labels, foreground locations and box targets (there is no dataset in this environment).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, parallel, synthetic
from .head import RetinaNetHead, head_param_count


# ------------------------------------------------------------------------------------------------------------
# scaffolding: ResNet-C4..C5 + FPN in PyTorch (detectron/lib/modeling/ResNet.py:88-278, FPN.py:94-249)
# ------------------------------------------------------------------------------------------------------------
class AffineChannel(nn.Module):
    """Frozen batch-norm stand-in: y = x * scale + bias per channel, parameters not trained."""

    def __init__(self, c):
        super().__init__()
        self.register_buffer("scale", torch.ones(1, c, 1, 1))
        self.register_buffer("bias", torch.zeros(1, c, 1, 1))

    def forward(self, x):
        return torch.addcmul(self.bias, x, self.scale)


class _ConvAffine(torch.autograd.Function):
    """act(conv(x, w) * scale + bias (+ z)) as ONE cuDNN call: the frozen AffineChannel is folded into the weights
    (conv(x, w) * s = conv(x, w * s)) and bias, residual add and ReLU ride in the convolution's epilogue
    (cudnnConvolutionBiasActivationForward).  Unfused, every convolution of the bodies is followed by 2-3 elementwise
    passes over its output, which cost more than the convolutions themselves (measured: 4.7 of 13.5 ms at bs = 2).
    Backward: ReluGradient from the saved output (one pass), cuDNN data / weight gradients of the folded convolution,
    dW = dW_folded * s.  `prefolded`: w * s computed by the caller for every convolution at once (FullDistillStep: one
    multi-tensor launch per step instead of one per convolution); the UNSCALED weight gradient is then handed to the caller
    through `sink` (a list and a slot in it) instead of autograd, and the caller forms grad += dW_folded * s for every
    convolution in one multi-tensor launch after the backward pass."""

    @staticmethod
    def forward(ctx, x, w, scale, bias, z, stride, padding, relu, groups=1, prefolded=None, sink=None):
        w_eff = prefolded if prefolded is not None else (w if scale is None else w * scale.view(-1, 1, 1, 1))
        ctx.sink = sink if prefolded is not None else None
        if prefolded is not None:
            scale = None
        s, p, d = (stride, stride), (padding, padding), (1, 1)
        if z is not None:
            out = torch.cudnn_convolution_add_relu(x, w_eff, z, 1.0, bias, s, p, d, groups)
        elif relu:
            out = torch.cudnn_convolution_relu(x, w_eff, bias, s, p, d, groups)
        else:
            out = F.conv2d(x, w_eff, bias, stride, padding, 1, groups)
        ctx.conv = (s, p, d, groups)
        ctx.masked = relu or z is not None
        ctx.has_z = z is not None
        ctx.save_for_backward(x, w_eff, scale, out if ctx.masked else None)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, w_eff, scale, out = ctx.saved_tensors
        s, p, d, groups = ctx.conv
        g = torch.ops.aten.threshold_backward(dy, out, 0) if ctx.masked else dy
        dx, dw, _ = torch.ops.aten.convolution_backward(g, x, w_eff, None, s, p, d, False, [0, 0], groups,
                                                        [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        if dw is not None and scale is not None:
            dw = dw * scale.view(-1, 1, 1, 1)
        if dw is not None and ctx.sink is not None:
            ctx.sink[0][ctx.sink[1]] = dw
            dw = None
        dz = g if (ctx.has_z and ctx.needs_input_grad[4]) else None
        return dx, dw, None, None, dz, None, None, None, None, None, None


def conv_affine(conv, aff, x, z=None, relu=True):
    """Fused conv -> AffineChannel (-> + z) (-> ReLU).  Frozen convolutions (the teacher, res2 and below) fold once."""
    stride, padding = conv.stride[0], conv.padding[0]
    if not conv.weight.requires_grad:
        folded = getattr(conv, "_folded", None)
        if folded is None:
            with torch.no_grad():
                folded = (conv.weight * aff.scale.view(-1, 1, 1, 1)).contiguous(memory_format=torch.channels_last)
            conv._folded = folded
        return _ConvAffine.apply(x, folded, None, aff.bias.view(-1), z, stride, padding, relu, conv.groups)
    pre = getattr(conv, "_prefolded", None)       # (w * s, (gradient list, slot)) set per step by FullDistillStep
    return _ConvAffine.apply(x, conv.weight, aff.scale, aff.bias.view(-1), z, stride, padding, relu, conv.groups,
                             pre[0] if pre else None, pre[1] if pre else None)


class Bottleneck(nn.Module):
    def __init__(self, cin, cout, cmid, stride, groups=1, stride_1x1=True):
        super().__init__()
        # stride on the first 1x1 (RESNETS.STRIDE_1X1 = True for the R-50 / R-101 configs) or on the 3x3 (False: the ResNeXt
        # teacher, configs/focal_distillation/retinanet_X-101-64x4d-FPN_1x_teacher.yaml:20-24; ResNet.py:236-278)
        s1, s3 = (stride, 1) if stride_1x1 else (1, stride)
        self.c1, self.a1 = nn.Conv2d(cin, cmid, 1, stride=s1, bias=False), AffineChannel(cmid)
        self.c2, self.a2 = nn.Conv2d(cmid, cmid, 3, stride=s3, padding=1, groups=groups, bias=False), AffineChannel(cmid)
        self.c3, self.a3 = nn.Conv2d(cmid, cout, 1, bias=False), AffineChannel(cout)
        self.short = None
        self.fused = True
        if cin != cout or stride != 1:
            self.short = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), AffineChannel(cout))

    def forward(self, x):
        if self.fused:
            y = conv_affine(self.c1, self.a1, x)
            y = conv_affine(self.c2, self.a2, y)
            z = x if self.short is None else conv_affine(self.short[0], self.short[1], x, relu=False)
            return conv_affine(self.c3, self.a3, y, z=z)
        y = F.relu(self.a1(self.c1(x)), inplace=True)
        y = F.relu(self.a2(self.c2(y)), inplace=True)
        y = self.a3(self.c3(y))
        return F.relu(y + (x if self.short is None else self.short(x)), inplace=True)


class ResNetFPN(nn.Module):
    """C3..C5 -> P3..P7 (FPN.DIM = 256; P6, P7 by stride-2 3x3 convolutions: FPN.EXTRA_CONV_LEVELS, FPN.py:199-219)."""

    def __init__(self, blocks=(3, 4, 6, 3), dim=256, fused=True, groups=1, width_per_group=64, stride_1x1=True):
        """groups / width_per_group: RESNETS.NUM_GROUPS / WIDTH_PER_GROUP (1 / 64 = ResNet, 64 / 4 = ResNeXt-101-64x4d: the
        bottleneck width is groups * width_per_group * 2^stage, ResNet.py:94-97)."""
        super().__init__()
        self.fused = fused
        self.stem = nn.Sequential(nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), AffineChannel(64), nn.ReLU(inplace=True),
                                  nn.MaxPool2d(3, stride=2, padding=1))
        stages, cin = [], 64
        for i, n in enumerate(blocks):
            cout, cmid = 256 * 2 ** i, groups * width_per_group * 2 ** i
            stages.append(nn.Sequential(*[Bottleneck(cin if j == 0 else cout, cout, cmid, (1 if i == 0 else 2) if j == 0 else 1,
                                                     groups=groups, stride_1x1=stride_1x1) for j in range(n)]))
            cin = cout
        self.res2, self.res3, self.res4, self.res5 = stages
        self.lat = nn.ModuleList([nn.Conv2d(c, dim, 1) for c in (512, 1024, 2048)])
        self.out = nn.ModuleList([nn.Conv2d(dim, dim, 3, padding=1) for _ in range(3)])
        self.p6 = nn.Conv2d(2048, dim, 3, stride=2, padding=1)
        self.p7 = nn.Conv2d(dim, dim, 3, stride=2, padding=1)
        for p in list(self.stem.parameters()) + list(self.res2.parameters()):   # TRAIN.FREEZE_AT = 2
            p.requires_grad_(False)
        for m in self.modules():
            if isinstance(m, Bottleneck):
                m.fused = m.fused and fused

    def forward(self, x):
        with torch.no_grad():
            if self.fused:
                c2 = self.res2(self.stem[3](conv_affine(self.stem[0], self.stem[1], x)))
            else:
                c2 = self.res2(self.stem(x))
        c3 = self.res3(c2)
        c4 = self.res4(c3)
        c5 = self.res5(c4)
        # d(c4) is complete once res5 AND every FPN layer have run their backward; d(c3) once res4 has: the points at which
        # the gradient exchange of those parameters can start (FullDistillStep._bucket_ready)
        hooks = getattr(self, "grad_ready_hooks", None)
        if hooks and torch.is_grad_enabled():
            if c4.requires_grad and "c4" in hooks:
                c4.register_hook(lambda g, f=hooks["c4"]: f())
            if c3.requires_grad and "c3" in hooks:
                c3.register_hook(lambda g, f=hooks["c3"]: f())
        p5 = self.lat[2](c5)
        p4 = self.lat[1](c4) + F.interpolate(p5, scale_factor=2, mode="nearest")
        p3 = self.lat[0](c3) + F.interpolate(p4, scale_factor=2, mode="nearest")
        p6 = self.p6(c5)
        p7 = self.p7(F.relu(p6))
        return [self.out[0](p3), self.out[1](p4), self.out[2](p5), p6, p7]


class FullDistillStep:
    def __init__(self, n_images=2, scale_px=600, world=1, rank=0, seed=1234, student_blocks=(3, 4, 6, 3), teacher_blocks=(3, 4, 23, 3),
                 temperature=1.0, power=1.8, distill_alpha=0.5, distill_gamma=2.0, lr=0.01, momentum=0.9, weight_decay=1e-4,
                 fused_body=True, overlap_teacher=True, teacher_body=None, teacher_head_f16=False, student_head_f16=False,
                 overlap_exchange=True):
        """teacher_body: extra ResNetFPN arguments of the teacher, e.g. dict(groups=64, width_per_group=4, stride_1x1=False)
        for the ResNeXt-101-64x4d teacher of BASELINE.json configs[4].  teacher_head_f16: the forward-only teacher head on fp16
        operands (tcgen05 kind::f16, fp32 accumulation; configs[4]: "mixed fp16 compute / fp32 loss accumulate"); student_head_f16: the
        student head's forward AND backward on fp16 operands (loss-scaled fp16 gradient tensors, fp32 parameters and parameter gradients)."""
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.world, self.rank, self.images = int(world), int(rank), int(n_images)
        self.overlap_teacher = bool(overlap_teacher)
        # overlap_exchange: the gradient allreduce runs in buckets on the exchange's own stream while the backward pass is still
        # producing the next bucket (head bucket under the whole body backward, res5 + FPN under res4 + res3, ...), and the
        # optimiser waits only for the join (optimizer.py:72-92 issues its NCCLAllreduce ops after the whole backward)
        self.overlap_exchange = bool(overlap_exchange)
        self.student_head_f16 = bool(student_head_f16)
        self.overflow = torch.zeros((), dtype=torch.int32, device=self.device)   # set by the device when a reduced gradient is inf / NaN
        self.loss_scaler = None
        self.teacher_stream = torch.cuda.Stream(device=self.device)
        self.lr, self.mom, self.wd = lr, momentum, weight_decay
        shapes = synthetic.level_shapes(scale_px)
        H, W = shapes[0][0] * 8, shapes[0][1] * 8
        torch.backends.cudnn.allow_tf32 = True          # the bodies run on cuDNN's TF32 tensor-core path (scaffolding)
        torch.backends.cudnn.benchmark = True
        torch.manual_seed(seed)                         # same weights on every rank
        self.student = ResNetFPN(student_blocks, fused=fused_body).to(self.device).to(memory_format=torch.channels_last)
        self.teacher = ResNetFPN(teacher_blocks, fused=fused_body, **(teacher_body or {})).to(self.device).to(memory_format=torch.channels_last).eval()
        for p in self.teacher.parameters():
            p.requires_grad_(False)
        # ONE flat parameter buffer and ONE flat gradient buffer laid out [head weights | head biases | body weights | body biases]:
        # the head writes its gradient slice, autograd accumulates into views of the body slice; one allreduce, one SGD launch
        body = [p for p in self.student.parameters() if p.requires_grad]
        self.body_params = [p for p in body if p.dim() > 1] + [p for p in body if p.dim() == 1]
        n_body_w = sum(p.numel() for p in self.body_params if p.dim() > 1)
        n_body = sum(p.numel() for p in self.body_params)
        n_head = head_param_count()   # (no throw-away head: its arena is a cudaMalloc of hundreds of MB)
        self.flat_grads = torch.zeros(n_head + n_body, dtype=torch.float32, device=self.device)
        self.flat_params = torch.zeros(n_head + n_body, dtype=torch.float32, device=self.device)
        self.head = RetinaNetHead(n_images, shapes, device=self.device, seed=seed, grad_buffer=self.flat_grads[:n_head],
                                  param_buffer=self.flat_params[:n_head], compute_f16=student_head_f16)
        self.teacher_head = RetinaNetHead(n_images, shapes, device=self.device, seed=seed + 1, cls_output_sigmoid=True,
                                          compute_f16=teacher_head_f16)
        off = n_head
        for p in self.body_params:
            k = p.numel()
            if p.dim() == 4:   # stored channels-last inside the flat buffers: cuDNN takes the views as they are (no per-step copies)
                co, ci, kh, kw = p.shape
                view = lambda flat: flat[off:off + k].view(co, kh, kw, ci).permute(0, 3, 1, 2)
            else:
                view = lambda flat: flat[off:off + k].view(p.shape)
            view(self.flat_params).copy_(p.detach())
            p.data = view(self.flat_params)
            p.grad = view(self.flat_grads)
            off += k
        # AffineChannel folding of every trainable fused convolution in ONE multi-tensor launch per step (see _ConvAffine)
        self._fold = []
        if fused_body:
            for m in self.student.modules():
                if isinstance(m, Bottleneck) and m.fused:
                    pairs = [(m.c1, m.a1), (m.c2, m.a2), (m.c3, m.a3)] + ([(m.short[0], m.short[1])] if m.short is not None else [])
                    self._fold += [(c, a) for c, a in pairs if c.weight.requires_grad]
        self._fold_dw = []
        self._fold_w = [c.weight.detach() for c, _ in self._fold]
        self._fold_g = [c.weight.grad for c, _ in self._fold]
        self.refold()
        self.n_head, self.n_body = n_head, n_body
        self.exchange = parallel.make_exchange(self.flat_grads, world=self.world, rank=self.rank)
        # buckets of the flat gradient buffer in the order the backward pass completes them:
        #   "head" (after the head's backward) -> "c4" (res5 + FPN weights, when d(c4) is formed) -> "c3" (res4 weights) ->
        #   "end" (res3 weights + every body bias)
        stage_of = {}
        for name in ("res3", "res4", "res5", "lat", "out", "p6", "p7"):
            for q in getattr(self.student, name).parameters():
                stage_of[id(q)] = name
        ranges, off2 = {}, n_head
        for q in self.body_params:
            key = ("bias" if q.dim() == 1 else stage_of[id(q)])
            lo, hi = ranges.get(key, (off2, off2))
            assert hi == off2, "bucket %s is not contiguous in the flat buffer" % key
            ranges[key] = (lo, off2 + q.numel())
            off2 += q.numel()
        fpn_lo = min(ranges[k][0] for k in ("res5", "lat", "out", "p6", "p7"))
        fpn_hi = max(ranges[k][1] for k in ("res5", "lat", "out", "p6", "p7"))
        assert fpn_hi - fpn_lo == sum(ranges[k][1] - ranges[k][0] for k in ("res5", "lat", "out", "p6", "p7"))
        self.buckets = {"head": [(0, n_head)], "c4": [(fpn_lo, fpn_hi)], "c3": [ranges["res4"]],
                        "end": [ranges["res3"]] + ([ranges["bias"]] if "bias" in ranges else [])}
        assert sum(hi - lo for rs in self.buckets.values() for lo, hi in rs) == n_head + n_body
        # folded convolutions per bucket (their raw weight gradients are scaled into the flat buffer when the bucket closes)
        conv_stage = {}
        for name in ("res3", "res4", "res5"):
            for m in getattr(self.student, name).modules():
                if isinstance(m, nn.Conv2d):
                    conv_stage[id(m)] = {"res3": "end", "res4": "c3", "res5": "c4"}[name]
        self._fold_bucket = [conv_stage[id(c)] for c, _ in self._fold]
        self.student.grad_ready_hooks = {"c4": lambda: self._bucket_ready("c4"), "c3": lambda: self._bucket_ready("c3")}
        self.momentum = torch.zeros_like(self.flat_params)
        self.lr_dev = torch.tensor(lr, dtype=torch.float32, device=self.device)
        hw, hb = self.head.sgd_segments(weight_decay)
        self.sgd_segments = [hw, hb, (n_body_w, 1.0, weight_decay), (n_body - n_body_w, 2.0, 0.0)]
        # this rank's synthetic shard
        g = torch.Generator(device=self.device).manual_seed(seed + 7919 * (rank + 1))
        N, A = n_images, synthetic.NUM_ANCHORS
        self.images_t = torch.randn(N, 3, H, W, device=self.device, generator=g).contiguous(memory_format=torch.channels_last)
        self.labels = []
        for h, w in shapes:
            u = torch.rand(N, A, h, w, device=self.device, generator=g)
            lab = torch.zeros(N, A, h, w, dtype=torch.int32, device=self.device)
            lab[u < 0.005] = -1
            fg = (u >= 0.005) & (u < 0.006)
            lab[fg] = torch.randint(1, synthetic.NUM_CLASSES + 1, (int(fg.sum()),), device=self.device, generator=g, dtype=torch.int32)
            self.labels.append(lab)
        self.fg_num = torch.stack([(l > 0).sum() for l in self.labels]).sum().float().reshape(1)
        # foreground anchors as SelectSmoothL1Loss takes them (roi_data/retinanet.py:178-195): rows {image, 4 * anchor, y, x} as floats
        # and an (M, 4) target per level
        self.box_locs, self.box_targets = [], []
        for lab in self.labels:
            idx = (lab > 0).nonzero()
            locs = torch.stack([idx[:, 0], idx[:, 1] * 4, idx[:, 2], idx[:, 3]], dim=1).float().contiguous()
            self.box_locs.append(locs)
            self.box_targets.append((torch.randn(locs.shape[0], 4, device=self.device, generator=g) * 0.2).contiguous())
        self.box_ws = [ops.focal_workspace(self.device) for _ in shapes]
        self.box_losses = [torch.zeros((), device=self.device) for _ in shapes]
        self.focal_ws = [ops.focal_workspace(self.device) for _ in shapes]
        self.focal_losses = [torch.zeros((), device=self.device) for _ in shapes]
        self.cls, self.box = self.head.alloc_outputs()
        self.t_prob, self.t_box = self.teacher_head.alloc_outputs()
        self.d_fpn = [torch.empty(N, 256, h, w, device=self.device) for h, w in shapes]
        self.d_box = [torch.empty_like(b) for b in self.box]
        self.loss_scale = 1.0 / self.world                       # detector.py:650-655
        self.plan = ops.DistillPlan(list(zip(self.cls, self.t_prob, self.labels)), power=power, gamma=distill_gamma,
                                    alpha=distill_alpha, beta=0.0, scale=parallel.distill_loss_scale(temperature, self.world),
                                    num_classes=synthetic.NUM_CLASSES, ignored_label=-1)
        self.last = {}

    def _bucket_ready(self, key):
        """Every gradient of bucket `key` has been produced on the current stream: finish the folded convolutions' weight
        gradients of that bucket (grad += dW_folded * s, one multi-tensor launch) and, when overlapping, hand the bucket to
        the exchange (returns at once; the allreduce runs on the exchange's stream behind an event)."""
        if not self._exchanging:
            return
        idx = [i for i, b in enumerate(self._fold_bucket) if b == key and self._fold_dw[i] is not None]
        if idx:
            torch._foreach_addcmul_([self._fold_g[i] for i in idx], [self._fold_dw[i] for i in idx], [self._fold_s[i] for i in idx])
            for i in idx:
                self._fold_dw[i] = None
        if self.world > 1 or self._count_buckets:
            for lo, hi in self.buckets[key]:
                self.exchange.reduce_bucket(lo, hi)

    def refold(self):
        """(Re)build what depends on the frozen AffineChannel values: the per-weight expanded scales of the student and the
        folded weights of the frozen convolutions.  Call after loading AffineChannel parameters."""
        self._fold_s = [torch.empty_like(w).copy_(a.scale.view(-1, 1, 1, 1).expand_as(w)) for w, (_, a) in zip(self._fold_w, self._fold)]
        for net in (self.student, self.teacher):
            for m in net.modules():
                if isinstance(m, nn.Conv2d) and hasattr(m, "_folded"):
                    del m._folded

    def forward_backward(self):
        # The teacher does not depend on the student until the distillation loss: it runs on its own stream beside the
        # student's forward pass (at bs = 2 the res4 / res5 convolutions of either body fill well under 148 SMs on their own).
        main = torch.cuda.current_stream()
        if self.overlap_teacher:
            self.teacher_stream.wait_stream(main)
        with torch.cuda.stream(self.teacher_stream if self.overlap_teacher else main), torch.no_grad():
            # teacher: forward only (model.train = False)
            t_fpn = [f.contiguous() for f in self.teacher(self.images_t)]
            # teacher/retnet_cls_prob_fpnL: the Sigmoid of retinanet_heads.py:153-163 runs in the prediction convolution's epilogue
            self.teacher_head.forward(t_fpn, training=False, out=(self.t_prob, self.t_box))
        self.flat_grads[self.n_head:].zero_()
        if self._fold:
            self._fold_dw = [None] * len(self._fold)
            with torch.no_grad():
                for i, ((conv, _), w_eff) in enumerate(zip(self._fold, torch._foreach_mul(self._fold_w, self._fold_s))):
                    conv._prefolded = (w_eff, (self._fold_dw, i))
        fpn = self.student(self.images_t)                                # PyTorch graph ends here ...
        fpn_c = [f.detach().contiguous() for f in fpn]
        L = len(self.cls)
        self.head.forward(fpn_c, training=True, out=(self.cls, self.box))
        if self.overlap_teacher:
            main.wait_stream(self.teacher_stream)
        self.plan.run()                                                  # PowSum + distillation loss + d(logits), one launch
        for l in range(L):
            # SigmoidFocalLoss + gradient, added into the distillation gradient (the autograd Sum of the two consumers of
            # retnet_cls_pred_fpnL, core.py:695,792-842)
            ops.sigmoid_focal_loss(self.cls[l], self.labels[l], self.fg_num, accumulate_into=self.plan.grads[l], workspace=self.focal_ws[l],
                                   loss_out=self.focal_losses[l], gamma=2.0, alpha=0.25, scale=self.loss_scale,
                                   num_classes=synthetic.NUM_CLASSES)
            # SelectSmoothL1Loss + gradient (RETINANET.BBOX_REG_BETA = 0.11)
            ops.select_smooth_l1_loss(self.box[l], self.box_targets[l], self.box_locs[l], self.fg_num, beta=0.11, scale=self.loss_scale,
                                      workspace=self.box_ws[l], loss_out=self.box_losses[l], grad_out=self.d_box[l])
        d_fpn = self.head.backward(self.plan.grads, self.d_box, want_d_fpn=True, d_fpn=self.d_fpn)
        self._exchanging = self.overlap_exchange and hasattr(self.exchange, "reduce_bucket")
        self._count_buckets = False
        if self._exchanging:
            self._bucket_ready("head")                                   # the head's gradients are final: exchange them under the body's backward
        torch.autograd.backward(fpn, d_fpn)                              # ... and resumes here: FPN and ResNet body backward
        if self._exchanging:
            self._bucket_ready("end")
            self.exchange.join()                                         # the optimiser's stream waits for the last bucket
            assert all(d is None for d in self._fold_dw)
        elif self._fold:
            # grad (zeroed above) += dW_folded * s for every folded convolution: one multi-tensor launch instead of an
            # AccumulateGrad add per parameter
            torch._foreach_addcmul_(self._fold_g, [d for d in self._fold_dw], self._fold_s)
        self.last = {"bbox": self.box_losses, "focal": self.focal_losses, "distill": [x for x in self.plan.losses],
                     "normalizer": self.plan.normalizer}

    def capture(self, warmup=3):
        """Capture forward_backward() (teacher forward, student forward + backward, all losses: several hundred launches,
        cuDNN and this repository's kernels alike) into ONE CUDA graph.  Returns True when the graph was built; on any
        capture error the step stays eager."""
        try:
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(warmup):
                    self.forward_backward()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            if hasattr(self.exchange, "plan_reset"):
                self.exchange.plan_reset()      # buckets announced by a previous capture
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.forward_backward()         # reduce_bucket calls made here only leave event-record nodes in the graph ...
            self.graph = g
            return True
        except Exception as e:  # leave a usable eager step behind
            self.graph, self.capture_error = None, "%s: %s" % (type(e).__name__, str(e)[:200])
            torch.cuda.synchronize()
            return False

    def close(self):
        """Release the step's native objects in the order NCCL needs: the captured graph (it references the exchange's communicator)
        before the exchange, then the heads."""
        torch.cuda.synchronize()
        self.graph = None
        import gc
        gc.collect()
        if hasattr(self.exchange, "close"):
            self.exchange.close()
        self.head.close()
        self.teacher_head.close()

    def run(self):
        if getattr(self, "graph", None) is not None:
            self.graph.replay()
            if self.overlap_exchange and hasattr(self.exchange, "flush"):
                self.exchange.flush()           # ... and the exchanges run here, beside the graph, each behind its bucket's event
        else:
            self.forward_backward()

    def allreduce(self):
        """The step's exchange.  With overlap_exchange the buckets are already on their way (enqueued from inside forward_backward,
        or by run() beside the captured graph); what is left is to make this stream wait for the last of them."""
        if self.overlap_exchange and hasattr(self.exchange, "reduce_bucket"):
            self.exchange.join()
            return None
        return self.exchange.allreduce()

    @torch.no_grad()
    def sgd(self):
        """Scale(2x bias gradients) + WeightedSum weight decay + MomentumSGDUpdate of every trainable blob (optimizer.py:95-130)
        in ONE launch over the flat [head | body] buffers."""
        if self.student_head_f16:
            # mixed fp16 (configs[4]): an overflow of the loss-scaled fp16 gradient tensors shows up as inf / NaN in the head's reduced
            # parameter gradients; the flag makes the optimiser launch a no-op on every rank alike (solver.LossScaler lowers the scale)
            ops.nonfinite_flag(self.flat_grads[:self.n_head], self.overflow)
            ops.momentum_sgd(self.flat_params, self.flat_grads, self.momentum, self.sgd_segments, self.lr_dev, momentum=self.mom,
                             skip_flag=self.overflow)
            return
        ops.momentum_sgd(self.flat_params, self.flat_grads, self.momentum, self.sgd_segments, self.lr_dev, momentum=self.mom)

    def update_loss_scale(self, force=False):
        """Host side of dynamic loss scaling (solver.LossScaler): call once per step; looks at the device flag every few dozen steps,
        halves / doubles the head's loss scale and re-captures the step graph when the scale changed."""
        if not self.student_head_f16:
            return False
        if self.loss_scaler is None:
            from . import solver
            self.loss_scaler = solver.LossScaler(self.head, self.overflow, init_scale=self.head.f16_grad_scale())
        changed = self.loss_scaler.update(force=force)
        if changed and getattr(self, "graph", None) is not None:
            self.capture()
        return changed

    def update_lr(self, cur_iter, solver_cfg):
        """model.UpdateWorkspaceLr(cur_iter, lr_policy.get_lr_at_iter(cur_iter)) (utils/train.py loop, detector.py:598-648): sets the
        device `lr` scalar the optimiser launch reads and rescales the flat update history when the rate jumps."""
        from . import solver
        if getattr(self, "_lr_state", None) is None or self._lr_state.solver is not solver_cfg:
            self._lr_state = solver.LearningRate(solver_cfg, self.lr_dev, [self.momentum])
        return self._lr_state.update(cur_iter)

    def step(self, cur_iter=None, solver_cfg=None):
        if solver_cfg is not None:
            self.update_lr(cur_iter, solver_cfg)
        self.run()
        self.allreduce()
        self.sgd()

    def losses(self):
        return {"bbox": [float(x) for x in self.last["bbox"]], "focal": [float(x) for x in self.last["focal"]],
                "distill": [float(x) for x in self.last["distill"]], "normalizer": float(self.last["normalizer"])}

    def param_count(self):
        return {"head": self.n_head, "body_trainable": self.n_body}
