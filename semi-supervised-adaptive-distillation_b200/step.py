"""One data-parallel distillation step of the RetinaNet head on one GPU's image shard — the hot path end to end:

    student head forward (retinanet_heads.py:63-245)  ->  PowSum over the teacher probabilities (:320-328)
    ->  SigmoidAdaptiveDistillLoss + Gradient per level (:331-348, fused, one launch)
    ->  head backward (ConvGradient / ReluGradient / Sum, core.py:695-842)
    ->  ONE allreduce of the flat head-gradient buffer (optimizer.py:72-92 issues one per parameter blob)
    ->  MomentumSGDUpdate with the bias / weight-decay preamble (optimizer.py:95-130): one launch over the flat buffers

The path shards by image (optimizer.py:62-69 builds the ops per GPU scope; the normaliser is per GPU), so the
only exchange is the gradient allreduce; the loss is pre-scaled by 1 / world (detector.py:650-655) so the SUM is the mean.
The device work of one step (everything but the collective) can be captured into a CUDA graph: 60-odd launches
replayed with one host call — the reference synchronises the stream after every operator (operator.h:369-382).
"""
import torch

from . import ops, parallel, synthetic
from .head import RetinaNetHead


class DistillHeadStep:
    def __init__(self, n_images=2, scale_px=600, world=1, rank=0, seed=1234, temperature=1.0, power=1.8, alpha=0.5,
                 gamma=2.0, beta=0.0, device=None, level_shapes=None, dim=256, num_convs=4, with_bbox_branch=True, compute_f16=False, compute_f32x3=False):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.world, self.rank = int(world), int(rank)
        shapes = level_shapes if level_shapes is not None else synthetic.level_shapes(scale_px)
        # same weights on every rank; compute_f16: the head's convolutions (forward + backward) on fp16 operands, fp32 everything else
        self.head = RetinaNetHead(n_images, shapes, dim=dim, num_convs=num_convs, device=self.device, seed=seed, compute_f16=compute_f16,
                                  compute_f32x3=compute_f32x3)
        g = torch.Generator(device=self.device).manual_seed(seed + 7919 * (rank + 1))              # different images per rank
        N, A, Cc = n_images, synthetic.NUM_ANCHORS, synthetic.NUM_CLASSES
        # synthetic FPN features: post-conv, zero-mean
        self.fpn = [torch.randn(N, dim, h, w, device=self.device, generator=g) * 0.5 for h, w in shapes]
        self.teacher = [torch.sigmoid(torch.randn(N, A * Cc, h, w, device=self.device, generator=g) * 2.5 + synthetic.CLS_BIAS)
                        .clamp_(1e-6, 1 - 1e-6) for h, w in shapes]
        self.labels = []
        for h, w in shapes:
            u = torch.rand(N, A, h, w, device=self.device, generator=g)
            lab = torch.zeros(N, A, h, w, dtype=torch.int32, device=self.device)
            lab[u < 0.005] = -1
            lab[(u >= 0.005) & (u < 0.006)] = 1
            self.labels.append(lab)
        self.cls, self.box = self.head.alloc_outputs()
        # box-regression gradient stand-in (SelectSmoothL1LossGradient is an §8f "next" row): a fixed dense tensor so
        # that the box tower's backward does its full work
        self.d_box = [torch.randn(N, self.head.bbox_out, h, w, device=self.device, generator=g) * 1e-3 for h, w in shapes] \
            if with_bbox_branch else None
        self.d_fpn = [torch.empty_like(x) for x in self.fpn]
        self.plan = ops.DistillPlan(list(zip(self.cls, self.teacher, self.labels)), power=power, gamma=gamma, alpha=alpha, beta=beta,
                                    scale=parallel.distill_loss_scale(temperature, self.world), num_classes=Cc, ignored_label=-1)
        self.exchange = parallel.make_exchange(self.head.flat_grads, world=self.world, rank=self.rank)
        self.momentum = torch.zeros_like(self.head.flat_params)
        self.lr = torch.tensor(0.01, dtype=torch.float32, device=self.device)
        self.graph = None
        self.images = N
        self.anchors = int(sum(l.numel() for l in self.labels))

    # ---- device work of one step (no collective) ----
    def forward_backward(self):
        self.head.forward(self.fpn, training=True, out=(self.cls, self.box))
        self.plan.run()                                   # PowSum -> normaliser; fused loss + d(logits) for all levels
        self.head.backward(self.plan.grads, self.d_box, want_d_fpn=True, d_fpn=self.d_fpn)

    def capture(self, warmup=3):
        """Capture forward_backward() into a CUDA graph (call once; run() then replays it)."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.forward_backward()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.forward_backward()
        return self

    def close(self):
        torch.cuda.synchronize()
        self.graph = None
        if hasattr(self.exchange, "close"):
            self.exchange.close()
        self.head.close()

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.forward_backward()

    def allreduce(self):
        """The step's only collective: SUM of the flat head-gradient buffer over the ranks."""
        return self.exchange.allreduce()

    def sgd(self, momentum=0.9, weight_decay=1e-4):
        """Scale(2x bias gradients) + WeightedSum weight decay + MomentumSGDUpdate of all 20 head blobs (optimizer.py:95-130)
        in ONE launch over the flat buffers; the learning rate is the device scalar self.lr (the `lr` blob)."""
        ops.momentum_sgd(self.head.flat_params, self.head.flat_grads, self.momentum, self.head.sgd_segments(weight_decay), self.lr,
                         momentum=momentum)

    def losses(self):
        return [l.item() for l in self.plan.losses]

    # 2 * pixels * Cout * 9 * Cin per conv, forward; backward = data gradient + weight gradient = 2x
    def flops(self):
        pix = sum(x.shape[0] * x.shape[2] * x.shape[3] for x in self.fpn)
        d, nc = self.head.dim, self.head.num_convs
        fwd = 2.0 * pix * 9 * d * (2 * nc * d + self.head.cls_out + self.head.bbox_out)
        return fwd, 2.0 * fwd
