"""Host mirror of the RetinaNet head on the path: detectron/lib/modeling/retinanet_heads.py:63-245
(`add_fpn_retinanet_outputs`) and its gradient, over the C ABI `sad_head_*` (include/sad_b200.h).

Parameters live under the reference's blob names (level k_min owns them, the other levels share:
retinanet_heads.py:101-152,188-245) and are initialised as the reference does: GaussianFill(std=0.01)
weights, zero biases, classification-prediction bias -log((1 - pi) / pi) with pi = RETINANET.PRIOR_PROB
(retinanet_heads.py:29-60).  All gradients live in ONE flat buffer so the data-parallel exchange is a
single allreduce (the reference issues one NCCLAllreduce per parameter blob, optimizer.py:72-92).
torch supplies device memory and streams only; there is no CPU path.
"""
import ctypes as C
import math

import torch

from . import native
from .native import HeadConfig, HeadTensors, check, lib

K_MIN = 3  # cfg.FPN.RPN_MIN_LEVEL


def param_names(num_convs=4, k_min=K_MIN):
    """Blob names in the order the flat parameter / gradient buffers are laid out: every weight first, then every bias, so
    that the optimiser's two parameter classes (weights: weight decay; biases: 2x gradient, no decay — optimizer.py:115-124)
    are two contiguous segments of the flat buffer (one sad_momentum_sgd_f32 launch)."""
    names = []
    for suffix in ("_w", "_b"):
        for tower in ("cls", "bbox"):
            for i in range(num_convs):
                names.append("retnet_%s_conv_n%d_fpn%d%s" % (tower, i, k_min, suffix))
            names.append("retnet_%s_pred_fpn%d%s" % (tower, k_min, suffix))
    return names


def head_param_count(dim=256, num_convs=4, num_anchors=9, num_classes=80):
    """Elements of the head's flat parameter / gradient buffer (retinanet_heads.py:63-245: two towers of num_convs 3x3 convolutions
    dim -> dim and the two prediction convolutions, each with a bias) — without creating a head."""
    tower = num_convs * (dim * dim * 9 + dim)
    return 2 * tower + (num_anchors * num_classes) * (dim * 9 + 1) + (num_anchors * 4) * (dim * 9 + 1)


class RetinaNetHead:
    def __init__(self, n_images, level_shapes, dim=256, num_convs=4, num_anchors=9, num_classes=80, prior_prob=0.01,
                 device="cuda", seed=0, grad_buffer=None, cls_output_sigmoid=False, param_buffer=None, compute_f16=False, f16_grad_scale=0.0,
                 compute_f32x3=False):
        self.N, self.level_shapes = int(n_images), [tuple(s) for s in level_shapes]
        self.dim, self.num_convs = int(dim), int(num_convs)
        self.cls_out, self.bbox_out = num_anchors * num_classes, num_anchors * 4
        self.device = torch.device(device)
        cfg = HeadConfig()
        lib().sad_head_default_config(C.byref(cfg))
        cfg.n_levels, cfg.N = len(self.level_shapes), self.N
        for l, (h, w) in enumerate(self.level_shapes):
            cfg.H[l], cfg.W[l] = h, w
        cfg.dim, cfg.num_convs, cfg.cls_out, cfg.bbox_out = self.dim, self.num_convs, self.cls_out, self.bbox_out
        # teacher mode (retinanet_heads.py:153-163): the classification output is retnet_cls_prob_fpnL = Sigmoid(logits)
        cfg.cls_output_sigmoid = 1 if cls_output_sigmoid else 0
        # fp16 operands for every convolution of the head, forward and backward (BASELINE.json configs[4]: mixed fp16 compute, fp32
        # accumulation, fp32 parameters and parameter gradients); f16_grad_scale: loss scale of the fp16 gradient tensors (0 = 4096)
        cfg.compute_f16 = 1 if compute_f16 else 0
        cfg.f16_grad_scale = float(f16_grad_scale)
        self.compute_f16 = bool(compute_f16)
        # 3xTF32: the fp32-accurate mode (the reference's head convolution is fp32, conv_op_cudnn.cc:494-498); split [hi | lo] operands
        cfg.compute_f32x3 = 1 if compute_f32x3 else 0
        self.compute_f32x3 = bool(compute_f32x3)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().sad_head_create(C.byref(cfg), C.byref(self.handle)))
        # flat parameter and gradient buffers, views per blob name
        self.names = param_names(self.num_convs)
        shapes = {}
        for n in self.names:
            out = self.dim if "_conv_" in n else (self.cls_out if "_cls_" in n else self.bbox_out)
            shapes[n] = (out, self.dim, 3, 3) if n.endswith("_w") else (out,)
        self.shapes = shapes
        total = sum(math.prod(s) for s in shapes.values())
        if param_buffer is None:
            param_buffer = torch.zeros(total, dtype=torch.float32, device=self.device)
        if not (param_buffer.is_cuda and param_buffer.dtype == torch.float32 and param_buffer.is_contiguous() and param_buffer.numel() == total
                and param_buffer.data_ptr() % 16 == 0):
            raise ValueError("param_buffer must be a contiguous 16-byte aligned CUDA fp32 tensor of %d elements" % total)
        self.flat_params = param_buffer.zero_()
        self.n_weights = sum(math.prod(s) for n, s in shapes.items() if n.endswith("_w"))
        if grad_buffer is None:
            grad_buffer = torch.zeros(total, dtype=torch.float32, device=self.device)
        # a caller-owned slice lets the head's gradients live inside a larger flat buffer (one allreduce for the whole model)
        if not (grad_buffer.is_cuda and grad_buffer.dtype == torch.float32 and grad_buffer.is_contiguous() and grad_buffer.numel() == total
                and grad_buffer.data_ptr() % 16 == 0):
            raise ValueError("grad_buffer must be a contiguous 16-byte aligned CUDA fp32 tensor of %d elements" % total)
        self.flat_grads = grad_buffer
        self.params, self.grads, off = {}, {}, 0
        for n in self.names:
            k = math.prod(shapes[n])
            assert off % 4 == 0  # 16-byte alignment of every view
            self.params[n] = self.flat_params[off:off + k].view(shapes[n])
            self.grads[n] = self.flat_grads[off:off + k].view(shapes[n])
            off += k
        g = torch.Generator(device=self.device).manual_seed(seed)
        for n in self.names:
            if n.endswith("_w"):
                self.params[n].normal_(0.0, 0.01, generator=g)
        self.params["retnet_cls_pred_fpn%d_b" % K_MIN].fill_(-math.log((1.0 - prior_prob) / prior_prob))
        self._w = self._tensors(self.params)
        self._g = self._tensors(self.grads)

    def _tensors(self, d):
        t = HeadTensors()
        for tower, (wa, ba) in (("cls", (t.cls_tower_w, t.cls_tower_b)), ("bbox", (t.bbox_tower_w, t.bbox_tower_b))):
            for i in range(self.num_convs):
                wa[i] = d["retnet_%s_conv_n%d_fpn%d_w" % (tower, i, K_MIN)].data_ptr()
                ba[i] = d["retnet_%s_conv_n%d_fpn%d_b" % (tower, i, K_MIN)].data_ptr()
        t.cls_pred_w, t.cls_pred_b = d["retnet_cls_pred_fpn3_w"].data_ptr(), d["retnet_cls_pred_fpn3_b"].data_ptr()
        t.bbox_pred_w, t.bbox_pred_b = d["retnet_bbox_pred_fpn3_w"].data_ptr(), d["retnet_bbox_pred_fpn3_b"].data_ptr()
        return t

    def views(self, flat_buffer):
        """{blob name: view} over another flat buffer laid out like flat_params (e.g. the momentum buffer: the reference's
        `<name>_momentum` blobs, optimizer.py:103-107) — what weights_io loads and saves by name."""
        if flat_buffer.numel() != self.flat_params.numel():
            raise ValueError("flat buffer must have %d elements" % self.flat_params.numel())
        out, off = {}, 0
        for n in self.names:
            k = math.prod(self.shapes[n])
            out[n] = flat_buffer[off:off + k].view(self.shapes[n])
            off += k
        return out

    def sgd_segments(self, weight_decay):
        """[(count, gradient multiplier, weight decay)] of the flat buffers for ops.momentum_sgd (optimizer.py:115-124)."""
        return [(self.n_weights, 1.0, weight_decay), (self.flat_params.numel() - self.n_weights, 2.0, 0.0)]

    def close(self):
        if getattr(self, "handle", None):
            lib().sad_head_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_f16_grad_scale(self, scale):
        """Loss scale of the fp16 gradient tensors for the following backward passes (re-capture a CUDA graph after changing it)."""
        check(lib().sad_head_set_f16_grad_scale(self.handle, float(scale)))

    def f16_grad_scale(self):
        return float(lib().sad_head_f16_grad_scale(self.handle))

    def device_bytes(self):
        return lib().sad_head_device_bytes(self.handle)

    def _ptrs(self, tensors, channels, name):
        if tensors is None:
            return None
        if len(tensors) != len(self.level_shapes):
            raise ValueError("%s: expected %d levels" % (name, len(self.level_shapes)))
        for t, (h, w) in zip(tensors, self.level_shapes):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (self.N, channels, h, w)):
                raise ValueError("%s: every level must be a contiguous CUDA fp32 tensor (N, %d, H, W)" % (name, channels))
        return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    def activation(self, tower, conv, level):
        """Kept activation of the last forward as an NCHW tensor: tower 'cls' / 'bbox'; conv -1 = the input fpn_L,
        i = output of tower conv i after ReLU (blob retnet_<tower>_conv_n<i>_fpn<L>), tf32-rounded as stored."""
        h, w = self.level_shapes[level]
        cs = lib().sad_conv3x3_split_channels(self.dim)
        shape = (self.N, h, w, 2 * cs) if self.compute_f32x3 else (self.N, h, w, self.dim)
        out = torch.empty(shape, dtype=torch.float16 if self.compute_f16 else torch.float32, device=self.device)
        check(lib().sad_head_copy_activation(self.handle, 0 if tower == "cls" else 1, int(conv), int(level),
                                             C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        if self.compute_f32x3:   # split rows [hi | lo]: the stored value is hi + lo (exact in fp32)
            out = out[..., :self.dim] + out[..., cs:cs + self.dim]
        return out.float().permute(0, 3, 1, 2).contiguous()

    def alloc_outputs(self):
        cls = [torch.empty((self.N, self.cls_out, h, w), dtype=torch.float32, device=self.device) for h, w in self.level_shapes]
        box = [torch.empty((self.N, self.bbox_out, h, w), dtype=torch.float32, device=self.device) for h, w in self.level_shapes]
        return cls, box

    def forward(self, fpn, training=True, out=None):
        """fpn: list of (N, dim, H_l, W_l), finest level first.  Returns (cls_logits, bbox_preds), NCHW."""
        cls, box = out if out is not None else self.alloc_outputs()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib().sad_head_forward(self.handle, C.byref(self._w), self._ptrs(fpn, self.dim, "fpn"),
                                     self._ptrs(cls, self.cls_out, "cls_logits"), self._ptrs(box, self.bbox_out, "bbox_pred"),
                                     1 if training else 0, st))
        return cls, box

    def backward(self, d_cls, d_bbox, want_d_fpn=True, accumulate=False, d_fpn=None):
        """Gradients of every head parameter (into self.grads / self.flat_grads) and, optionally, of the FPN inputs."""
        if want_d_fpn and d_fpn is None:
            d_fpn = [torch.empty((self.N, self.dim, h, w), dtype=torch.float32, device=self.device) for h, w in self.level_shapes]
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib().sad_head_backward(self.handle, C.byref(self._w), self._ptrs(d_cls, self.cls_out, "d_cls_logits"),
                                      self._ptrs(d_bbox, self.bbox_out, "d_bbox_pred"), C.byref(self._g),
                                      self._ptrs(d_fpn, self.dim, "d_fpn") if want_d_fpn else None,
                                      1 if accumulate else 0, st))
        return d_fpn if want_d_fpn else None
