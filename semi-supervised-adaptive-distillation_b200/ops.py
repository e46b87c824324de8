"""torch-tensor convenience layer over the C ABI (torch supplies device memory and streams only).

Every function enqueues on torch's current CUDA stream and returns device tensors; nothing here
computes on the CPU and nothing falls back if the CUDA library is unavailable.
"""
import ctypes as C

import torch

from . import native
from .native import ConvLevel, DistillLevel, DistillParams, HostLevel, WgradLevel, check, default_params, lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise ValueError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous (NCHW)" % name)


class Workspace:
    """Caller-owned kernel scratch (see include/sad_b200.h: sad_workspace_init)."""

    def __init__(self, nbytes, device):
        nbytes = max(256, (int(nbytes) + 255) // 256 * 256)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        assert self.buf.data_ptr() % 256 == 0
        check(lib().sad_workspace_init(C.c_void_p(self.buf.data_ptr()), nbytes, _stream()))
        self.nbytes = nbytes

    @property
    def ptr(self):
        return C.c_void_p(self.buf.data_ptr())


def pow_sum_workspace(tensors):
    sizes = (C.c_int64 * len(tensors))(*[t.numel() for t in tensors])
    return Workspace(lib().sad_pow_sum_workspace_bytes(sizes, len(tensors)), tensors[0].device)


def pow_sum(tensors, power=1.0, out=None, workspace=None):
    """PowSum operator: scalar sum of x**power over all inputs (reference pow_sum_op.cu:25-43)."""
    tensors = list(tensors)
    if not 1 <= len(tensors) <= native.SAD_MAX_INPUTS:
        raise ValueError("PowSum takes 1..%d inputs" % native.SAD_MAX_INPUTS)
    for i, t in enumerate(tensors):
        _require_cuda(t, torch.float32, "input %d" % i)
    if out is None:
        out = torch.empty((), dtype=torch.float32, device=tensors[0].device)
    if workspace is None:
        workspace = pow_sum_workspace(tensors)
    ptrs = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    sizes = (C.c_int64 * len(tensors))(*[t.numel() for t in tensors])
    check(lib().sad_pow_sum_f32(ptrs, sizes, len(tensors), float(power), C.c_void_p(out.data_ptr()),
                                workspace.ptr, workspace.nbytes, _stream()))
    return out


def _levels_struct(levels, want_loss, want_grad, d_loss):
    n = len(levels)
    arr = (DistillLevel * n)()
    losses, grads = [], []
    for i, (x, t, g) in enumerate(levels):
        _require_cuda(x, torch.float32, "logits[%d]" % i)
        _require_cuda(t, torch.float32, "teacher_prob[%d]" % i)
        _require_cuda(g, torch.int32, "labels[%d]" % i)
        if x.dim() != 4 or t.shape != x.shape:
            raise ValueError("logits/teacher_prob must be 4-D of equal shape")
        L = arr[i]
        L.logits, L.teacher_prob, L.labels = x.data_ptr(), t.data_ptr(), g.data_ptr()
        L.N, L.D, L.H, L.W = x.shape
        if want_loss:
            losses.append(torch.empty((), dtype=torch.float32, device=x.device))
            L.loss = losses[-1].data_ptr()
        if want_grad:
            grads.append(torch.empty_like(x))
            L.d_logits = grads[-1].data_ptr()
            if d_loss is not None:
                dl = d_loss[i] if isinstance(d_loss, (list, tuple)) else d_loss
                _require_cuda(dl, torch.float32, "d_loss")
                L.d_loss = dl.data_ptr()
    return arr, losses, grads


def distill_workspace(levels):
    arr, _, _ = _levels_struct(levels, False, False, None)
    return Workspace(lib().sad_distill_workspace_bytes(arr, len(levels)), levels[0][0].device)


def distill(levels, normalizer, want_loss=True, want_grad=True, d_loss=None, workspace=None, **args):
    """SigmoidAdaptiveDistillLoss and/or its gradient for a list of (logits, teacher_prob, labels)
    levels in one launch.  Returns (losses, d_logits) lists (empty when not requested).
    Keyword args are the operator arguments: gamma, alpha, beta, scale, num_classes, ignored_label."""
    levels = list(levels)
    params = default_params(**args)
    _require_cuda(normalizer, torch.float32, "normalizer")
    arr, losses, grads = _levels_struct(levels, want_loss, want_grad, d_loss)
    if want_loss and workspace is None:
        workspace = distill_workspace(levels)
    wptr, wbytes = (workspace.ptr, workspace.nbytes) if workspace is not None else (None, 0)
    check(lib().sad_distill_f32(arr, len(levels), C.c_void_p(normalizer.data_ptr()), C.byref(params), wptr, wbytes,
                                _stream()))
    return losses, grads


def focal_workspace(device):
    return Workspace(lib().sad_focal_workspace_bytes(), device)


def sigmoid_focal_loss(logits, labels, fg_num, want_loss=True, want_grad=True, d_loss=None, accumulate_into=None, workspace=None,
                       loss_out=None, gamma=1.0, alpha=0.25, scale=1.0, num_classes=80):
    """SigmoidFocalLoss and/or its gradient for one level.  Returns (loss or None, d_logits or None).
    accumulate_into: an existing d_logits tensor the gradient is ADDED to (the autograd Sum with the distillation gradient)."""
    _require_cuda(logits, torch.float32, "logits")
    _require_cuda(labels, torch.int32, "labels")
    _require_cuda(fg_num, torch.float32, "fg_num")
    n, d, h, w = logits.shape
    p = native.FocalParams()
    lib().sad_focal_default_params(C.byref(p))
    p.gamma, p.alpha, p.scale, p.num_classes = float(gamma), float(alpha), float(scale), int(num_classes)
    loss = (loss_out if loss_out is not None else torch.empty((), dtype=torch.float32, device=logits.device)) if want_loss else None
    grad = None
    if want_grad:
        grad = accumulate_into if accumulate_into is not None else torch.empty_like(logits)
    if want_loss and workspace is None:
        workspace = Workspace(lib().sad_focal_workspace_bytes(), logits.device)
    check(lib().sad_sigmoid_focal_loss_f32(
        C.c_void_p(logits.data_ptr()), C.c_void_p(labels.data_ptr()), C.c_void_p(fg_num.data_ptr()), n, d, h, w, C.byref(p),
        C.c_void_p(loss.data_ptr()) if loss is not None else None, C.c_void_p(d_loss.data_ptr()) if d_loss is not None else None,
        C.c_void_p(grad.data_ptr()) if grad is not None else None, 1 if accumulate_into is not None else 0,
        workspace.ptr if workspace is not None else None, workspace.nbytes if workspace is not None else 0, _stream()))
    return loss, grad


def select_smooth_l1_loss(y_hat, y, locs, fg_num, beta=1.0, scale=1.0, want_loss=True, want_grad=True, d_loss=None, workspace=None,
                          loss_out=None, grad_out=None):
    """SelectSmoothL1Loss and/or its gradient for one level: y_hat (N, A*4, H, W), y (M, 4), locs (M, 4) float {n, c, y, x}."""
    _require_cuda(y_hat, torch.float32, "y_hat")
    m = int(y.shape[0]) if y.numel() else 0
    if m:
        _require_cuda(y, torch.float32, "y")
        _require_cuda(locs, torch.float32, "locs")
    n, d, h, w = y_hat.shape
    loss = (loss_out if loss_out is not None else torch.empty((), dtype=torch.float32, device=y_hat.device)) if want_loss else None
    grad = (grad_out if grad_out is not None else torch.empty_like(y_hat)) if want_grad else None
    if want_loss and workspace is None:
        workspace = Workspace(lib().sad_smooth_l1_workspace_bytes(), y_hat.device)
    check(lib().sad_select_smooth_l1_loss_f32(
        C.c_void_p(y_hat.data_ptr()), C.c_void_p(y.data_ptr()) if m else None, C.c_void_p(locs.data_ptr()) if m else None,
        C.c_void_p(fg_num.data_ptr()), n, d, h, w, m, float(beta), float(scale),
        C.c_void_p(loss.data_ptr()) if loss is not None else None, C.c_void_p(d_loss.data_ptr()) if d_loss is not None else None,
        C.c_void_p(grad.data_ptr()) if grad is not None else None,
        workspace.ptr if workspace is not None else None, workspace.nbytes if workspace is not None else 0, _stream()))
    return loss, grad


def nonfinite_flag(x, flag):
    """flag |= 1 (CUDA int32 scalar tensor) when any element of the fp32 CUDA tensor x is inf or NaN; no host synchronisation."""
    _require_cuda(x, torch.float32, "x")
    _require_cuda(flag, torch.int32, "flag")
    check(lib().sad_nonfinite_flag_f32(C.c_void_p(x.data_ptr()), x.numel(), C.c_void_p(flag.data_ptr()), _stream()))
    return flag


def momentum_sgd(param, grad, momentum_buf, segments, lr, momentum=0.9, nesterov=False, skip_flag=None):
    """In-place momentum SGD over flat buffers.  segments: [(count, grad_multiplier, weight_decay)] tiling the buffers;
    lr: CUDA fp32 scalar tensor (the `lr` blob, updated by the learning-rate policy between steps).  skip_flag: CUDA int32 scalar
    tensor; when it is not 0 the launch leaves every buffer untouched (an overflowed mixed-precision step is skipped on the device)."""
    for name, t in (("param", param), ("grad", grad), ("momentum", momentum_buf)):
        _require_cuda(t, torch.float32, name)
    _require_cuda(lr, torch.float32, "lr")
    arr = (native.SgdSegment * len(segments))()
    for i, (cnt, mult, wd) in enumerate(segments):
        arr[i].count, arr[i].grad_multiplier, arr[i].weight_decay = int(cnt), float(mult), float(wd)
    if sum(int(c) for c, _, _ in segments) != param.numel() or grad.numel() != param.numel() or momentum_buf.numel() != param.numel():
        raise ValueError("the segments must tile the flat buffers exactly")
    if skip_flag is not None:
        _require_cuda(skip_flag, torch.int32, "skip_flag")
        check(lib().sad_momentum_sgd_guarded_f32(C.c_void_p(param.data_ptr()), C.c_void_p(grad.data_ptr()), C.c_void_p(momentum_buf.data_ptr()),
                                                 arr, len(segments), C.c_void_p(lr.data_ptr()), float(momentum), 1 if nesterov else 0,
                                                 C.c_void_p(skip_flag.data_ptr()), _stream()))
        return
    check(lib().sad_momentum_sgd_f32(C.c_void_p(param.data_ptr()), C.c_void_p(grad.data_ptr()), C.c_void_p(momentum_buf.data_ptr()),
                                     arr, len(segments), C.c_void_p(lr.data_ptr()), float(momentum), 1 if nesterov else 0, _stream()))


def distill_step(levels, power=1.8, workspace=None, d_loss=None, **args):
    """PowSum over the levels' teacher probabilities + loss + gradient of every level through the single-launch entry
    point sad_distill_fused_f32.  Returns (normalizer, losses, d_logits)."""
    levels = list(levels)
    params = default_params(**args)
    arr, losses, grads = _levels_struct(levels, True, True, d_loss)
    dev = levels[0][0].device
    norm = torch.empty((), dtype=torch.float32, device=dev)
    if workspace is None:
        workspace = Workspace(lib().sad_distill_fused_workspace_bytes(arr, len(levels), params.num_classes), dev)
    check(lib().sad_distill_fused_f32(arr, len(levels), float(power), C.c_void_p(norm.data_ptr()), C.byref(params),
                                      workspace.ptr, workspace.nbytes, _stream()))
    return norm, losses, grads


class DistillPlan:
    """Pre-bound PowSum + fused multi-level loss+grad for fixed device tensors: the launch
    descriptors are built once so the timed loop only enqueues two kernels."""

    def __init__(self, levels, power=1.8, **args):
        self.levels = list(levels)
        self.params = default_params(**args)
        dev = self.levels[0][0].device
        self.teacher = [t for (_, t, _) in self.levels]
        n = len(self.levels)
        self.power = float(power)
        self.normalizer = torch.empty((), dtype=torch.float32, device=dev)
        self.arr, self.losses, self.grads = _levels_struct(self.levels, True, True, None)
        self.ws_pow = pow_sum_workspace(self.teacher)
        self.ws_dist = distill_workspace(self.levels)
        self._ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in self.teacher])
        self._sizes = (C.c_int64 * n)(*[t.numel() for t in self.teacher])
        self.n = n
        self.ws_fused = Workspace(lib().sad_distill_fused_workspace_bytes(self.arr, n, self.params.num_classes), dev)

    def run(self):
        """The whole loss step in one launch (sad_distill_fused_f32)."""
        check(lib().sad_distill_fused_f32(self.arr, self.n, self.power, C.c_void_p(self.normalizer.data_ptr()),
                                          C.byref(self.params), self.ws_fused.ptr, self.ws_fused.nbytes, _stream()))

    def run_two_launches(self):
        """PowSum, then the fused multi-level loss + gradient: the operator-by-operator form."""
        st = _stream()
        l = lib()
        check(l.sad_pow_sum_f32(self._ptrs, self._sizes, self.n, self.power, C.c_void_p(self.normalizer.data_ptr()),
                                self.ws_pow.ptr, self.ws_pow.nbytes, st))
        check(l.sad_distill_f32(self.arr, self.n, C.c_void_p(self.normalizer.data_ptr()), C.byref(self.params),
                                self.ws_dist.ptr, self.ws_dist.nbytes, st))

    def run_pow_sum(self):
        check(lib().sad_pow_sum_f32(self._ptrs, self._sizes, self.n, self.power,
                                    C.c_void_p(self.normalizer.data_ptr()), self.ws_pow.ptr, self.ws_pow.nbytes, _stream()))

    def run_distill(self):
        check(lib().sad_distill_f32(self.arr, self.n, C.c_void_p(self.normalizer.data_ptr()), C.byref(self.params),
                                    self.ws_dist.ptr, self.ws_dist.nbytes, _stream()))


class HostStep:
    """sad_distill_step_host: the whole loss step on HOST (ideally pinned) tensors."""

    def __init__(self, device=0):
        self.handle = C.c_void_p()
        check(lib().sad_ctx_create(int(device), C.byref(self.handle)))

    def set_chunk_bytes(self, nbytes):
        """Pipeline granularity (sad_ctx_set_host_chunk_bytes): logits bytes per chunk of whole anchors; 0 = default (16 MB)."""
        check(lib().sad_ctx_set_host_chunk_bytes(self.handle, int(nbytes)))

    def close(self):
        if self.handle:
            lib().sad_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bind(self, levels, grads_out, power=1.8, **args):
        """levels: [(logits, teacher_prob, labels)] CPU tensors; grads_out: CPU tensors or None."""
        n = len(levels)
        self._keep = (levels, grads_out)
        self._arr = (HostLevel * n)()
        for i, (x, t, g) in enumerate(levels):
            for name, ten, dt in (("logits", x, torch.float32), ("teacher_prob", t, torch.float32), ("labels", g, torch.int32)):
                if ten.is_cuda or ten.dtype != dt or not ten.is_contiguous():
                    raise ValueError("%s[%d] must be a contiguous CPU %s tensor" % (name, i, dt))
            L = self._arr[i]
            L.logits, L.teacher_prob, L.labels = x.data_ptr(), t.data_ptr(), g.data_ptr()
            L.d_logits = grads_out[i].data_ptr() if grads_out is not None else None
            L.N, L.D, L.H, L.W = x.shape
        self._n = n
        self._power = float(power)
        self._params = default_params(**args)
        self._losses = (C.c_float * n)()
        self._norm = C.c_float()

    def run(self):
        check(lib().sad_distill_step_host(self.handle, self._arr, self._n, self._power, C.byref(self._params),
                                          self._losses, C.byref(self._norm)))
        return list(self._losses), self._norm.value


# ---------------------------------------------------------------------------------------------
# RetinaNet head convolution (3x3, stride 1, pad 1, NCHW fp32; weights shared by all levels)
# ---------------------------------------------------------------------------------------------
def conv3x3_pack(weight, mode=0):
    """Repack (Cout, Cin, 3, 3) weights for the tensor-core kernels: mode 0 forward, 1 data gradient."""
    _require_cuda(weight, torch.float32, "weight")
    cout, cin = weight.shape[0], weight.shape[1]
    if tuple(weight.shape[2:]) != (3, 3):
        raise ValueError("weight must be (Cout, Cin, 3, 3)")
    packed = torch.empty(9 * cin * cout, dtype=torch.float32, device=weight.device)
    check(lib().sad_conv3x3_pack_weights_f32(C.c_void_p(weight.data_ptr()), cin, cout, int(mode),
                                             C.c_void_p(packed.data_ptr()), _stream()))
    return packed


def to_nhwc(xs):
    """NCHW fp32 -> channels-last (N, H, W, C) fp32 rounded to tf32, every level in one launch."""
    xs = list(xs)
    arr = (native.LayoutLevel * len(xs))()
    outs = []
    for i, x in enumerate(xs):
        _require_cuda(x, torch.float32, "x[%d]" % i)
        if x.shape[1] != xs[0].shape[1]:
            raise ValueError("all levels must have the same channel count")
        n, c, h, w = x.shape
        outs.append(torch.empty((n, h, w, c), dtype=torch.float32, device=x.device))
        arr[i].src_nchw, arr[i].dst_nhwc = x.data_ptr(), outs[-1].data_ptr()
        arr[i].N, arr[i].H, arr[i].W = n, h, w
    check(lib().sad_nchw_to_nhwc_f32(arr, len(xs), xs[0].shape[1], _stream()))
    return outs


def sign_bits_like(n, c, h, w, device):
    nbytes = lib().sad_conv3x3_sign_bits_bytes(n, c, h, w)
    return torch.empty(max(1, nbytes // 4), dtype=torch.int32, device=device)


def _conv_run(xs_nhwc, packed, bias, k, m, relu, want_nchw, want_nhwc, masks_nhwc=None, bits_in=None, want_bits=False):
    arr = (ConvLevel * len(xs_nhwc))()
    ys, yts, bits = [], [], []
    for i, xt in enumerate(xs_nhwc):
        _require_cuda(xt, torch.float32, "x_nhwc[%d]" % i)
        n, h, w, c = xt.shape
        if c != k:
            raise ValueError("input channels do not match the weights")
        arr[i].x_nhwc = xt.data_ptr()
        arr[i].N, arr[i].H, arr[i].W = n, h, w
        if masks_nhwc is not None:
            _require_cuda(masks_nhwc[i], torch.float32, "relu_mask_nhwc[%d]" % i)
            if tuple(masks_nhwc[i].shape) != (n, h, w, m):
                raise ValueError("relu mask must be channels-last (N, H, W, Cout_of_this_pass)")
            arr[i].relu_mask_nhwc = masks_nhwc[i].data_ptr()
        if bits_in is not None:
            arr[i].relu_bits_in = bits_in[i].data_ptr()
        if want_bits:
            bits.append(sign_bits_like(n, m, h, w, xt.device))
            arr[i].relu_bits_out = bits[-1].data_ptr()
        if want_nchw:
            ys.append(torch.empty((n, m, h, w), dtype=torch.float32, device=xt.device))
            arr[i].y_nchw = ys[-1].data_ptr()
        if want_nhwc:
            yts.append(torch.empty((n, h, w, m), dtype=torch.float32, device=xt.device))
            arr[i].y_nhwc = yts[-1].data_ptr()
    b = C.c_void_p(bias.data_ptr()) if bias is not None else None
    check(lib().sad_conv3x3_fwd_f32(arr, len(xs_nhwc), C.c_void_p(packed.data_ptr()), b, k, m, 1 if relu else 0, _stream()))
    return (ys, yts, bits) if want_bits else (ys, yts)


def conv3x3_forward(xs, weight, bias=None, relu=False, packed=None, xs_nhwc=None, want_nchw=True, want_nhwc=False, want_bits=False):
    """Conv (+bias, + optional fused ReLU) of every level in one launch.  xs: list of (N, Cin, H, W)
    (or pass xs_nhwc, channels-last copies from to_nhwc / a previous call).  Returns (ys_nchw, ys_nhwc)."""
    cout, cin = weight.shape[0], weight.shape[1]
    if packed is None:
        packed = conv3x3_pack(weight, 0)
    if xs_nhwc is None:
        xs_nhwc = to_nhwc(xs)
    return _conv_run(xs_nhwc, packed, bias, cin, cout, relu, want_nchw, want_nhwc, want_bits=want_bits)


def conv3x3_dgrad(dys, weight, packed=None, dys_nhwc=None, want_nchw=True, want_nhwc=False, relu_masks_nhwc=None, relu_bits=None):
    """Data gradient dX = conv(dY, W^T with flipped taps) of every level in one launch.  dys: (N, Cout, H, W).
    relu_masks_nhwc: channels-last forward outputs Y (N, H, W, Cin) of the layer below; fuses its ReluGradient."""
    cout, cin = weight.shape[0], weight.shape[1]
    if packed is None:
        packed = conv3x3_pack(weight, 1)
    if dys_nhwc is None:
        dys_nhwc = to_nhwc(dys)
    return _conv_run(dys_nhwc, packed, None, cout, cin, False, want_nchw, want_nhwc, relu_masks_nhwc, bits_in=relu_bits)


def conv3x3_wgrad(xs_nhwc, dys_nhwc, want_bias=True, accumulate_into=None, workspace=None):
    """Weight (+ bias) gradient summed over every level in one launch.  xs_nhwc: channels-last forward inputs
    (N, H, W, Cin); dys_nhwc: channels-last output gradients (N, H, W, Cout).  Returns (dW (Cout, Cin, 3, 3), db)."""
    n = len(xs_nhwc)
    arr = (WgradLevel * n)()
    cin, cout = xs_nhwc[0].shape[3], dys_nhwc[0].shape[3]
    for i, (xt, dt) in enumerate(zip(xs_nhwc, dys_nhwc)):
        _require_cuda(xt, torch.float32, "x_nhwc[%d]" % i)
        _require_cuda(dt, torch.float32, "dy_nhwc[%d]" % i)
        if xt.shape[:3] != dt.shape[:3] or xt.shape[3] != cin or dt.shape[3] != cout:
            raise ValueError("level %d: x (N,H,W,Cin) and dy (N,H,W,Cout) do not match" % i)
        arr[i].x_nhwc, arr[i].dy_nhwc = xt.data_ptr(), dt.data_ptr()
        arr[i].N, arr[i].H, arr[i].W = xt.shape[:3]
    dev = xs_nhwc[0].device
    if accumulate_into is not None:
        dw, db = accumulate_into
    else:
        dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=dev)
        db = torch.empty((cout,), dtype=torch.float32, device=dev) if want_bias else None
    if workspace is None:
        nbytes = lib().sad_conv3x3_wgrad_workspace_bytes(arr, n, cin, cout)
        workspace = torch.empty(max(256, nbytes), dtype=torch.uint8, device=dev)
    check(lib().sad_conv3x3_wgrad_f32(arr, n, cin, cout, C.c_void_p(dw.data_ptr()),
                                      C.c_void_p(db.data_ptr()) if db is not None else None,
                                      1 if accumulate_into is not None else 0,
                                      C.c_void_p(workspace.data_ptr()), workspace.numel(), _stream()))
    return dw, db


def affine_channel(x, scale, bias, out=None):
    """AffineChannel (affine_channel_op.cu:52-75): y = x * scale[c] + bias[c] over (N, C, H, W); out may be x (in place)."""
    _require_cuda(x, torch.float32, "x")
    _require_cuda(scale, torch.float32, "scale")
    _require_cuda(bias, torch.float32, "bias")
    if x.dim() != 4 or scale.numel() != x.shape[1] or bias.numel() != x.shape[1]:
        raise ValueError("x must be (N, C, H, W) and scale / bias must have C elements")
    y = torch.empty_like(x) if out is None else out
    check(lib().sad_affine_channel_f32(C.c_void_p(x.data_ptr()), C.c_void_p(scale.data_ptr()), C.c_void_p(bias.data_ptr()),
                                       C.c_void_p(y.data_ptr()), x.shape[0], x.shape[1], x.shape[2] * x.shape[3], _stream()))
    return y


def affine_channel_grad(scale, dy, out=None):
    """AffineChannelGradient (affine_channel_op.cu:77-98): dx = dy * scale[c]; out may be dy (in place)."""
    _require_cuda(dy, torch.float32, "dy")
    _require_cuda(scale, torch.float32, "scale")
    if dy.dim() != 4 or scale.numel() != dy.shape[1]:
        raise ValueError("dy must be (N, C, H, W) and scale must have C elements")
    dx = torch.empty_like(dy) if out is None else out
    check(lib().sad_affine_channel_f32(C.c_void_p(dy.data_ptr()), C.c_void_p(scale.data_ptr()), None, C.c_void_p(dx.data_ptr()),
                                       dy.shape[0], dy.shape[1], dy.shape[2] * dy.shape[3], _stream()))
    return dx


def _outer_hw(t):
    if t.dim() not in (3, 4):
        raise ValueError("UpsampleNearest takes a 3-D or 4-D tensor")
    outer = 1
    for d in t.shape[:-2]:
        outer *= d
    return outer, t.shape[-2], t.shape[-1]


def upsample_nearest(x, scale=2):
    """UpsampleNearest (upsample_nearest_op.cu:116-158): nearest-neighbour upsampling of the last two dimensions."""
    _require_cuda(x, torch.float32, "x")
    outer, h, w = _outer_hw(x)
    y = torch.empty(tuple(x.shape[:-2]) + (h * scale, w * scale), dtype=torch.float32, device=x.device)
    check(lib().sad_upsample_nearest_f32(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), outer, h, w, int(scale), _stream()))
    return y


def upsample_nearest_grad(x, dy, scale=2):
    """UpsampleNearestGradient (upsample_nearest_op.cu:162-211): dx like x, each element the sum of its scale x scale block of dy."""
    _require_cuda(dy, torch.float32, "dy")
    outer, h, w = _outer_hw(x)
    if dy.numel() != x.numel() * scale * scale:
        raise ValueError("dy must be the upsampled shape of x")
    dx = torch.empty(tuple(x.shape), dtype=torch.float32, device=dy.device)
    check(lib().sad_upsample_nearest_grad_f32(C.c_void_p(dy.data_ptr()), C.c_void_p(dx.data_ptr()), outer, h, w, int(scale), _stream()))
    return dx


def upsample_nearest_add(top, lateral, out=None):
    """FPN top-down merge (FPN.py:230-249) in one pass: lateral + UpsampleNearest(top, 2).  (N, C, H, W) tensors, contiguous NCHW
    or channels-last (both operands in the same format); out may be `lateral` (in place)."""
    _require_cuda_any(top, "top")
    _require_cuda_any(lateral, "lateral")
    n, c, h, w = lateral.shape
    if tuple(top.shape) != (n, c, h // 2, w // 2) or h % 2 or w % 2:
        raise ValueError("top must be (N, C, H/2, W/2) of the lateral's (N, C, H, W)")
    cl = lateral.is_contiguous(memory_format=torch.channels_last) and not lateral.is_contiguous()
    fmt = torch.channels_last if cl else torch.contiguous_format
    if not (top.is_contiguous(memory_format=fmt) and lateral.is_contiguous(memory_format=fmt)):
        raise ValueError("top and lateral must both be contiguous NCHW or both channels-last")
    if out is None:
        out = torch.empty_like(lateral, memory_format=fmt)
    outer, inner = (n, c) if cl else (n * c, 1)
    check(lib().sad_upsample_nearest_add_f32(C.c_void_p(top.data_ptr()), C.c_void_p(lateral.data_ptr()), C.c_void_p(out.data_ptr()),
                                             outer, h, w, inner, _stream()))
    return out


def _require_cuda_any(t, name):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise TypeError("%s must be a CUDA float32 tensor" % name)


def scale_(x, alpha):
    """Scale in place (scale_op.h:31-50): x *= alpha — the momentum correction of a learning-rate change (detector.py:628-648)."""
    _require_cuda(x, torch.float32, "x")
    check(lib().sad_scale_f32(C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), x.numel(), float(alpha), _stream()))
    return x


# ---------------------------------------------------------------------------------------------
# fp16-operand forward convolution (BASELINE.json configs[4]: mixed fp16 compute, fp32 accumulate) — forward only
# ---------------------------------------------------------------------------------------------
def conv3x3_pack_f16(weight, mode=0):
    """(Cout, Cin, 3, 3) fp32 -> fp16 (round to nearest) for sad_conv3x3_fwd_f16: mode 0 [tap][Cout][Cin] (forward), mode 1
    [tap][Cin][pad8(Cout)] with flipped taps (data gradient; the K axis is zero-padded to a multiple of 8)."""
    _require_cuda(weight, torch.float32, "weight")
    cout, cin = weight.shape[0], weight.shape[1]
    if tuple(weight.shape[2:]) != (3, 3):
        raise ValueError("weight must be (Cout, Cin, 3, 3)")
    m, k = (cout, cin) if mode == 0 else (cin, cout)
    packed = torch.empty(9 * m * ((k + 7) // 8 * 8), dtype=torch.float16, device=weight.device)
    item = (native.PackItem * 1)()
    item[0].weight, item[0].packed, item[0].cin, item[0].cout, item[0].mode = weight.data_ptr(), packed.data_ptr(), cin, cout, int(mode)
    check(lib().sad_conv3x3_pack_weights_multi_f16(item, 1, _stream()))
    return packed


def to_nhwc_f16(xs, channels_dst=None, scale=1.0):
    """NCHW fp32 -> channels-last (N, H, W, channels_dst or C) fp16, every level in one launch; values are multiplied by `scale`
    before rounding (the loss scale of gradient tensors) and channels beyond C are zero."""
    xs = list(xs)
    arr = (native.LayoutLevel * len(xs))()
    outs = []
    for i, x in enumerate(xs):
        _require_cuda(x, torch.float32, "x[%d]" % i)
        if x.shape[1] != xs[0].shape[1]:
            raise ValueError("all levels must have the same channel count")
        n, c, h, w = x.shape
        outs.append(torch.empty((n, h, w, channels_dst or c), dtype=torch.float16, device=x.device))
        arr[i].src_nchw, arr[i].dst_nhwc = x.data_ptr(), outs[-1].data_ptr()
        arr[i].N, arr[i].H, arr[i].W = n, h, w
    check(lib().sad_nchw_to_nhwc_f16(arr, len(xs), xs[0].shape[1], int(channels_dst or 0), float(scale), _stream()))
    return outs


def conv3x3_forward_f16(xs_nhwc_f16, packed_f16, cout, bias=None, relu=0, want_nchw=True, want_nhwc=False, nchw_scale=1.0, relu_bits=None,
                        want_bits=False):
    """Conv (+bias, + activation: 0 none, 1 ReLU, 2 Sigmoid) of every level in one launch with fp16 operands.
    xs_nhwc_f16: channels-last fp16 inputs (to_nhwc_f16 or a previous call's fp16 output).  Returns (ys_nchw fp32, ys_nhwc fp16)."""
    arr = (ConvLevel * len(xs_nhwc_f16))()
    ys, yts, bits = [], [], []
    cin = xs_nhwc_f16[0].shape[3]
    for i, xt in enumerate(xs_nhwc_f16):
        _require_cuda(xt, torch.float16, "x_nhwc[%d]" % i)
        n, h, w, c = xt.shape
        if c != cin:
            raise ValueError("all levels must have the same channel count")
        arr[i].x_nhwc = xt.data_ptr()
        arr[i].N, arr[i].H, arr[i].W = n, h, w
        if want_nchw:
            ys.append(torch.empty((n, cout, h, w), dtype=torch.float32, device=xt.device))
            arr[i].y_nchw = ys[-1].data_ptr()
        if want_nhwc:
            yts.append(torch.empty((n, h, w, cout), dtype=torch.float16, device=xt.device))
            arr[i].y_nhwc = yts[-1].data_ptr()
        if relu_bits is not None:      # ReluGradient fused into a data-gradient pass: sign bits a forward pass left
            arr[i].relu_bits_in = relu_bits[i].data_ptr()
        if want_bits:
            bits.append(sign_bits_like(n, cout, h, w, xt.device))
            arr[i].relu_bits_out = bits[-1].data_ptr()
    _require_cuda(packed_f16, torch.float16, "packed")
    if packed_f16.numel() != 9 * cin * cout:
        raise ValueError("packed weights must hold 9 * Cin * Cout fp16 elements")
    b = C.c_void_p(bias.data_ptr()) if bias is not None else None
    check(lib().sad_conv3x3_fwd_f16(arr, len(xs_nhwc_f16), C.c_void_p(packed_f16.data_ptr()), b, cin, cout, int(relu), float(nchw_scale), _stream()))
    return (ys, yts, bits) if want_bits else (ys, yts)


def conv3x3_wgrad_f16(xs_nhwc_f16, dys_nhwc_f16, cout=None, out_scale=1.0, want_bias=True):
    """Weight (+ bias) gradient summed over every level with fp16 operands.  xs: (N, H, W, Cin) fp16; dys: (N, H, W, C_dy) fp16 with
    C_dy >= cout (channel-padded gradient tensors); the result is multiplied by out_scale (1 / loss scale).  Returns (dW, db) fp32."""
    n = len(xs_nhwc_f16)
    arr = (WgradLevel * n)()
    cin, cdy = xs_nhwc_f16[0].shape[3], dys_nhwc_f16[0].shape[3]
    cout = cdy if cout is None else int(cout)
    for i, (xt, dt) in enumerate(zip(xs_nhwc_f16, dys_nhwc_f16)):
        _require_cuda(xt, torch.float16, "x_nhwc[%d]" % i)
        _require_cuda(dt, torch.float16, "dy_nhwc[%d]" % i)
        if xt.shape[:3] != dt.shape[:3] or xt.shape[3] != cin or dt.shape[3] != cdy:
            raise ValueError("level %d: x (N,H,W,Cin) and dy (N,H,W,C_dy) do not match" % i)
        arr[i].x_nhwc, arr[i].dy_nhwc = xt.data_ptr(), dt.data_ptr()
        arr[i].N, arr[i].H, arr[i].W = xt.shape[:3]
    dev = xs_nhwc_f16[0].device
    dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=dev)
    db = torch.empty((cout,), dtype=torch.float32, device=dev) if want_bias else None
    nbytes = lib().sad_conv3x3_wgrad_workspace_bytes(arr, n, cin, cdy)
    workspace = torch.empty(max(256, nbytes), dtype=torch.uint8, device=dev)
    check(lib().sad_conv3x3_wgrad_f16(arr, n, cin, cdy, cout, float(out_scale), C.c_void_p(dw.data_ptr()),
                                      C.c_void_p(db.data_ptr()) if db is not None else None, 0,
                                      C.c_void_p(workspace.data_ptr()), workspace.numel(), _stream()))
    return dw, db


# ---------------------------------------------------------------------------------------------
# 3xTF32: the fp32-accurate convolution mode (split [hi | lo] operands; include/sad_b200.h, sad_conv3x3_fwd_f32x3)
# ---------------------------------------------------------------------------------------------
def split_channels(c):
    """Offset of the lo half of a split row: round_up(c, 32); rows are 2 * split_channels(c) floats long."""
    return lib().sad_conv3x3_split_channels(int(c))


def join_split(t_nhwc_split, channels):
    """Split channels-last tensor (N, H, W, 2 * split_channels(C)) -> the fp32 values it carries, NCHW: hi + lo."""
    cs = t_nhwc_split.shape[3] // 2
    return (t_nhwc_split[..., :channels] + t_nhwc_split[..., cs:cs + channels]).permute(0, 3, 1, 2).contiguous()


def conv3x3_pack_f32x3(weight, mode=0):
    """(Cout, Cin, 3, 3) fp32 -> [tap][M][hi(0..K) pad | lo(0..K) pad] tf32 pairs: mode 0 forward (M = Cout, K = Cin), mode 1 data
    gradient (M = Cin, K = Cout, taps flipped)."""
    _require_cuda(weight, torch.float32, "weight")
    cout, cin = weight.shape[0], weight.shape[1]
    if tuple(weight.shape[2:]) != (3, 3):
        raise ValueError("weight must be (Cout, Cin, 3, 3)")
    packed = torch.empty(lib().sad_conv3x3_packed_bytes_f32x3(cin, cout, int(mode)) // 4, dtype=torch.float32, device=weight.device)
    item = (native.PackItem * 1)()
    item[0].weight, item[0].packed, item[0].cin, item[0].cout, item[0].mode = weight.data_ptr(), packed.data_ptr(), cin, cout, int(mode)
    check(lib().sad_conv3x3_pack_weights_multi_f32x3(item, 1, _stream()))
    return packed


def to_nhwc_f32x3(xs):
    """NCHW fp32 -> split channels-last (N, H, W, 2 * split_channels(C)), every level in one launch."""
    xs = list(xs)
    arr = (native.LayoutLevel * len(xs))()
    outs = []
    for i, x in enumerate(xs):
        _require_cuda(x, torch.float32, "x[%d]" % i)
        if x.shape[1] != xs[0].shape[1]:
            raise ValueError("all levels must have the same channel count")
        n, c, h, w = x.shape
        outs.append(torch.empty((n, h, w, 2 * split_channels(c)), dtype=torch.float32, device=x.device))
        arr[i].src_nchw, arr[i].dst_nhwc = x.data_ptr(), outs[-1].data_ptr()
        arr[i].N, arr[i].H, arr[i].W = n, h, w
    check(lib().sad_nchw_to_nhwc_f32x3(arr, len(xs), xs[0].shape[1], _stream()))
    return outs


def conv3x3_forward_f32x3(xs_split, packed_x3, cin, cout, bias=None, relu=0, want_nchw=True, want_nhwc=False, relu_bits=None, want_bits=False):
    """Conv (+bias, + activation: 0 none, 1 ReLU, 2 Sigmoid) of every level in one launch in 3xTF32 arithmetic; with mode-1 packed
    weights (and cin / cout swapped) the data gradient.  xs_split: split channels-last inputs.  Returns (ys_nchw, ys_split)."""
    arr = (ConvLevel * len(xs_split))()
    ys, yts, bits = [], [], []
    for i, xt in enumerate(xs_split):
        _require_cuda(xt, torch.float32, "x_split[%d]" % i)
        n, h, w, c = xt.shape
        if c != 2 * split_channels(cin):
            raise ValueError("x_split rows must be 2 * split_channels(cin) floats")
        arr[i].x_nhwc = xt.data_ptr()
        arr[i].N, arr[i].H, arr[i].W = n, h, w
        if want_nchw:
            ys.append(torch.empty((n, cout, h, w), dtype=torch.float32, device=xt.device))
            arr[i].y_nchw = ys[-1].data_ptr()
        if want_nhwc:   # pad channels are never written by the kernel and must read as zero
            yts.append(torch.zeros((n, h, w, 2 * split_channels(cout)), dtype=torch.float32, device=xt.device))
            arr[i].y_nhwc = yts[-1].data_ptr()
        if relu_bits is not None:
            arr[i].relu_bits_in = relu_bits[i].data_ptr()
        if want_bits:
            bits.append(sign_bits_like(n, cout, h, w, xt.device))
            arr[i].relu_bits_out = bits[-1].data_ptr()
    b = C.c_void_p(bias.data_ptr()) if bias is not None else None
    check(lib().sad_conv3x3_fwd_f32x3(arr, len(xs_split), C.c_void_p(packed_x3.data_ptr()), b, int(cin), int(cout), int(relu), _stream()))
    return (ys, yts, bits) if want_bits else (ys, yts)


def conv3x3_wgrad_f32x3(xs_split, dys_split, cin, cout, want_bias=True):
    """Weight (+ bias) gradient summed over every level in 3xTF32 arithmetic from split tensors.  Returns (dW, db) fp32."""
    n = len(xs_split)
    arr = (WgradLevel * n)()
    for i, (xt, dt) in enumerate(zip(xs_split, dys_split)):
        _require_cuda(xt, torch.float32, "x_split[%d]" % i)
        _require_cuda(dt, torch.float32, "dy_split[%d]" % i)
        if xt.shape[:3] != dt.shape[:3] or xt.shape[3] != 2 * split_channels(cin) or dt.shape[3] != 2 * split_channels(cout):
            raise ValueError("level %d: split x / dy rows do not match cin / cout" % i)
        arr[i].x_nhwc, arr[i].dy_nhwc = xt.data_ptr(), dt.data_ptr()
        arr[i].N, arr[i].H, arr[i].W = xt.shape[:3]
    dev = xs_split[0].device
    dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=dev)
    db = torch.empty((cout,), dtype=torch.float32, device=dev) if want_bias else None
    nbytes = lib().sad_conv3x3_wgrad_workspace_bytes(arr, n, int(cin), int(cout))
    workspace = torch.empty(max(256, nbytes), dtype=torch.uint8, device=dev)
    check(lib().sad_conv3x3_wgrad_f32x3(arr, n, int(cin), int(cout), C.c_void_p(dw.data_ptr()),
                                        C.c_void_p(db.data_ptr()) if db is not None else None, 0,
                                        C.c_void_p(workspace.data_ptr()), workspace.numel(), _stream()))
    return dw, db
