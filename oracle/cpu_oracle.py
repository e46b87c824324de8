"""TEST INFRASTRUCTURE — numpy-facing loader of the CPU oracle (oracle/liboracle_cpu.so).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_cpu.so")
REF_GPU_LIB = os.path.join(HERE, "_ref", "libref_ops.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(verbose=False):
    """Compile the C restatement (and, when /root/reference exists, the GPU oracle from the
    unmodified reference sources).  Building the checker is not using it."""
    r = subprocess.run(["make", "-C", HERE, "all"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode:
        print(r.stdout)
    if r.returncode:
        raise RuntimeError("oracle build failed")


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        l.oracle_num_threads.restype = C.c_int
        l.oracle_set_num_threads.argtypes = [C.c_int]
        l.oracle_ref_order_sum.restype = C.c_float
        l.oracle_ref_order_sum.argtypes = [_f32p, C.c_int64]
        l.oracle_pow_sum.restype = C.c_float
        l.oracle_pow_sum.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_float, _f32p]
        l.oracle_label_index.restype = C.c_int
        l.oracle_label_index.argtypes = [C.c_int] * 5
        l.oracle_distill_loss_elem.restype = C.c_float
        l.oracle_distill_loss_elem.argtypes = [C.c_float, C.c_float, C.c_int] + [C.c_float] * 4
        l.oracle_distill_grad_elem.restype = C.c_float
        l.oracle_distill_grad_elem.argtypes = [C.c_float, C.c_float, C.c_int] + [C.c_float] * 6
        l.oracle_distill_loss.restype = C.c_float
        l.oracle_distill_loss.argtypes = [C.c_int] * 5 + [_f32p, _f32p, _i32p, _f32p] + [C.c_float] * 3 + [C.c_int, C.c_float, _f32p]
        l.oracle_distill_grad.restype = None
        l.oracle_distill_grad.argtypes = [C.c_int] * 5 + [_f32p, _f32p, _i32p, _f32p] + [C.c_float] * 3 + [C.c_int, C.c_float, _f32p, _f32p]
        l.oracle_distill_elem_f64.restype = None
        l.oracle_distill_elem_f64.argtypes = [C.c_double, C.c_double, C.c_int] + [C.c_double] * 6 + [C.POINTER(C.c_double)] * 2
        l.oracle_conv_out_dim.restype = C.c_int
        l.oracle_conv_out_dim.argtypes = [C.c_int] * 4
        l.oracle_conv2d_fwd.restype = None
        l.oracle_conv2d_fwd.argtypes = [_f32p, _f32p, C.c_void_p, _f32p] + [C.c_int] * 9
        l.oracle_conv2d_bwd.restype = None
        l.oracle_conv2d_bwd.argtypes = [_f32p, _f32p, _f32p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 9
        l.oracle_relu.argtypes = [_f32p, _f32p, C.c_int64]
        l.oracle_relu_grad.argtypes = [_f32p, _f32p, _f32p, C.c_int64]
        l.oracle_focal_loss.restype = C.c_float
        l.oracle_focal_loss.argtypes = [C.c_int] * 4 + [_f32p, _i32p, _f32p, C.c_float, C.c_float, C.c_int, C.c_float, _f32p]
        l.oracle_select_smooth_l1_loss.restype = C.c_float
        l.oracle_select_smooth_l1_loss.argtypes = [C.c_int] * 5 + [_f32p] * 4 + [C.c_float, C.c_float, _f32p]
        l.oracle_select_smooth_l1_grad.restype = None
        l.oracle_select_smooth_l1_grad.argtypes = [C.c_int] * 5 + [_f32p] * 4 + [C.c_float, C.c_float, _f32p, _f32p]
        l.oracle_upsample2_add.restype = None
        l.oracle_upsample2_add.argtypes = [_f32p, _f32p, _f32p, C.c_int64, C.c_int, C.c_int, C.c_int64]
        l.oracle_weighted_sum.restype = None
        l.oracle_weighted_sum.argtypes = [C.c_int64, C.c_int, C.POINTER(C.c_void_p), _f32p, _f32p]
        l.oracle_momentum_sgd.restype = None
        l.oracle_momentum_sgd.argtypes = [C.c_int64, _f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float]
        l.oracle_affine_channel.restype = None
        l.oracle_affine_channel.argtypes = [C.c_int64, C.c_int, C.c_int64, _f32p, _f32p, C.c_void_p, _f32p]
        l.oracle_upsample_nearest.restype = None
        l.oracle_upsample_nearest.argtypes = [_f32p, _f32p, C.c_int64] + [C.c_int] * 4
        l.oracle_upsample_nearest_grad.restype = None
        l.oracle_upsample_nearest_grad.argtypes = [_f32p, _f32p, C.c_int64] + [C.c_int] * 4
        l.oracle_focal_grad.restype = None
        l.oracle_focal_grad.argtypes = [C.c_int] * 4 + [_f32p, _i32p, _f32p, C.c_float, C.c_float, C.c_int, C.c_float, _f32p, _f32p]
        _lib = l
    return _lib


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def pow_sum(inputs, power):
    inputs = [np.ascontiguousarray(x, dtype=np.float32).reshape(-1) for x in inputs]
    n = len(inputs)
    ptrs = (C.c_void_p * n)(*[x.ctypes.data for x in inputs])
    sizes = (C.c_int64 * n)(*[x.size for x in inputs])
    scratch = np.empty(max(1, max(x.size for x in inputs)), dtype=np.float32)
    return np.float32(lib().oracle_pow_sum(ptrs, sizes, n, float(power), scratch))


def distill_loss(logits, teacher_prob, labels, normalizer, gamma=1.0, alpha=0.25, beta=0.0, scale=1.0,
                 num_classes=80, ignored_label=-1, return_elements=False):
    x = np.ascontiguousarray(logits, dtype=np.float32)
    t = np.ascontiguousarray(teacher_prob, dtype=np.float32)
    g = np.ascontiguousarray(labels, dtype=np.int32)
    wp = np.asarray([normalizer], dtype=np.float32).reshape(1)
    N, D, H, W = x.shape
    losses = np.empty(max(1, x.size), dtype=np.float32)
    val = lib().oracle_distill_loss(N, D, H, W, int(ignored_label), x.reshape(-1), t.reshape(-1), g.reshape(-1), wp,
                                    float(gamma), float(alpha), float(beta), int(num_classes), float(scale), losses)
    if return_elements:
        return np.float32(val), losses[:x.size].reshape(x.shape)
    return np.float32(val)


def distill_grad(logits, teacher_prob, labels, normalizer, d_loss=1.0, gamma=1.0, alpha=0.25, beta=0.0, scale=1.0,
                 num_classes=80, ignored_label=-1):
    x = np.ascontiguousarray(logits, dtype=np.float32)
    t = np.ascontiguousarray(teacher_prob, dtype=np.float32)
    g = np.ascontiguousarray(labels, dtype=np.int32)
    wp = np.asarray([normalizer], dtype=np.float32).reshape(1)
    dl = np.asarray([d_loss], dtype=np.float32).reshape(1)
    N, D, H, W = x.shape
    dX = np.empty(max(1, x.size), dtype=np.float32)
    lib().oracle_distill_grad(N, D, H, W, int(ignored_label), x.reshape(-1), t.reshape(-1), g.reshape(-1), wp,
                              float(gamma), float(alpha), float(beta), int(num_classes), float(scale), dl, dX)
    return dX[:x.size].reshape(x.shape)


def focal_loss(logits, labels, fg_num, gamma=1.0, alpha=0.25, scale=1.0, num_classes=80, return_elements=False):
    """SigmoidFocalLoss (sigmoid_focal_loss_op.cu:26-66,112-144): scalar loss in the reference's summation order."""
    x = np.ascontiguousarray(logits, dtype=np.float32)
    g = np.ascontiguousarray(labels, dtype=np.int32)
    wp = np.asarray([fg_num], dtype=np.float32).reshape(1)
    N, D, H, W = x.shape
    losses = np.empty(max(1, x.size), dtype=np.float32)
    val = lib().oracle_focal_loss(N, D, H, W, x.reshape(-1), g.reshape(-1), wp, float(gamma), float(alpha), int(num_classes),
                                  float(scale), losses)
    if return_elements:
        return np.float32(val), losses[:x.size].reshape(x.shape)
    return np.float32(val)


def focal_grad(logits, labels, fg_num, d_loss=1.0, gamma=1.0, alpha=0.25, scale=1.0, num_classes=80):
    """SigmoidFocalLossGradient (sigmoid_focal_loss_op.cu:68-109,147-173)."""
    x = np.ascontiguousarray(logits, dtype=np.float32)
    g = np.ascontiguousarray(labels, dtype=np.int32)
    wp = np.asarray([fg_num], dtype=np.float32).reshape(1)
    dl = np.asarray([d_loss], dtype=np.float32).reshape(1)
    N, D, H, W = x.shape
    dX = np.empty(max(1, x.size), dtype=np.float32)
    lib().oracle_focal_grad(N, D, H, W, x.reshape(-1), g.reshape(-1), wp, float(gamma), float(alpha), int(num_classes),
                            float(scale), dl, dX)
    return dX[:x.size].reshape(x.shape)


def select_smooth_l1(y_hat, y, locs, fg_num, beta=1.0, scale=1.0, d_loss=1.0):
    """(loss, d_y_hat) of SelectSmoothL1Loss / Gradient (select_smooth_l1_loss_op.cu:23-86, 90-181)."""
    yh = np.ascontiguousarray(y_hat, dtype=np.float32)
    N, D, H, W = yh.shape
    y = np.ascontiguousarray(y, dtype=np.float32).reshape(-1, 4)
    locs = np.ascontiguousarray(locs, dtype=np.float32).reshape(-1, 4)
    M = y.shape[0]
    S = np.asarray([fg_num], dtype=np.float32)
    dl = np.asarray([d_loss], dtype=np.float32)
    buff = np.empty(max(1, yh.size), dtype=np.float32)
    grad = np.empty(max(1, yh.size), dtype=np.float32)
    yy = y.reshape(-1) if M else np.zeros(4, np.float32)
    ll = locs.reshape(-1) if M else np.zeros(4, np.float32)
    loss = lib().oracle_select_smooth_l1_loss(N, D, H, W, M, yh.reshape(-1), yy, ll, S, float(beta), float(scale), buff)
    lib().oracle_select_smooth_l1_grad(N, D, H, W, M, yh.reshape(-1), yy, ll, S, float(beta), float(scale), dl, grad)
    return np.float32(loss), grad[:yh.size].reshape(yh.shape)


def momentum_sgd(param, grad, mom, lr, momentum=0.9, nesterov=False, grad_mult=1.0, wd=0.0):
    """Returns updated copies (param, grad, momentum) — optimizer.py:115-130 + momentum_sgd_op_gpu.cu:23-54."""
    p, g, m = (np.array(a, dtype=np.float32, copy=True).reshape(-1) for a in (param, grad, mom))
    lib().oracle_momentum_sgd(p.size, p, g, m, float(lr), float(momentum), 1 if nesterov else 0, float(grad_mult), float(wd))
    return p, g, m


def distill_elem_f64(x, pt, keep=1, wp=1.0, gamma=2.0, alpha=0.5, beta=0.0, d_loss=1.0, scale=1.0):
    lo, gr = C.c_double(), C.c_double()
    lib().oracle_distill_elem_f64(x, pt, int(keep), wp, gamma, alpha, beta, d_loss, scale, C.byref(lo), C.byref(gr))
    return lo.value, gr.value


def label_index(i, D, H, W, num_classes):
    return lib().oracle_label_index(int(i), D, H, W, num_classes)


def conv2d_fwd(X, Wt, b=None, pad=1, stride=1):
    X = np.ascontiguousarray(X, dtype=np.float32)
    Wt = np.ascontiguousarray(Wt, dtype=np.float32)
    N, Cc, H, W = X.shape
    M, _, kh, kw = Wt.shape
    Ho, Wo = lib().oracle_conv_out_dim(H, kh, pad, stride), lib().oracle_conv_out_dim(W, kw, pad, stride)
    Y = np.empty((N, M, Ho, Wo), dtype=np.float32)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
    lib().oracle_conv2d_fwd(X, Wt, None if bb is None else bb.ctypes.data, Y, N, Cc, H, W, M, kh, kw, pad, stride)
    return Y


def conv2d_bwd(X, Wt, dY, pad=1, stride=1, need_dx=True):
    X = np.ascontiguousarray(X, dtype=np.float32)
    Wt = np.ascontiguousarray(Wt, dtype=np.float32)
    dY = np.ascontiguousarray(dY, dtype=np.float32)
    N, Cc, H, W = X.shape
    M, _, kh, kw = Wt.shape
    dW = np.empty_like(Wt)
    db = np.empty(M, dtype=np.float32)
    dX = np.empty_like(X) if need_dx else None
    lib().oracle_conv2d_bwd(X, Wt, dY, dW.ctypes.data, db.ctypes.data, None if dX is None else dX.ctypes.data,
                            N, Cc, H, W, M, kh, kw, pad, stride)
    return dW, db, dX


def relu(X):
    X = np.ascontiguousarray(X, dtype=np.float32)
    Y = np.empty_like(X)
    lib().oracle_relu(X.reshape(-1), Y.reshape(-1), X.size)
    return Y


def relu_grad(Y, dY):
    Y = np.ascontiguousarray(Y, dtype=np.float32)
    dY = np.ascontiguousarray(dY, dtype=np.float32)
    dX = np.empty_like(Y)
    lib().oracle_relu_grad(Y.reshape(-1), dY.reshape(-1), dX.reshape(-1), Y.size)
    return dX


def affine_channel(x, scale, bias=None):
    """AffineChannel (bias given) / AffineChannelGradient (bias None) — affine_channel_op.cu:22-48."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    scale = np.ascontiguousarray(scale, dtype=np.float32)
    N, Cc, H, W = x.shape
    out = np.empty_like(x)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    lib().oracle_affine_channel(x.size, Cc, H * W, x.reshape(-1), scale, None if b is None else b.ctypes.data, out.reshape(-1))
    return out


def _d123(shape):
    """(d1, d2, d3) as upsample_nearest_op.cu:129-138 takes them from a 3-D or 4-D shape."""
    return tuple(shape[-3:])


def upsample_nearest(x, scale=2):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape[:-2] + (x.shape[-2] * scale, x.shape[-1] * scale), dtype=np.float32)
    d1, d2, d3 = _d123(out.shape)
    lib().oracle_upsample_nearest(x.reshape(-1), out.reshape(-1), out.size, scale, d1, d2, d3)
    return out


def upsample_nearest_grad(x_shape, dy, scale=2):
    dy = np.ascontiguousarray(dy, dtype=np.float32)
    dx = np.empty(tuple(x_shape), dtype=np.float32)
    d1, d2, d3 = _d123(dx.shape)
    lib().oracle_upsample_nearest_grad(dx.reshape(-1), dy.reshape(-1), dx.size, scale, d1, d2, d3)
    return dx


def weighted_sum(xs, ws):
    """WeightedSum (utility_ops.h:333-378): xs[0] * ws[0] + xs[1] * ws[1] + ... with the GPU's one-FMA-per-term rounding."""
    xs = [np.ascontiguousarray(x, dtype=np.float32).reshape(-1) for x in xs]
    w = np.ascontiguousarray(ws, dtype=np.float32)
    out = np.empty_like(xs[0])
    ptrs = (C.c_void_p * len(xs))(*[x.ctypes.data for x in xs])
    lib().oracle_weighted_sum(xs[0].size, len(xs), ptrs, w, out)
    return out


def upsample2_add(top, lateral, inner=None):
    """lateral + UpsampleNearest(top, 2).  inner=None: arrays (..., H, W) (NCHW planes); inner=C: channels-last arrays (N, H, W, C)."""
    top = np.ascontiguousarray(top, dtype=np.float32)
    lateral = np.ascontiguousarray(lateral, dtype=np.float32)
    out = np.empty_like(lateral)
    if inner is None:
        Ho, Wo = lateral.shape[-2:]
        outer, inner = lateral.size // (Ho * Wo), 1
    else:
        outer, Ho, Wo = lateral.shape[0], lateral.shape[1], lateral.shape[2]
        assert lateral.shape[3] == inner
    lib().oracle_upsample2_add(top.reshape(-1), lateral.reshape(-1), out.reshape(-1), outer, Ho, Wo, inner)
    return out
