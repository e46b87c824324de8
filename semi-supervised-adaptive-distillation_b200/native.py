"""ctypes binding of the C ABI declared in include/sad_b200.h.

There is no fallback: if libsad_b200.so is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsad_b200.so")
OPS_LIB_PATH = os.path.join(HERE, "libcaffe2_detectron_ops_gpu.so")

SAD_MAX_LEVELS = 8
SAD_MAX_INPUTS = 16


class SadError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sad_b200 error %d: %s" % (code, msg))
        self.code = code


class DistillLevel(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("teacher_prob", C.c_void_p), ("labels", C.c_void_p),
                ("d_logits", C.c_void_p), ("loss", C.c_void_p), ("d_loss", C.c_void_p),
                ("N", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


class DistillParams(C.Structure):
    _fields_ = [("gamma", C.c_float), ("alpha", C.c_float), ("beta", C.c_float), ("scale", C.c_float),
                ("num_classes", C.c_int32), ("ignored_label", C.c_int32)]


class FocalParams(C.Structure):
    _fields_ = [("gamma", C.c_float), ("alpha", C.c_float), ("scale", C.c_float), ("num_classes", C.c_int32)]


class SgdSegment(C.Structure):
    _fields_ = [("count", C.c_int64), ("grad_multiplier", C.c_float), ("weight_decay", C.c_float)]


class ConvLevel(C.Structure):
    _fields_ = [("x_nhwc", C.c_void_p), ("y_nchw", C.c_void_p), ("y_nhwc", C.c_void_p),
                ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("relu_mask_nhwc", C.c_void_p),
                ("accumulate_nchw", C.c_int32), ("relu_bits_out", C.c_void_p), ("relu_bits_in", C.c_void_p)]


class LayoutLevel(C.Structure):
    _fields_ = [("src_nchw", C.c_void_p), ("dst_nhwc", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


class WgradLevel(C.Structure):
    _fields_ = [("x_nhwc", C.c_void_p), ("dy_nhwc", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


SAD_HEAD_MAX_CONVS = 8


class HeadConfig(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("N", C.c_int32), ("H", C.c_int32 * SAD_MAX_LEVELS), ("W", C.c_int32 * SAD_MAX_LEVELS),
                ("dim", C.c_int32), ("num_convs", C.c_int32), ("cls_out", C.c_int32), ("bbox_out", C.c_int32),
                ("cls_output_sigmoid", C.c_int32), ("compute_f16", C.c_int32), ("f16_grad_scale", C.c_float),
                ("compute_f32x3", C.c_int32)]


class PackItem(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("packed", C.c_void_p), ("cin", C.c_int32), ("cout", C.c_int32), ("mode", C.c_int32)]


class HeadTensors(C.Structure):
    """sad_head_weights / sad_head_grads (same layout: const-ness differs only in C)."""
    _fields_ = [("cls_tower_w", C.c_void_p * SAD_HEAD_MAX_CONVS), ("cls_tower_b", C.c_void_p * SAD_HEAD_MAX_CONVS),
                ("bbox_tower_w", C.c_void_p * SAD_HEAD_MAX_CONVS), ("bbox_tower_b", C.c_void_p * SAD_HEAD_MAX_CONVS),
                ("cls_pred_w", C.c_void_p), ("cls_pred_b", C.c_void_p), ("bbox_pred_w", C.c_void_p), ("bbox_pred_b", C.c_void_p)]


class HostLevel(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("teacher_prob", C.c_void_p), ("labels", C.c_void_p),
                ("d_logits", C.c_void_p),
                ("N", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


_lib = None


def lib():
    """The loaded libsad_b200.so (built in-tree by build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libsad_b200.so is missing (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or semi-supervised-adaptive-distillation_b200/build.py; there is no non-CUDA fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.sad_last_error.restype = C.c_char_p
        l.sad_version.restype = C.c_char_p
        l.sad_launch_count.restype = C.c_uint64
        l.sad_workspace_init.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        l.sad_pow_sum_workspace_bytes.restype = C.c_size_t
        l.sad_pow_sum_workspace_bytes.argtypes = [C.POINTER(C.c_int64), C.c_int]
        l.sad_pow_sum_f32.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_float, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]
        l.sad_distill_default_params.argtypes = [C.POINTER(DistillParams)]
        l.sad_distill_default_params.restype = None
        l.sad_distill_workspace_bytes.restype = C.c_size_t
        l.sad_distill_workspace_bytes.argtypes = [C.POINTER(DistillLevel), C.c_int]
        l.sad_distill_f32.argtypes = [C.POINTER(DistillLevel), C.c_int, C.c_void_p, C.POINTER(DistillParams),
                                      C.c_void_p, C.c_size_t, C.c_void_p]
        l.sad_distill_fused_workspace_bytes.restype = C.c_size_t
        l.sad_distill_fused_workspace_bytes.argtypes = [C.POINTER(DistillLevel), C.c_int, C.c_int]
        l.sad_distill_fused_f32.argtypes = [C.POINTER(DistillLevel), C.c_int, C.c_float, C.c_void_p, C.POINTER(DistillParams),
                                            C.c_void_p, C.c_size_t, C.c_void_p]
        l.sad_focal_default_params.argtypes = [C.POINTER(FocalParams)]
        l.sad_focal_default_params.restype = None
        l.sad_focal_workspace_bytes.restype = C.c_size_t
        l.sad_sigmoid_focal_loss_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.POINTER(FocalParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                 C.c_size_t, C.c_void_p]
        l.sad_smooth_l1_workspace_bytes.restype = C.c_size_t
        l.sad_select_smooth_l1_loss_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                    C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.sad_momentum_sgd_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SgdSegment), C.c_int, C.c_void_p, C.c_float, C.c_int,
                                           C.c_void_p]
        l.sad_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        l.sad_ctx_destroy.argtypes = [C.c_void_p]
        l.sad_ctx_destroy.restype = None
        l.sad_distill_step_host.argtypes = [C.c_void_p, C.POINTER(HostLevel), C.c_int, C.c_float,
                                            C.POINTER(DistillParams), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        l.sad_ctx_set_host_chunk_bytes.argtypes = [C.c_void_p, C.c_size_t]
        l.sad_ctx_device_d_logits.restype = C.c_void_p
        l.sad_ctx_device_d_logits.argtypes = [C.c_void_p, C.c_int]
        l.sad_nchw_to_nhwc_f32.argtypes = [C.POINTER(LayoutLevel), C.c_int, C.c_int, C.c_void_p]
        l.sad_conv3x3_packed_bytes.restype = C.c_size_t
        l.sad_conv3x3_packed_bytes.argtypes = [C.c_int, C.c_int]
        l.sad_conv3x3_pack_weights_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.sad_conv3x3_fwd_f32.argtypes = [C.POINTER(ConvLevel), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p]
        l.sad_conv3x3_wgrad_workspace_bytes.restype = C.c_size_t
        l.sad_conv3x3_wgrad_workspace_bytes.argtypes = [C.POINTER(WgradLevel), C.c_int, C.c_int, C.c_int]
        l.sad_conv3x3_wgrad_f32.argtypes = [C.POINTER(WgradLevel), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_size_t, C.c_void_p]
        l.sad_head_default_config.argtypes = [C.POINTER(HeadConfig)]
        l.sad_head_default_config.restype = None
        l.sad_head_create.argtypes = [C.POINTER(HeadConfig), C.POINTER(C.c_void_p)]
        l.sad_head_destroy.argtypes = [C.c_void_p]
        l.sad_head_destroy.restype = None
        l.sad_head_device_bytes.argtypes = [C.c_void_p]
        l.sad_head_device_bytes.restype = C.c_size_t
        l.sad_head_forward.argtypes = [C.c_void_p, C.POINTER(HeadTensors), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
        l.sad_head_backward.argtypes = [C.c_void_p, C.POINTER(HeadTensors), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(HeadTensors), C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
        l.sad_head_copy_activation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.sad_relu_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.sad_sigmoid_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.sad_relu_grad_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.sad_nchw_to_nhwc_f16.argtypes = [C.POINTER(LayoutLevel), C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
        l.sad_conv3x3_pack_weights_multi_f16.argtypes = [C.POINTER(PackItem), C.c_int, C.c_void_p]
        l.sad_conv3x3_fwd_f16.argtypes = [C.POINTER(ConvLevel), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
        l.sad_conv3x3_wgrad_f16.argtypes = [C.POINTER(WgradLevel), C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_size_t, C.c_void_p]
        l.sad_conv3x3_split_channels.argtypes = [C.c_int]
        l.sad_conv3x3_packed_bytes_f32x3.restype = C.c_size_t
        l.sad_conv3x3_packed_bytes_f32x3.argtypes = [C.c_int, C.c_int, C.c_int]
        l.sad_nchw_to_nhwc_f32x3.argtypes = [C.POINTER(LayoutLevel), C.c_int, C.c_int, C.c_void_p]
        l.sad_conv3x3_pack_weights_multi_f32x3.argtypes = [C.POINTER(PackItem), C.c_int, C.c_void_p]
        l.sad_conv3x3_fwd_f32x3.argtypes = [C.POINTER(ConvLevel), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        l.sad_conv3x3_wgrad_f32x3.argtypes = [C.POINTER(WgradLevel), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_size_t, C.c_void_p]
        l.sad_scale_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
        l.sad_affine_channel_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]
        l.sad_upsample_nearest_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
        l.sad_upsample_nearest_grad_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
        l.sad_nonfinite_flag_f32.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        l.sad_momentum_sgd_guarded_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SgdSegment), C.c_int, C.c_void_p, C.c_float, C.c_int,
                                                   C.c_void_p, C.c_void_p]
        l.sad_head_set_f16_grad_scale.argtypes = [C.c_void_p, C.c_float]
        l.sad_head_f16_grad_scale.argtypes = [C.c_void_p]
        l.sad_head_f16_grad_scale.restype = C.c_float
        l.sad_upsample_nearest_add_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_void_p]
        l.sad_momentum_sgd_update_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                  C.c_float, C.c_int, C.c_void_p]
        l.sad_weighted_sum_f32.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        l.sad_conv3x3_sign_bits_bytes.restype = C.c_size_t
        l.sad_conv3x3_sign_bits_bytes.argtypes = [C.c_int] * 4
        l.sad_conv3x3_pack_weights_multi_f32.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise SadError(rc, lib().sad_last_error().decode())


def default_params(**kw):
    p = DistillParams()
    lib().sad_distill_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown distillation argument %r" % k)
        setattr(p, k, v)
    return p
