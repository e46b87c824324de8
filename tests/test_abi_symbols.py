"""CPU: the C-ABI library loads without a GPU and exports every symbol include/*.h declares;
argument validation (which needs no device) behaves as documented."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "sad_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(sad_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ["sad_pow_sum_f32", "sad_distill_f32", "sad_workspace_init", "sad_distill_step_host",
                 "sad_pow_sum_workspace_bytes", "sad_distill_workspace_bytes", "sad_last_error"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    from sad_b200 import native
    lib = native.lib()
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, "declared in include/sad_b200.h but not exported: %s" % missing
    assert b"sm_100a" in lib.sad_version()


def test_argument_validation_without_device():
    from sad_b200 import native
    lib = native.lib()
    p = native.default_params()
    assert (p.gamma, p.alpha, p.beta, p.scale, p.num_classes, p.ignored_label) == (1.0, 0.25, 0.0, 1.0, 80, -1)
    lv = (native.DistillLevel * 1)()
    lv[0].N, lv[0].D, lv[0].H, lv[0].W = 1, 81, 2, 2      # D not a multiple of num_classes
    lv[0].logits = lv[0].teacher_prob = lv[0].labels = 256
    lv[0].loss = 256
    dummy = C.c_void_p(256)
    rc = lib.sad_distill_f32(lv, 1, dummy, C.byref(p), None, 0, None)
    assert rc == -1 and b"multiple of num_classes" in lib.sad_last_error()
    lv[0].D = 80
    p.scale = -1.0                                        # reference: CAFFE_ENFORCE(scale_ >= 0)
    rc = lib.sad_distill_f32(lv, 1, dummy, C.byref(p), None, 0, None)
    assert rc == -1 and b"scale" in lib.sad_last_error()
    p.scale = 1.0
    rc = lib.sad_distill_f32(lv, 1, dummy, C.byref(p), None, 0, None)   # loss wanted but no workspace
    assert rc == -3
    rc = lib.sad_distill_f32(lv, 0, dummy, C.byref(p), None, 0, None)
    assert rc == -1
    sizes = (C.c_int64 * 2)(100, 9000)
    assert lib.sad_pow_sum_workspace_bytes(sizes, 2) >= 256 + 3 * 4
    assert lib.sad_pow_sum_workspace_bytes(sizes, 0) == 0


def test_operator_library_exports_handle_api():
    from sad_b200 import native
    lib = C.CDLL(native.OPS_LIB_PATH)
    for name in ["c2_workspace_create", "c2_feed_external", "c2_run_operator_once", "c2_create_net", "c2_run_net",
                 "c2_fetch", "c2_gradient_defs", "c2_has_operator", "c2_fuse_adaptive_distill_ops"]:
        assert hasattr(lib, name), name


def test_new_entry_points_validate_arguments_without_device():
    """AffineChannel / UpsampleNearest / Scale / fp16 convolution: bad arguments are refused before any CUDA call."""
    from sad_b200 import native
    lib = native.lib()
    p = C.c_void_p(256)
    assert lib.sad_affine_channel_f32(p, p, p, p, 1, 0, 4, None) == -1 and b"bad shape" in lib.sad_last_error()
    assert lib.sad_affine_channel_f32(None, p, p, p, 1, 2, 4, None) == -1 and b"null" in lib.sad_last_error()
    assert lib.sad_affine_channel_f32(p, p, p, p, 0, 2, 4, None) == 0                      # empty tensor: nothing to do
    assert lib.sad_affine_channel_f32(p, p, p, p, 1 << 15, 1 << 10, 1 << 10, None) == -4   # 2^35 elements: int indexing refused
    assert lib.sad_upsample_nearest_f32(p, p, 1, 2, 2, 0, None) == -1 and b"scale" in lib.sad_last_error()
    assert lib.sad_upsample_nearest_f32(p, p, 0, 2, 2, 2, None) == 0
    assert lib.sad_upsample_nearest_grad_f32(None, p, 1, 2, 2, 2, None) == -1
    assert lib.sad_scale_f32(None, p, 4, 1.0, None) == -1 and lib.sad_scale_f32(p, p, 0, 1.0, None) == 0
    lv = (native.ConvLevel * 1)()
    assert lib.sad_conv3x3_fwd_f16(lv, 0, p, None, 64, 64, 0, 1.0, None) == -1
    assert lib.sad_conv3x3_fwd_f16(lv, 1, None, None, 64, 64, 0, 1.0, None) == -1
    lv[0].N, lv[0].H, lv[0].W = 1, 4, 4
    lv[0].x_nhwc, lv[0].y_nchw = 256, 256
    assert lib.sad_conv3x3_fwd_f16(lv, 1, p, None, 36, 64, 0, 1.0, None) == -4 and b"Cin % 8" in lib.sad_last_error()
    lv[0].relu_mask_nhwc = 256
    assert lib.sad_conv3x3_fwd_f16(lv, 1, p, None, 64, 64, 0, 1.0, None) == -4 and b"sign bits" in lib.sad_last_error()
    wl = (native.WgradLevel * 1)()
    wl[0].N, wl[0].H, wl[0].W = 1, 4, 4
    wl[0].x_nhwc = wl[0].dy_nhwc = 256
    assert lib.sad_conv3x3_wgrad_f16(wl, 1, 64, 40, 44, 1.0, p, None, 0, p, 1 << 30, None) in (-1, -2)   # cout > dY channels (-2: no device)


def test_exchange_library_exports_every_declared_symbol():
    """include/sad_exchange.h (the gradient exchange: host C++ over NCCL).  Loads without a GPU and without NCCL (resolved at
    run time); argument validation needs neither."""
    from sad_b200 import exchange
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "sad_exchange.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(sad_exchange_[a-z0-9_]+)\s*\(", src)))
    assert "sad_exchange_allreduce_async_f32" in names and "sad_exchange_join" in names and len(names) >= 11
    lib = exchange.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    out = C.c_void_p()
    assert lib.sad_exchange_create(None, 2, 2, C.byref(out)) == -1 and b"rank" in lib.sad_exchange_last_error()
    assert lib.sad_exchange_allreduce_async_f32(None, None, 0, None) == -1
    assert lib.sad_exchange_join(None, None) == -1
    assert lib.sad_exchange_world(None) == 0


def test_exchange_fails_loudly_without_a_device_and_validates_the_copy_engine_form():
    """No CPU path behind the exchange either: creating one where there is no CUDA device reports the CUDA error; the copy-engine
    form validates its capacity; its capability query and the empty slot sum need neither a device nor NCCL."""
    import torch
    from sad_b200 import exchange
    lib = exchange.lib()
    out = C.c_void_p()
    assert lib.sad_exchange_create_gather(None, 0, 1, 0, C.byref(out)) == -1 and b"capacity" in lib.sad_exchange_last_error()
    assert lib.sad_exchange_gather_supported() in (0, 1)
    assert lib.sad_exchange_slot_sum_f32(None, 0, 2, None, 0, None) == 0          # nothing to add
    assert lib.sad_exchange_slot_sum_f32(None, 128, 2, None, 4, None) != 0         # null slots: cudaErrorInvalidValue, not a crash
    assert lib.sad_exchange_gather_capacity(None) == 0 and lib.sad_exchange_gathered(None) == 0
    if not torch.cuda.is_available():
        rc = lib.sad_exchange_create(None, 0, 1, C.byref(out))
        assert rc == -2 and not out.value, (rc, lib.sad_exchange_last_error())       # SAD_EXCHANGE_ERR_CUDA
        assert lib.sad_exchange_last_error()
