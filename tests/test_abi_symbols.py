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
