"""CPU: oracle/conv_oracle.c pinned to the reference's OWN CPU convolution operators.

Two anchors:
  * tests/golden/ref_cpu_conv_kat.npz — outputs of the unmodified reference `Conv` / `ConvGradient` CPU operators
    (conv_op.cc, conv_gradient_op.cc, conv_op_impl.h:31-180,357-560 compiled from /root/reference into
    oracle/_ref/libref_ops.so) on seeded inputs; made by tests/golden/make_ref_cpu_conv_golden.py;
  * the same operators run live from oracle/_ref/libref_ops.so when that library is present (this container and the
    GPU box: it travels with the snapshot), on further shapes.
Gate: 2e-6 * max|ref| (both sides are fp32 dot products over K <= 2304 terms; measured difference: 0).
"""
import os
import sys

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import make_ref_cpu_conv_golden as gen  # noqa: E402

KAT = np.load(os.path.join(GOLDEN, "ref_cpu_conv_kat.npz"))
RTOL = 2e-6


def _close(got, ref, what):
    m = float(np.abs(ref).max()) or 1.0
    d = float(np.abs(np.asarray(got, np.float64) - ref).max())
    assert d <= RTOL * m, "%s: max|d| %.3g vs %.3g * %.3g" % (what, d, RTOL, m)


@pytest.mark.parametrize("case", gen.CASES, ids=[c[0] for c in gen.CASES])
def test_conv_oracle_matches_reference_cpu_operator_golden(oracle, case):
    name = case[0]
    x, w, b, dy = gen.make_inputs(*case)
    y = oracle.conv2d_fwd(x, w, b)
    dw, db, dx = oracle.conv2d_bwd(x, w, dy, need_dx=True)
    _close(gen.sample(y), KAT[name + "_y"], name + " Y")
    _close(gen.sample(dw), KAT[name + "_dw"], name + " dW")
    _close(gen.sample(dx), KAT[name + "_dx"], name + " dX")
    if b is not None:
        _close(gen.sample(db), KAT[name + "_db"], name + " db")


def _reflib():
    from oracle import cpu_oracle
    from sad_b200 import c2
    if not os.path.exists(cpu_oracle.REF_GPU_LIB):
        pytest.skip("oracle/_ref/libref_ops.so not built (needs /root/reference)")
    return c2.OperatorLibrary(cpu_oracle.REF_GPU_LIB)


def test_reference_library_registers_the_cpu_convolution():
    from sad_b200 import c2
    lib = _reflib()
    for name in ("Conv", "ConvGradient"):
        assert lib.HasOperator(name, c2.CPU), name
    assert lib.SchemaArity("Conv") == (2, 3, 1, 1)            # conv_op.cc:72-76
    assert lib.SchemaArity("ConvGradient") == (2, 3, 1, 3)    # conv_gradient_op.cc:24


@pytest.mark.parametrize("shape", [(1, 3, 7, 5, 6), (2, 32, 48, 10, 16), (1, 64, 36, 8, 14)],
                         ids=lambda s: "x".join(map(str, s)))
def test_conv_oracle_matches_reference_cpu_operator_live(oracle, shape):
    n, cin, cout, h, w_ = shape
    lib = _reflib()
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal((n, cin, h, w_)).astype(np.float32)
    w = rng.standard_normal((cout, cin, 3, 3)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    dy = rng.standard_normal((n, cout, h, w_)).astype(np.float32)
    ref = gen.run_reference(lib, x, w, b, dy)
    _close(oracle.conv2d_fwd(x, w, b), ref["y"], "Y")
    dw, db, dx = oracle.conv2d_bwd(x, w, dy, need_dx=True)
    _close(dw, ref["dw"], "dW")
    _close(db, ref["db"], "db")
    _close(dx, ref["dx"], "dX")


# The body's convolution geometries (detectron/lib/modeling/ResNet.py:88-129,157-278, FPN.py:94-249): 1x1
# (bottleneck reduce / expand, FPN laterals), 1x1 stride 2 (STRIDE_1X1 bottlenecks and projection shortcuts), 3x3 stride 2 (ResNeXt
# bottlenecks, FPN P6 / P7), 7x7 stride 2 pad 3 (the stem).  The oracle these pin is what a native body convolution is tested against.
BODY_GEOMETRIES = [
    # name, (n, cin, cout, h, w), kernel, pad, stride, bias
    ("1x1", (2, 64, 48, 9, 14), 1, 0, 1, False),
    ("1x1_lateral_bias", (1, 96, 32, 7, 12), 1, 0, 1, True),
    ("1x1_s2", (2, 32, 64, 10, 15), 1, 0, 2, False),
    ("3x3_s2", (1, 24, 40, 11, 16), 3, 1, 2, True),
    ("7x7_s2_stem", (1, 3, 16, 23, 30), 7, 3, 2, False),
]


@pytest.mark.parametrize("name,shape,kernel,pad,stride,bias", BODY_GEOMETRIES, ids=[g[0] for g in BODY_GEOMETRIES])
def test_conv_oracle_matches_reference_cpu_operator_on_body_geometries(oracle, name, shape, kernel, pad, stride, bias):
    n, cin, cout, h, w_ = shape
    lib = _reflib()
    rng = np.random.default_rng(sum(shape) + 31 * kernel + stride)
    x = rng.standard_normal((n, cin, h, w_)).astype(np.float32)
    w = rng.standard_normal((cout, cin, kernel, kernel)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32) if bias else None
    ho, wo = (h + 2 * pad - kernel) // stride + 1, (w_ + 2 * pad - kernel) // stride + 1
    dy = rng.standard_normal((n, cout, ho, wo)).astype(np.float32)
    ref = gen.run_reference(lib, x, w, b, dy, kernel=kernel, pad=pad, stride=stride)
    assert ref["y"].shape == (n, cout, ho, wo)
    _close(oracle.conv2d_fwd(x, w, b, pad=pad, stride=stride), ref["y"], name + " Y")
    dw, db, dx = oracle.conv2d_bwd(x, w, dy, pad=pad, stride=stride, need_dx=True)
    _close(dw, ref["dw"], name + " dW")
    _close(dx, ref["dx"], name + " dX")
    if bias:
        _close(db, ref["db"], name + " db")


def test_reference_gradient_maker_emits_no_bias_for_two_input_conv():
    # conv_gradient_op.cc:56-71: a bias-less Conv gets ConvGradient(no_bias=1) -> {dW, dX}
    from sad_b200 import c2
    lib = _reflib()
    op = c2.CreateOperator("Conv", ["X", "W"], ["Y"], kernel=3, pad=1, stride=1, device_option=c2.DeviceOption(c2.CPU))
    text = lib.GetGradientDefs(op, ["Y_grad"])
    assert 'name: "no_bias"' in text and 'output: "W_grad"' in text and 'output: "X_grad"' in text and "b_grad" not in text
