"""CPU: the restatement of AffineChannel(+Gradient) / UpsampleNearest(+Gradient) (oracle/body_oracle.c, following
affine_channel_op.cu:22-48 and upsample_nearest_op.cu:66-113) against independent numpy definitions, and the host side of the four
operators: registration under the reference's names on both devices, schema arity, gradient makers, CAFFE_NOT_IMPLEMENTED on CPU.
The reference holds no test or golden vector for these operators; on the GPU the reference's own kernels run beside
the product (tests/test_body_ops_gpu.py)."""
import numpy as np
import pytest

from sad_b200 import c2


@pytest.fixture(scope="module")
def oplib():
    return c2.OperatorLibrary()


@pytest.mark.parametrize("shape", [(2, 3, 4, 8), (1, 5, 3, 7), (2, 64, 1, 1)])
def test_affine_channel_oracle_is_one_fma_per_element(oracle, shape):
    rng = np.random.default_rng(sum(shape))
    x = rng.normal(size=shape).astype(np.float32)
    s = rng.normal(1.0, 0.3, size=shape[1]).astype(np.float32)
    b = rng.normal(size=shape[1]).astype(np.float32)
    y = oracle.affine_channel(x, s, b)
    exact = x.astype(np.float64) * s.astype(np.float64)[None, :, None, None] + b.astype(np.float64)[None, :, None, None]
    # x * s is exact in fp64; adding b and rounding once more can differ from the single-rounded FMA by a double rounding only
    np.testing.assert_array_max_ulp(y, exact.astype(np.float32), maxulp=1)
    assert np.mean(y == exact.astype(np.float32)) > 0.999
    dx = oracle.affine_channel(x, s, None)
    assert np.array_equal(dx, x * s[None, :, None, None])


@pytest.mark.parametrize("shape,scale", [((2, 3, 4, 8), 2), ((1, 2, 3, 5), 2), ((1, 2, 3, 5), 3), ((4, 6, 8), 2), ((1, 1, 1, 1), 1)])
def test_upsample_nearest_oracle_matches_numpy(oracle, shape, scale):
    rng = np.random.default_rng(len(shape) * 100 + scale)
    x = rng.normal(size=shape).astype(np.float32)
    y = oracle.upsample_nearest(x, scale)
    assert np.array_equal(y, np.repeat(np.repeat(x, scale, axis=-2), scale, axis=-1))
    dy = rng.normal(size=y.shape).astype(np.float32)
    dx = oracle.upsample_nearest_grad(x.shape, dy, scale)
    # the reference's order: x offset outer, y offset inner, fp32 running sum starting from the zero fill
    acc = np.zeros(shape, np.float32)
    for i in range(scale):
        for j in range(scale):
            acc = acc + dy[..., j::scale, i::scale]
    assert np.array_equal(dx, acc)


def test_body_operators_registered_under_reference_names(oplib):
    # affine_channel_op.cc:21-24, .cu:99-102; upsample_nearest_op.cc:21-24, .cu:214-217
    for dev in (c2.CPU, c2.CUDA):
        for name in ("AffineChannel", "AffineChannelGradient", "UpsampleNearest", "UpsampleNearestGradient"):
            assert oplib.HasOperator(name, dev), (name, dev)
    assert oplib.SchemaArity("AffineChannel") == (3, 3, 1, 1)
    assert oplib.SchemaArity("AffineChannelGradient") == (2, 2, 1, 1)
    assert oplib.SchemaArity("UpsampleNearest") == (1, 1, 1, 1)
    assert oplib.SchemaArity("UpsampleNearestGradient") == (2, 2, 1, 1)


def _fields(text):
    line = lambda key: [l.split('"')[1] for l in text.splitlines() if l.strip().startswith(key + ":")]
    return line("type"), line("input"), line("output")


def test_body_gradient_makers_follow_reference(oplib):
    dev = c2.DeviceOption(c2.CUDA, 0)
    # affine_channel_op.cc:71-80: AffineChannelGradient(scale, dY) -> dX
    op = c2.CreateOperator("AffineChannel", ["x", "s", "b"], ["y"], device_option=dev)
    types, ins, outs = _fields(oplib.GetGradientDefs(op, ["y_grad"]))
    assert types == ["AffineChannelGradient"] and ins == ["s", "y_grad"] and outs == ["x_grad"]
    # upsample_nearest_op.cc:61-72: UpsampleNearestGradient(X, dY) -> dX, the scale argument copied from the forward def
    op = c2.CreateOperator("UpsampleNearest", ["x"], ["y"], device_option=dev, scale=2)
    text = oplib.GetGradientDefs(op, ["y_grad"])
    types, ins, outs = _fields(text)
    assert types == ["UpsampleNearestGradient"] and ins == ["x", "y_grad"] and outs == ["x_grad"]
    assert 'name: "scale"' in text and "i: 2" in text


def test_body_operators_not_implemented_on_cpu_like_the_reference(oplib):
    # affine_channel_op.h:33-36, upsample_nearest_op.h:37-40
    ws = oplib.Workspace()
    ws.FeedBlob("x", np.ones((1, 2, 2, 2), np.float32))
    ws.FeedBlob("s", np.ones((2,), np.float32))
    with pytest.raises(c2.EnforceNotMet, match="Not Implemented"):
        ws.RunOperatorOnce(c2.CreateOperator("AffineChannel", ["x", "s", "s"], ["y"]))
    with pytest.raises(c2.EnforceNotMet, match="Not Implemented"):
        ws.RunOperatorOnce(c2.CreateOperator("UpsampleNearest", ["x"], ["y"], scale=2))


@pytest.mark.parametrize("shape", [(2, 3, 4, 6), (1, 5, 2, 2), (1, 1, 10, 14)], ids=lambda s: "x".join(map(str, s)))
def test_upsample_add_oracle_matches_numpy(oracle, shape):
    # FPN.py:230-249: td = UpsampleNearest(top, 2); out = Sum([lateral, td]) — in both memory views
    rng = np.random.default_rng(sum(shape))
    n, c, h, w = shape
    lat = rng.standard_normal(shape).astype(np.float32)
    top = rng.standard_normal((n, c, h // 2, w // 2)).astype(np.float32)
    ref = lat + top.repeat(2, axis=2).repeat(2, axis=3)
    assert np.array_equal(oracle.upsample2_add(top, lat), ref)
    got_cl = oracle.upsample2_add(top.transpose(0, 2, 3, 1), lat.transpose(0, 2, 3, 1), inner=c)
    assert np.array_equal(got_cl.transpose(0, 3, 1, 2), ref)
