"""CPU: size-independent properties of the reference's formulas as restated by the oracle (hypothesis-driven shapes and arguments).
They pin behaviours the golden vectors only sample: the ignore mask is per (image, anchor, y, x) and independent of the class;
the normaliser enters as 1 / max(wp, 1); `scale` and `d_loss` are plain multipliers applied after the per-element work
(sigmoid_adaptive_distillation_loss_op.cu:35-51, 63-64, 98-102, 136-138, 166-168); PowSum is additive over its inputs in input
order (pow_sum_op.cu:34-40); SigmoidFocalLoss shares the label indexing (sigmoid_focal_loss_op.cu:34-43)."""
import numpy as np
from hypothesis import example, given, settings
from hypothesis import strategies as st

from oracle import cpu_oracle as O

shape_st = st.tuples(st.integers(1, 2), st.integers(1, 3), st.integers(1, 4), st.integers(1, 5), st.integers(1, 7))  # N, A, C, H, W


def _case(seed, n, a, c, h, w):
    rng = np.random.default_rng(seed)
    x = rng.normal(-2.0, 2.0, size=(n, a * c, h, w)).astype(np.float32)
    t = np.clip(1.0 / (1.0 + np.exp(-rng.normal(-2.0, 2.5, size=x.shape))), 1e-6, 1 - 1e-6).astype(np.float32)
    g = rng.choice(np.array([-1, 0, 0, 0, 1, 7], dtype=np.int32), size=(n, a, h, w)).astype(np.int32)
    return x, t, g


@settings(max_examples=40, deadline=None, derandomize=True)
@given(shape_st, st.integers(0, 2 ** 20), st.sampled_from([0.0, 1.0, 2.0]), st.sampled_from([0.25, 0.5]), st.sampled_from([0.0, 1.0]))
@example(shape=(1, 2, 2, 5, 5), seed=1088, gamma=0.0, alpha=0.25, beta=1.0)   # gamma = 0 with an adaptive target of exactly 0: 0 * powf(0, -1) = NaN
def test_ignore_mask_is_per_anchor_location_and_class_independent(shape, seed, gamma, alpha, beta):
    n, a, c, h, w = shape
    x, t, g = _case(seed, n, a, c, h, w)
    args = dict(gamma=gamma, alpha=alpha, beta=beta, num_classes=c)
    _, elems = O.distill_loss(x, t, g, 3.0, return_elements=True, **args)
    grad = O.distill_grad(x, t, g, 3.0, **args)
    keep = np.repeat(g != -1, c, axis=1)                      # (N, A, H, W) -> (N, A*C, H, W): channel a*C + d reads anchor a
    assert np.all(elems[~keep] == 0)
    # the gradient kernel MULTIPLIES by the mask (loss_op.cu:98-101: ... * d_loss * (t != ignored_label)), so an ignored position is
    # exactly 0 unless the unmasked expression itself is not finite — gamma = 0 with an adaptive target of 0 gives
    # 0 * powf(0, -1) = NaN in the reference, and NaN * 0 stays NaN.  Everything else must be an exact zero.
    unmasked = O.distill_grad(x, t, np.zeros_like(g), 3.0, **args)
    ignored = grad[~keep]
    assert np.all((ignored == 0) | (np.isnan(ignored) & ~np.isfinite(unmasked[~keep])))
    # changing the VALUE of a kept label (any class id, background) changes nothing: only `!= ignored_label` is read
    g2 = np.where(g != -1, 55, -1).astype(np.int32)
    _, elems2 = O.distill_loss(x, t, g2, 3.0, return_elements=True, **args)
    assert np.array_equal(elems, elems2) and np.array_equal(grad, O.distill_grad(x, t, g2, 3.0, **args), equal_nan=True)
    # another ignored_label value moves the mask with it
    _, elems3 = O.distill_loss(x, t, g, 3.0, return_elements=True, ignored_label=7, **args)
    assert np.all(elems3[np.repeat(g == 7, c, axis=1)] == 0)


@settings(max_examples=30, deadline=None, derandomize=True)
@given(shape_st, st.integers(0, 2 ** 20))
def test_normaliser_clamp_scale_and_upstream_gradient_are_multipliers(shape, seed):
    n, a, c, h, w = shape
    x, t, g = _case(seed, n, a, c, h, w)
    args = dict(gamma=2.0, alpha=0.5, beta=0.0, num_classes=c)
    # Np = max(wp, 1): every normaliser <= 1 gives the same numbers
    l1, g1 = O.distill_loss(x, t, g, 1.0, **args), O.distill_grad(x, t, g, 1.0, **args)
    for wp in (0.0, 0.3, -5.0):
        assert O.distill_loss(x, t, g, wp, **args) == l1 and np.array_equal(O.distill_grad(x, t, g, wp, **args), g1)
    # exact powers of two commute with fp32 rounding: scale and d_loss multiply, a 4x normaliser divides
    assert np.array_equal(O.distill_grad(x, t, g, 1.0, scale=0.5, **args), g1 * np.float32(0.5))
    assert np.array_equal(O.distill_grad(x, t, g, 1.0, d_loss=0.25, **args), g1 * np.float32(0.25))
    assert np.array_equal(O.distill_grad(x, t, g, 4.0, **args), g1 * np.float32(0.25))
    assert O.distill_loss(x, t, g, 1.0, scale=0.5, **args) == np.float32(l1 * np.float32(0.5))
    # scale = 0 (legal: CAFFE_ENFORCE(scale_ >= 0)) zeroes both
    assert O.distill_loss(x, t, g, 1.0, scale=0.0, **args) == 0 and not O.distill_grad(x, t, g, 1.0, scale=0.0, **args).any()


@settings(max_examples=30, deadline=None, derandomize=True)
@given(st.lists(st.integers(1, 300), min_size=1, max_size=5), st.integers(0, 2 ** 20), st.sampled_from([1.0, 1.8, 2.0, 0.5]))
def test_pow_sum_is_additive_over_inputs_in_input_order(sizes, seed, power):
    rng = np.random.default_rng(seed)
    xs = [rng.uniform(1e-6, 1.0, size=s).astype(np.float32) for s in sizes]
    total = np.float32(0)
    for x in xs:                                               # res = res + s, a float add per input (pow_sum_op.cu:39)
        total = np.float32(total + O.pow_sum([x], power))
    assert O.pow_sum(xs, power) == total
    if power == 1.0:
        assert abs(float(O.pow_sum(xs, 1.0)) - float(sum(x.astype(np.float64).sum() for x in xs))) <= 1e-5 * sum(sizes)


@settings(max_examples=30, deadline=None, derandomize=True)
@given(shape_st, st.integers(0, 2 ** 20))
def test_focal_loss_label_semantics(shape, seed):
    n, a, c, h, w = shape
    x, _, g = _case(seed, n, a, c, h, w)
    g = np.where(g > c, 1, g).astype(np.int32)                 # class ids in 1..C
    fg = np.asarray([2.0], np.float32)
    grad = O.focal_grad(x, g, fg, gamma=2.0, alpha=0.25, num_classes=c)
    # ignored anchors (label -1): zero for every class (sigmoid_focal_loss_op.cu:42-43 c2 = (t >= 0 & t != d + 1))
    assert not grad[np.repeat(g == -1, c, axis=1)].any()
    # a foreground anchor pushes its own class logit up (negative gradient) and every other class down
    gv = grad.reshape(n, a, c, h, w)
    for cls in range(1, c + 1):
        m = g == cls
        if m.any():
            own = gv[:, :, cls - 1][m]
            assert (own <= 0).all()
            for other in range(c):
                if other != cls - 1:
                    assert (gv[:, :, other][m] >= 0).all()
