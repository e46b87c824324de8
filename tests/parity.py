"""The parity gate of BASELINE.md §3 / SURVEY.md Appendix D item 2, in one place.

  loss     : |got - ref| <= 1e-4 * |ref|                               (fp32, relative)
  gradient : max|d| <= 1e-4 * max|ref|   AND   |d_i| <= 1e-4*|ref_i| + 1e-6*max|ref|  elementwise
  indexing : bit-exact (tests compare integer index maps / masks with array_equal)

The gradient gate has an absolute floor because the reference formula itself cancels
catastrophically where student ~= teacher (AT = 1 - exp(-DL), DL -> 0): the reference evaluated
in fp32 vs fp64 already violates a pure per-element 1e-4 relative bound on ~0.3 % of elements.
"""
import numpy as np

LOSS_RTOL = 1e-4
GRAD_RTOL = 1e-4
GRAD_FLOOR = 1e-6


def assert_loss_close(got, ref, what="loss"):
    got, ref = float(got), float(ref)
    if np.isnan(ref):
        assert np.isnan(got), "%s: reference is NaN, got %r" % (what, got)
        return
    assert abs(got - ref) <= LOSS_RTOL * abs(ref) + 1e-30, "%s: got %.9g ref %.9g rel %.3g" % (
        what, got, ref, abs(got - ref) / max(abs(ref), 1e-30))


def assert_grad_close(got, ref, what="grad"):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    nan_ref = np.isnan(ref)
    assert np.array_equal(np.isnan(got), nan_ref), "%s: NaN pattern differs from the reference" % what
    if nan_ref.all():
        return
    g, r = got[~nan_ref], ref[~nan_ref]
    m = np.max(np.abs(r)) if r.size else 0.0
    d = np.abs(g - r)
    assert d.max(initial=0.0) <= GRAD_RTOL * m + 1e-30, "%s: max|d| %.3g > 1e-4*max|ref| %.3g" % (what, d.max(), m)
    bound = GRAD_RTOL * np.abs(r) + GRAD_FLOOR * m
    bad = d > bound + 1e-30
    assert not bad.any(), "%s: %d of %d elements outside |d|<=1e-4|ref|+1e-6 max|ref| (worst %.3g vs %.3g)" % (
        what, int(bad.sum()), r.size, d[bad].max(), bound[bad][np.argmax(d[bad])])


def assert_reduced_close(got, ref_order, exact, what="sum", exact_rtol=2e-5):
    """Gate for a scalar that the reference obtains with its single-block fp32 Sum
    (math_gpu.cu:1021-1058: 128 lanes, each adding up to 115 k terms sequentially at config-2 size).
    That accumulation is itself up to ~1e-4 away from the exact sum of the same fp32 terms, so
    comparing only against the reference-order value would test the reference's rounding noise.
    `ref_order` is the oracle's value in the reference's summation order, `exact` the fp64 sum of the
    oracle's per-element fp32 terms.  Required: within exact_rtol of the exact value, and within
    1e-4 of the reference-order value once the reference's own deviation from exact is allowed for."""
    got, ref_order, exact = float(got), float(ref_order), float(exact)
    assert abs(got - exact) <= exact_rtol * abs(exact) + 1e-30, "%s: got %.9g exact %.9g rel %.3g" % (
        what, got, exact, abs(got - exact) / max(abs(exact), 1e-30))
    assert abs(got - ref_order) <= LOSS_RTOL * abs(ref_order) + abs(ref_order - exact) + 1e-30, \
        "%s: got %.9g reference-order %.9g (exact %.9g)" % (what, got, ref_order, exact)
