"""Seeded synthetic inputs with the shapes and value distributions of SURVEY.md §8(d).

Shapes follow the reference's padding/stride rules (detectron/lib/utils/blob.py:52-56 pads to
COARSEST_STRIDE=128; FPN levels 3..7 have strides 8..128): a 600 px image is 3x640x1024, a 500 px
one 3x512x896.  Anchors per location A = 9 (retinanet_heads.py:72), classes C = 80.
"""
import numpy as np

NUM_ANCHORS = 9
NUM_CLASSES = 80
CLS_BIAS = -4.59511985013459  # -log((1 - pi) / pi), pi = 0.01 (retinanet_heads.py:55-59)


def level_shapes(scale_px=600):
    """[(H, W)] for FPN levels 3..7.  scale_px: 600, 500, or a padded (height, width) in multiples of 128."""
    if isinstance(scale_px, (tuple, list)):
        h, w = int(scale_px[0]), int(scale_px[1])
        if h <= 0 or w <= 0 or h % 128 or w % 128:
            raise ValueError("padded image size must be positive multiples of COARSEST_STRIDE = 128")
    elif scale_px == 600:
        h, w = 640, 1024
    elif scale_px == 500:
        h, w = 512, 896
    else:
        raise ValueError("scale_px must be 600 or 500")
    return [(h // s, w // s) for s in (8, 16, 32, 64, 128)]


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def make_level(rng, n, h, w, a=NUM_ANCHORS, c=NUM_CLASSES, stress=False):
    """Returns (logits f32 (n, a*c, h, w), teacher_prob f32 same, labels i32 (n, a, h, w))."""
    shape = (n, a * c, h, w)
    if stress:  # dense-positive stress set: logits and teacher logits uniform in (-12, 12)
        x = rng.uniform(-12.0, 12.0, size=shape)
        t = _sigmoid(rng.uniform(-12.0, 12.0, size=shape))
    else:
        x = rng.normal(CLS_BIAS, 2.0, size=shape)
        t = _sigmoid(rng.normal(CLS_BIAS, 2.5, size=shape))
    t = np.clip(t, 1e-6, 1.0 - 1e-6)  # the reference is NaN at exactly 0 or 1 (loss_op.cu:59,93)
    u = rng.random(size=(n, a, h, w))
    labels = np.zeros((n, a, h, w), dtype=np.int32)
    labels[u < 0.005] = -1
    fg = (u >= 0.005) & (u < 0.006)
    labels[fg] = rng.integers(1, c + 1, size=int(fg.sum()), dtype=np.int32)
    return x.astype(np.float32), t.astype(np.float32), labels


def make_pyramid(seed, n, scale_px=600, a=NUM_ANCHORS, c=NUM_CLASSES, stress=False):
    """One entry per FPN level 3..7, seeded with seed + level index."""
    out = []
    for i, (h, w) in enumerate(level_shapes(scale_px)):
        rng = np.random.default_rng(seed + i)
        out.append(make_level(rng, n, h, w, a, c, stress))
    return out


def anchors_in(levels):
    """Anchor count (one (n, a, y, x) position with its C logits) of a list of level tuples."""
    return int(sum(l[2].size for l in levels))
