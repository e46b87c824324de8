"""Generates tests/golden/distill_kat.npz.   Run: python tests/golden/make_golden.py

The reference ships NO golden vectors for this path.  Two kinds of vectors are
committed instead:
  * `f64_*`  — known answers from an INDEPENDENT float64 numpy evaluation of the published formula
               (this file, `formula_f64`), not from the C oracle: they pin the oracle.
  * `ora_*`  — outputs of the C oracle (oracle/distill_oracle.c) on the same seeded inputs: a
               regression pin for the oracle itself and the expected values of the GPU tests.
A third set, tests/golden/ref_gpu_kat.npz, holds outputs of the UNMODIFIED reference CUDA
operators run on a B200 (tests/golden/make_ref_gpu_golden.py, executed under gpurun).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_oracle  # noqa: E402

CASES = [
    # name, (N, A, C, H, W), args
    ("vec", (2, 3, 4, 5, 8), dict(gamma=2.0, alpha=0.5, beta=0.0, scale=1.0, ignored_label=-1)),
    ("ragged", (1, 2, 3, 3, 5), dict(gamma=2.0, alpha=0.25, beta=0.0, scale=0.125, ignored_label=-1)),
    ("beta", (1, 2, 4, 4, 4), dict(gamma=2.0, alpha=0.25, beta=1.0, scale=1.0, ignored_label=-1)),
    ("gamma1", (1, 2, 4, 4, 4), dict(gamma=1.0, alpha=0.75, beta=0.5, scale=2.0, ignored_label=-1)),
    ("gamma3", (2, 1, 5, 2, 6), dict(gamma=3.0, alpha=0.5, beta=0.0, scale=1.0, ignored_label=7)),
]


def make_inputs(seed, shape):
    N, A, C, H, W = shape
    rng = np.random.default_rng(seed)
    x = rng.normal(-2.0, 3.0, size=(N, A * C, H, W)).astype(np.float32)
    t = (1.0 / (1.0 + np.exp(-rng.normal(-2.0, 3.0, size=(N, A * C, H, W))))).clip(1e-6, 1 - 1e-6).astype(np.float32)
    g = rng.integers(-1, C + 1, size=(N, A, H, W)).astype(np.int32)
    g[rng.random(size=g.shape) < 0.2] = 7  # a second candidate ignore value
    return x, t, g


def formula_f64(x, t, g, wp, gamma, alpha, beta, scale, ignored_label, num_classes, d_loss=1.0):
    """Independent float64 statement of the operator pair (loss value, gradient)."""
    x = x.astype(np.float64)
    pt = t.astype(np.float64)
    N, D, H, W = x.shape
    A = D // num_classes
    keep = (g != ignored_label).astype(np.float64)           # (N, A, H, W)
    keep = np.repeat(keep, num_classes, axis=1)              # channel c -> anchor c // num_classes
    Np = max(float(wp), 1.0)
    p = 1.0 / (1.0 + np.exp(-x))
    softplus_negabs = np.log1p(np.exp(-np.abs(x)))
    log_p = np.minimum(x, 0) - softplus_negabs
    log_1mp = -np.maximum(x, 0) - softplus_negabs
    bce = -(pt * log_p + (1 - pt) * log_1mp)
    neg_entropy = pt * np.log(pt) + (1 - pt) * np.log(1 - pt)
    dist = bce + beta * neg_entropy
    E = np.exp(-dist)
    AT = 1 - E
    weighted = alpha * pt * log_p + (1 - alpha) * (1 - pt) * log_1mp
    loss = np.sum(-(AT ** gamma) * weighted * keep) / Np * scale
    grad = -(-(pt - p) * gamma * AT ** (gamma - 1) * E * weighted +
             AT ** gamma * (alpha * (pt - p) - (1 - 2 * alpha) * (1 - pt) * p)) * keep * d_loss / Np * scale
    return loss, grad


def main():
    out = {}
    for k, (name, shape, args) in enumerate(CASES):
        x, t, g = make_inputs(100 + k, shape)
        C = shape[2]
        wp = np.float32(3.5 + k)
        d_loss = 1.0 if k % 2 == 0 else 0.5
        out[name + "_x"], out[name + "_t"], out[name + "_g"] = x, t, g
        out[name + "_wp"] = wp
        out[name + "_dloss"] = np.float32(d_loss)
        out[name + "_args"] = np.array([args["gamma"], args["alpha"], args["beta"], args["scale"], C, args["ignored_label"]], dtype=np.float64)
        lo, gr = formula_f64(x, t, g, wp, args["gamma"], args["alpha"], args["beta"], args["scale"], args["ignored_label"], C, d_loss)
        out[name + "_f64_loss"], out[name + "_f64_grad"] = np.float64(lo), gr
        out[name + "_ora_loss"] = cpu_oracle.distill_loss(x, t, g, wp, num_classes=C, **args)
        out[name + "_ora_grad"] = cpu_oracle.distill_grad(x, t, g, wp, d_loss=d_loss, num_classes=C, **args)
    # PowSum
    rng = np.random.default_rng(7)
    ps = [rng.random(size=s).astype(np.float32) for s in (1000, 37, 4096 + 3)]
    for i, a in enumerate(ps):
        out["ps_in%d" % i] = a
    for power in (1.0, 1.8, 2.0, 3.0):
        out["ps_f64_%g" % power] = np.float64(sum(np.sum(a.astype(np.float64) ** power) for a in ps))
        out["ps_ora_%g" % power] = cpu_oracle.pow_sum(ps, power)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "distill_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
