"""Runs the UNMODIFIED reference CPU convolution operators (caffe2/caffe2/operators/conv_op.cc, conv_gradient_op.cc
with conv_op_impl.h:31-180, 357-560, compiled from /root/reference into oracle/_ref/libref_ops.so by oracle/Makefile)
on seeded inputs and writes inputs' seeds + the operators' outputs:

    python tests/golden/make_ref_cpu_conv_golden.py        # -> tests/golden/ref_cpu_conv_kat.npz

No GPU needed (CPU operators).  The file pins oracle/conv_oracle.c (tests/test_conv_reference_pin.py) on machines
where neither /root/reference nor the prebuilt oracle/_ref exists.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# (name, N, Cin, Cout, H, W, bias): the head's shape class (3x3, stride 1, pad 1; retinanet_heads.py:105-152) from
# tiny / ragged to the real channel counts, including the 256->720 and 256->36 prediction convolutions on a small map
CASES = [
    ("tiny", 1, 4, 5, 3, 3, True),
    ("ragged", 2, 16, 24, 9, 13, True),
    ("one_px", 1, 8, 8, 1, 1, True),
    ("one_row", 1, 8, 16, 1, 7, False),
    ("tower", 1, 256, 256, 5, 8, True),
    ("cls_pred", 1, 256, 720, 5, 8, True),
    ("box_pred", 2, 256, 36, 4, 7, True),
]


def make_inputs(name, n, cin, cout, h, w, bias):
    seed = int.from_bytes(name.encode(), "little") % (2 ** 31)
    rng = np.random.default_rng(seed)
    x = np.maximum(rng.standard_normal((n, cin, h, w)), 0).astype(np.float32) * 0.5   # post-ReLU-like
    wt = (rng.standard_normal((cout, cin, 3, 3)) * 0.01).astype(np.float32)            # N(0, 0.01) (retinanet_heads.py:105-107)
    b = (rng.standard_normal(cout) * 0.1).astype(np.float32) if bias else None
    dy = rng.standard_normal((n, cout, h, w)).astype(np.float32)
    return x, wt, b, dy


def sample(v, limit=16384):
    """Fixture size: arrays above `limit` elements are stored as every k-th element of the flattened array (prime
    stride, so every channel / tap / pixel residue is visited); the test applies the same rule to what it compares."""
    flat = np.ascontiguousarray(v).ravel()
    if flat.size <= limit:
        return flat
    k = -(-flat.size // limit)
    while any(k % p == 0 for p in (2, 3, 5, 7)):
        k += 1
    return flat[::k]


def run_reference(reflib, x, w, b, dy, kernel=3, pad=1, stride=1):
    from sad_b200 import c2
    cpu = c2.DeviceOption(c2.CPU)
    ws = reflib.Workspace()
    ws.FeedBlob("X", x)
    ws.FeedBlob("W", w)
    ws.FeedBlob("dY", dy)
    ins = ["X", "W"]
    if b is not None:
        ws.FeedBlob("b", b)
        ins.append("b")
    fwd = c2.CreateOperator("Conv", ins, ["Y"], kernel=kernel, pad=pad, stride=stride, order="NCHW", device_option=cpu)
    ws.RunOperatorOnce(fwd)
    # the gradient operator exactly as the reference's own gradient maker emits it (conv_gradient_op.cc:35-77)
    ws.CreateNet("name: \"g\"\n" + reflib.GetGradientDefs(fwd, ["dY"]).split("external_output")[0])
    ws.RunNet("g")
    out = {"y": ws.FetchBlob("Y"), "dw": ws.FetchBlob("W_grad"), "dx": ws.FetchBlob("X_grad")}
    if b is not None:
        out["db"] = ws.FetchBlob("b_grad")
    return out


def main():
    from oracle import cpu_oracle
    from sad_b200 import c2
    reflib = c2.OperatorLibrary(cpu_oracle.REF_GPU_LIB)
    out = {}
    for case in CASES:
        x, w, b, dy = make_inputs(*case)
        for k, v in run_reference(reflib, x, w, b, dy).items():
            out["%s_%s" % (case[0], k)] = sample(v)
    path = os.path.join(HERE, "ref_cpu_conv_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
