"""Runs the UNMODIFIED reference CUDA operators (oracle/_ref/libref_ops.so: the reference's own
pow_sum_op.cu / sigmoid_adaptive_distillation_loss_op.cu behind the shim) on the inputs stored in
distill_kat.npz and writes their outputs.  Needs a GPU:

    gpurun -- python tests/golden/make_ref_gpu_golden.py      # -> gpurun_out/ref_gpu_kat.npz

The result is committed as tests/golden/ref_gpu_kat.npz: "outputs of the reference itself", the
vectors that pin the CPU oracle (tests/test_oracle_golden.py) on machines without a GPU.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = ["vec", "ragged", "beta", "gamma1", "gamma3"]


def run_reference_ops(reflib):
    import torch
    from sad_b200 import c2
    kat = np.load(os.path.join(HERE, "distill_kat.npz"))
    dev = c2.DeviceOption(c2.CUDA, 0)
    out = {}
    for name in CASES:
        gamma, alpha, beta, scale, C, ign = kat[name + "_args"]
        args = dict(gamma=float(gamma), alpha=float(alpha), beta=float(beta), scale=float(scale),
                    num_classes=int(C), ignored_label=int(ign))
        ws = reflib.Workspace()
        ws.FeedBlob("X", torch.from_numpy(kat[name + "_x"]).cuda())
        ws.FeedBlob("T", torch.from_numpy(kat[name + "_t"]).cuda())
        ws.FeedBlob("G", torch.from_numpy(kat[name + "_g"]).cuda())
        ws.FeedBlob("wp", torch.tensor([float(kat[name + "_wp"])], device="cuda"))
        ws.FeedBlob("dl", torch.tensor(float(kat[name + "_dloss"]), device="cuda"))
        ws.RunOperatorOnce(c2.CreateOperator("SigmoidAdaptiveDistillLoss", ["X", "T", "G", "wp"], ["loss"], device_option=dev, **args))
        ws.RunOperatorOnce(c2.CreateOperator("SigmoidAdaptiveDistillLossGradient", ["X", "T", "G", "wp", "dl"], ["dX"],
                                             device_option=dev, **args))
        out[name + "_loss"] = ws.FetchBlob("loss")
        out[name + "_grad"] = ws.FetchBlob("dX")
    ws = reflib.Workspace()
    for i in range(3):
        ws.FeedBlob("in%d" % i, torch.from_numpy(kat["ps_in%d" % i]).cuda())
    for power in (1.0, 1.8, 2.0, 3.0):
        ws.RunOperatorOnce(c2.CreateOperator("PowSum", ["in0", "in1", "in2"], ["s"], device_option=dev, power=float(power)))
        out["ps_%g" % power] = ws.FetchBlob("s")
    return out


def main():
    from oracle import cpu_oracle
    from sad_b200 import c2
    reflib = c2.OperatorLibrary(cpu_oracle.REF_GPU_LIB)
    out = run_reference_ops(reflib)
    dst = os.path.join(ROOT, "gpurun_out")
    os.makedirs(dst, exist_ok=True)
    path = os.path.join(dst, "ref_gpu_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
