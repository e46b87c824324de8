"""B200-native adaptive-distillation hot path (PowSum, SigmoidAdaptiveDistillLoss(+Gradient),
RetinaNet FPN head) behind the reference's Caffe2 operator interface.

Layout:
  csrc/kernels/      sm_100a CUDA kernels + the C ABI of include/sad_b200.h  -> libsad_b200.so
  csrc/caffe2_shim/  operator-boundary shim (Operator<CUDAContext>, registries, NetDef text)
  csrc/ops/          operator classes registered under the reference's names
                     -> libcaffe2_detectron_ops_gpu.so
  native.py          ctypes binding of the C ABI (fails loudly if the library is missing)
  ops.py             torch-tensor convenience layer over the C ABI (device pointers + streams)
  c2.py              host-side mirror of the Caffe2 python surface the path is driven through
  retinanet_heads.py mirror of detectron/lib/modeling/retinanet_heads.py:313-352 (graph wiring)
  synthetic.py       seeded COCO-shaped synthetic inputs
"""
__version__ = "0.1.0"
