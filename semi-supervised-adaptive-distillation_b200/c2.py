"""Host-side mirror of the slice of the Caffe2 Python surface this path is driven through.

The reference builds OperatorDefs in Python (caffe2/caffe2/python/core.py: CreateOperator, Net),
asks C++ for gradient ops (core.py:1818) and runs them through a Workspace
(caffe2/caffe2/python/workspace.py: FeedBlob / RunOperatorOnce / CreateNet / RunNet / FetchBlob).
This module keeps those names and argument meanings on top of the shim's C handle API
(csrc/caffe2_shim/shim_c_api.cc), so tests read like Caffe2 operator tests.  Tensors are torch
tensors (CUDA for device_option CUDA); blobs BORROW their memory.

`OperatorLibrary(path)` can load either the product library (libcaffe2_detectron_ops_gpu.so) or
the GPU oracle built from the unmodified reference sources (oracle/_ref/libref_ops.so): both
export the same handle API, which is what makes op-for-op comparison on identical inputs possible.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import native

CPU, CUDA = 0, 1
_DT = {torch.float32: 1, torch.int32: 2}
_DT_INV = {1: torch.float32, 2: torch.int32}


class EnforceNotMet(RuntimeError):
    """caffe2::EnforceNotMet surfaced through the C API."""


class DeviceOption:
    def __init__(self, device_type=CPU, cuda_gpu_id=0):
        self.device_type, self.cuda_gpu_id = device_type, cuda_gpu_id


def _quote(s):
    return '"' + s.replace("\\", "\\\\").replace('"', '\\"') + '"'


class OperatorDef:
    def __init__(self, type, inputs, outputs, name="", device_option=None, engine="", **kwargs):
        self.type, self.name, self.engine = type, name, engine
        self.input = [inputs] if isinstance(inputs, str) else list(inputs)
        self.output = [outputs] if isinstance(outputs, str) else list(outputs)
        self.device_option = device_option
        self.arg = dict(kwargs)
        self.is_gradient_op = False

    def to_text(self, indent=""):
        t = []
        for s in self.input:
            t.append('%sinput: %s' % (indent, _quote(s)))
        for s in self.output:
            t.append('%soutput: %s' % (indent, _quote(s)))
        if self.name:
            t.append('%sname: %s' % (indent, _quote(self.name)))
        t.append('%stype: %s' % (indent, _quote(self.type)))
        for k, v in self.arg.items():
            # caffe2/python/utils.py MakeArgument: python float -> f, int/bool -> i, str -> s, lists -> floats/ints
            if isinstance(v, bool):
                body = "i: %d" % int(v)
            elif isinstance(v, (int, np.integer)):
                body = "i: %d" % int(v)
            elif isinstance(v, (float, np.floating)):
                body = "f: %r" % float(v)
            elif isinstance(v, str):
                body = "s: %s" % _quote(v)
            elif isinstance(v, (list, tuple)) and all(isinstance(e, (int, np.integer)) for e in v):
                body = " ".join("ints: %d" % int(e) for e in v)
            elif isinstance(v, (list, tuple)):
                body = " ".join("floats: %r" % float(e) for e in v)
            else:
                raise TypeError("unsupported argument %s=%r" % (k, v))
            t.append('%sarg { name: %s %s }' % (indent, _quote(k), body))
        if self.device_option is not None:
            t.append('%sdevice_option { device_type: %d cuda_gpu_id: %d }' % (
                indent, self.device_option.device_type, self.device_option.cuda_gpu_id))
        if self.engine:
            t.append('%sengine: %s' % (indent, _quote(self.engine)))
        if self.is_gradient_op:
            t.append('%sis_gradient_op: true' % indent)
        return "\n".join(t) + "\n"


def CreateOperator(operator_type, inputs, outputs, name="", device_option=None, engine="", **kwargs):
    """caffe2.python.core.CreateOperator"""
    return OperatorDef(operator_type, inputs, outputs, name, device_option, engine, **kwargs)


class NetDef:
    def __init__(self, name, ops=(), device_option=None):
        self.name, self.op, self.device_option = name, list(ops), device_option
        self.text_override = None

    def to_text(self):
        if self.text_override is not None:
            return self.text_override
        t = ['name: %s' % _quote(self.name)]
        for op in self.op:
            t.append("op {\n%s}" % op.to_text("  "))
        if self.device_option is not None:
            t.append('device_option { device_type: %d cuda_gpu_id: %d }' % (
                self.device_option.device_type, self.device_option.cuda_gpu_id))
        return "\n".join(t) + "\n"


class OperatorLibrary:
    """One dlopen'ed operator library (registration happens in its static initialisers, as with
    dyndep.InitOpsLibrary in the reference: detectron/lib/utils/c2.py:39-42)."""

    def __init__(self, path=None):
        path = path or native.OPS_LIB_PATH
        if not os.path.exists(path):
            raise ImportError("operator library %s is missing; build it first (there is no fallback)" % path)
        self.path = path
        l = C.CDLL(path)
        l.c2_last_error.restype = C.c_char_p
        l.c2_workspace_create.restype = C.c_void_p
        l.c2_workspace_destroy.argtypes = [C.c_void_p]
        l.c2_feed_external.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_void_p]
        l.c2_tensor_info.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
        l.c2_fetch.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        l.c2_has_blob.argtypes = [C.c_void_p, C.c_char_p]
        l.c2_has_operator.argtypes = [C.c_char_p, C.c_int]
        l.c2_has_schema.argtypes = [C.c_char_p]
        l.c2_schema_arity.argtypes = [C.c_char_p] + [C.POINTER(C.c_int)] * 4
        l.c2_registered_operators.restype = C.c_char_p
        l.c2_registered_operators.argtypes = [C.c_int]
        l.c2_run_operator_once.argtypes = [C.c_void_p, C.c_char_p]
        l.c2_create_net.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        l.c2_run_net.argtypes = [C.c_void_p, C.c_char_p]
        l.c2_run_net_async.argtypes = [C.c_void_p, C.c_char_p]
        l.c2_gradient_defs.restype = C.c_char_p
        l.c2_gradient_defs.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int]
        l.c2_adopt_stream.argtypes = [C.c_int, C.c_int, C.c_void_p]
        l.c2_normalize_net_text.restype = C.c_char_p
        l.c2_normalize_net_text.argtypes = [C.c_char_p]
        if hasattr(l, "c2_fuse_adaptive_distill_ops"):
            l.c2_fuse_adaptive_distill_ops.restype = C.c_char_p
            l.c2_fuse_adaptive_distill_ops.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        self.l = l

    def _check(self, rc):
        if rc != 0:
            raise EnforceNotMet(self.l.c2_last_error().decode() or "operator returned false")

    def RegisteredOperators(self, device_type=CUDA):
        s = self.l.c2_registered_operators(device_type).decode()
        return sorted(s.split(",")) if s else []

    def HasOperator(self, type, device_type=CUDA):
        return bool(self.l.c2_has_operator(type.encode(), device_type))

    def SchemaArity(self, type):
        v = [C.c_int() for _ in range(4)]
        if self.l.c2_schema_arity(type.encode(), *[C.byref(x) for x in v]) != 0:
            return None
        return tuple(x.value for x in v)

    def GetGradientDefs(self, op, g_outputs):
        """Gradient op defs (as NetDef text) + gradient blob name per forward input."""
        arr = (C.c_char_p * len(g_outputs))(*[(g or "").encode() for g in g_outputs])
        r = self.l.c2_gradient_defs(op.to_text().encode(), arr, len(g_outputs))
        if r is None:
            raise EnforceNotMet(self.l.c2_last_error().decode())
        return r.decode()

    def NormalizeNetText(self, text):
        r = self.l.c2_normalize_net_text(text.encode())
        if r is None:
            raise EnforceNotMet(self.l.c2_last_error().decode())
        return r.decode()

    def FuseAdaptiveDistillOps(self, net_text):
        n = C.c_int()
        r = self.l.c2_fuse_adaptive_distill_ops(net_text.encode(), C.byref(n))
        if r is None:
            raise EnforceNotMet("fusion pass failed")
        return r.decode(), n.value

    def Workspace(self):
        return Workspace(self)


class Workspace:
    """caffe2.python.workspace, as an object."""

    def __init__(self, oplib):
        self.lib = oplib
        self.h = C.c_void_p(oplib.l.c2_workspace_create())
        self._keep = {}

    def __del__(self):
        try:
            if self.h:
                self.lib.l.c2_workspace_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def FeedBlob(self, name, tensor):
        """Blob `name` borrows `tensor` (torch, float32/int32, CPU or CUDA)."""
        if isinstance(tensor, np.ndarray):
            tensor = torch.from_numpy(np.ascontiguousarray(tensor))
        if tensor.dtype not in _DT:
            raise TypeError("only float32/int32 blobs cross this boundary")
        tensor = tensor.contiguous()
        self._keep[name] = tensor
        dims = (C.c_int64 * max(1, tensor.dim()))(*tensor.shape)
        self.lib._check(self.lib.l.c2_feed_external(self.h, name.encode(), CUDA if tensor.is_cuda else CPU,
                                                    _DT[tensor.dtype], dims, tensor.dim(),
                                                    C.c_void_p(tensor.data_ptr())))

    def HasBlob(self, name):
        return bool(self.lib.l.c2_has_blob(self.h, name.encode()))

    def FetchBlob(self, name):
        dev, dt, nd, ptr = C.c_int(), C.c_int(), C.c_int(), C.c_void_p()
        dims = (C.c_int64 * 8)()
        self.lib._check(self.lib.l.c2_tensor_info(self.h, name.encode(), C.byref(dev), C.byref(dt), dims, C.byref(nd), C.byref(ptr)))
        shape = tuple(dims[i] for i in range(nd.value))
        out = np.empty(shape, dtype=np.float32 if dt.value == 1 else np.int32)
        self.lib._check(self.lib.l.c2_fetch(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def RunOperatorOnce(self, op):
        text = op if isinstance(op, str) else op.to_text()
        self._adopt_current_stream()
        self.lib._check(self.lib.l.c2_run_operator_once(self.h, text.encode()))
        return True

    def CreateNet(self, net, overwrite=False):
        text = net if isinstance(net, str) else net.to_text()
        self._adopt_current_stream()
        self.lib._check(self.lib.l.c2_create_net(self.h, text.encode(), int(overwrite)))

    def RunNet(self, name):
        self._adopt_current_stream()
        self.lib._check(self.lib.l.c2_run_net(self.h, name.encode()))

    def RunNetAsync(self, name):
        self.lib._check(self.lib.l.c2_run_net_async(self.h, name.encode()))

    def _adopt_current_stream(self):
        # make ops enqueue on torch's current stream so they order with the tensors we feed
        if torch.cuda.is_available():
            dev = torch.cuda.current_device()
            self.lib._check(self.lib.l.c2_adopt_stream(dev, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
