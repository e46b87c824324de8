"""ctypes binding of include/sad_exchange.h: the gradient exchange of the data-parallel step as host C++ over NCCL.

Reference: detectron/lib/modeling/optimizer.py:72-92 (one NCCLAllreduce per gradient blob) and
caffe2/caffe2/contrib/nccl/cuda_nccl_gpu.cc:139-225 (event plumbing around ncclAllReduce).  Here the flat gradient buffer is
reduced in contiguous BUCKETS on a dedicated communication stream; a bucket is enqueued the moment the backward pass has
produced it, so the exchange overlaps the rest of the backward pass and the optimiser waits only for `join`.
torch.distributed is used once, to hand rank 0's NCCL id to the other ranks.  No fallback: a missing library or NCCL raises.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsad_exchange.so")
ID_BYTES = 128
_lib = None


class ExchangeError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libsad_exchange.so is missing (%s): build it with semi-supervised-adaptive-distillation_b200/build.py" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.sad_exchange_last_error.restype = C.c_char_p
        l.sad_exchange_unique_id.argtypes = [C.c_void_p]
        l.sad_exchange_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        l.sad_exchange_create_config.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        l.sad_exchange_max_ctas.argtypes = [C.c_void_p]
        l.sad_exchange_create_gather.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
        l.sad_exchange_gather_capacity.argtypes = [C.c_void_p]
        l.sad_exchange_gather_capacity.restype = C.c_size_t
        l.sad_exchange_gathered.argtypes = [C.c_void_p]
        l.sad_exchange_gathered.restype = C.c_uint64
        l.sad_exchange_slot_sum_f32.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        l.sad_exchange_destroy.argtypes = [C.c_void_p]
        l.sad_exchange_destroy.restype = None
        l.sad_exchange_world.argtypes = [C.c_void_p]
        l.sad_exchange_rank.argtypes = [C.c_void_p]
        l.sad_exchange_allreduce_async_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.sad_exchange_join.argtypes = [C.c_void_p, C.c_void_p]
        l.sad_exchange_flush.argtypes = [C.c_void_p]
        l.sad_exchange_plan_reset.argtypes = [C.c_void_p]
        l.sad_exchange_planned.argtypes = [C.c_void_p]
        l.sad_exchange_allreduce_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.sad_exchange_buckets.argtypes = [C.c_void_p]
        l.sad_exchange_buckets.restype = C.c_uint64
        l.sad_exchange_bytes.argtypes = [C.c_void_p]
        l.sad_exchange_bytes.restype = C.c_uint64
        _lib = l
    return _lib


def _check(rc):
    if rc < 0:
        raise ExchangeError("sad_exchange error %d: %s" % (rc, lib().sad_exchange_last_error().decode()))
    return rc


def nccl_version():
    return _check(lib().sad_exchange_nccl_version())


def gather_supported():
    """True when the resolved NCCL has what the copy-engine form needs (>= 2.28: symmetric windows, zero-CTA all-gather)."""
    return bool(lib().sad_exchange_gather_supported())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def slot_sum(slots, world, out):
    """out[i] = slots[0, i] + slots[1, i] + ... in that order (the local half of the copy-engine exchange).  slots: (world, n) fp32."""
    if not (slots.is_cuda and out.is_cuda and slots.dtype == out.dtype == torch.float32 and slots.dim() == 2 and slots.stride(1) == 1
            and out.is_contiguous() and slots.shape[0] == world and out.numel() <= slots.shape[1]):
        raise ValueError("slot_sum: (world, n) fp32 CUDA slots with unit inner stride and a contiguous fp32 output of <= n elements")
    rc = lib().sad_exchange_slot_sum_f32(C.c_void_p(slots.data_ptr()), slots.stride(0), int(world), C.c_void_p(out.data_ptr()), out.numel(), _stream())
    if rc != 0:
        raise ExchangeError("sad_exchange_slot_sum_f32: cudaError %d" % rc)
    return out


class NativeGradientExchange:
    """SUM-allreduce of one flat fp32 CUDA gradient buffer through libsad_exchange.so, whole or in buckets.

    Same interface as parallel.GradientExchange (allreduce / nbytes / bus_bytes) plus reduce_bucket / join for the overlapped
    form.  `world` ranks must construct it collectively (the NCCL id travels through torch.distributed once)."""

    def __init__(self, flat_grads, world=1, rank=0, group=None, max_ctas=0, gather=None, max_buckets=16):
        """gather: True = the copy-engine form (all-gather on the copy engines + a local rank-ordered sum; NCCL >= 2.28), False = the
        ncclAllReduce form, None = SAD_EXCHANGE_GATHER from the environment ("1" / "0"; default: the ncclAllReduce form).  Asking for
        the copy-engine form where the library cannot provide it raises — nothing is substituted silently."""
        if not (flat_grads.is_cuda and flat_grads.dtype == torch.float32 and flat_grads.is_contiguous() and flat_grads.dim() == 1):
            raise ValueError("the gradient buffer must be a contiguous 1-D fp32 CUDA tensor")
        self.flat, self.world, self.rank = flat_grads, int(world), int(rank)
        self.nbytes = flat_grads.numel() * 4
        self.calls = 0
        if gather is None:
            gather = os.environ.get("SAD_EXCHANGE_GATHER", "0") not in ("0", "", "false", "no")
        self.gather = bool(gather) and self.world > 1
        uid = (C.c_char * ID_BYTES)()
        if self.world > 1:
            import torch.distributed as dist
            if rank == 0:
                _check(lib().sad_exchange_unique_id(uid))
            t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
            dev = flat_grads.device if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = t.to(dev)
            dist.broadcast(t, src=0, group=group)
            uid = (C.c_char * ID_BYTES).from_buffer_copy(bytes(t.cpu().numpy().tobytes()))
        self.handle = C.c_void_p()
        with torch.cuda.device(flat_grads.device):
            if self.gather:
                capacity = flat_grads.numel() + 128 * int(max_buckets)
                _check(lib().sad_exchange_create_gather(uid, self.rank, self.world, capacity, C.byref(self.handle)))
            else:
                _check(lib().sad_exchange_create_config(uid, self.rank, self.world, int(max_ctas), C.byref(self.handle)))
        self.max_ctas = int(lib().sad_exchange_max_ctas(self.handle))
        self.mode = "copy-engine all-gather (NCCL zero-CTA, symmetric window) + local rank-ordered sum" if self.gather else "ncclAllReduce"

    def close(self):
        """Destroy the communicator.  NCCL keeps a reference for every CUDA graph that captured one of its collectives and
        ncclCommDestroy WAITS until those graphs are gone (measured: an exchange closed while its step's graph was still alive
        hung the process at exit) — destroy the graphs first (FullDistillStep.close does)."""
        if getattr(self, "handle", None):
            lib().sad_exchange_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        # deliberately not close(): object destruction order at interpreter exit is arbitrary, and tearing the communicator down
        # before a graph that references it blocks forever; an exchange that was not closed explicitly is left to process exit
        pass

    # ---- whole buffer on the caller's stream (no overlap) ----
    def allreduce(self, async_op=False):
        self.calls += 1
        _check(lib().sad_exchange_allreduce_f32(self.handle, C.c_void_p(self.flat.data_ptr()), self.flat.numel(), _stream()))
        return None

    # ---- overlapped form ----
    def reduce_bucket(self, begin, end):
        """Enqueue the exchange of flat[begin:end) behind everything enqueued so far on the current stream; returns at once."""
        if not (0 <= begin <= end <= self.flat.numel()):
            raise ValueError("bucket [%d, %d) outside the buffer" % (begin, end))
        _check(lib().sad_exchange_allreduce_async_f32(self.handle, C.c_void_p(self.flat.data_ptr() + 4 * begin), end - begin, _stream()))

    def plan_reset(self):
        """Forget the buckets announced during a previous graph capture (call before capturing the step again)."""
        _check(lib().sad_exchange_plan_reset(self.handle))

    def planned(self):
        return int(lib().sad_exchange_planned(self.handle))

    def flush(self):
        """After a launch of a graph that captured reduce_bucket calls: run those buckets' exchanges beside the graph, each behind
        the event its bucket recorded in that launch."""
        _check(lib().sad_exchange_flush(self.handle))

    def join(self):
        """The current stream waits for every bucket enqueued since the last join."""
        self.calls += 1
        _check(lib().sad_exchange_join(self.handle, _stream()))

    def bus_bytes(self):
        return 0 if self.world == 1 else 2.0 * (self.world - 1) / self.world * self.nbytes

    def stats(self):
        return {"buckets": int(lib().sad_exchange_buckets(self.handle)), "bytes": int(lib().sad_exchange_bytes(self.handle)),
                "gathered_buckets": int(lib().sad_exchange_gathered(self.handle)), "mode": self.mode}
