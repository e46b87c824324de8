"""Data-parallel plumbing of the path: one process per GPU, image-batch sharding, ONE gradient allreduce.

Reference: detectron/lib/modeling/optimizer.py:33-92 builds one replica of the graph per GPU inside a single
process, pre-scales every loss by 1 / NUM_GPUS (detector.py:650-655; for the distillation loss through the op
argument scale = T^2 / NUM_GPUS, retinanet_heads.py:342) and sums the gradients of each parameter blob over the
GPUs with one NCCLAllreduce op per blob (or the muji Add/Copy tree, muji.py:123-180).  Here: torch.distributed
(NCCL over NVLink/NVSwitch on the GPUs; gloo in the CPU tests) and a single collective over the flat gradient
buffer the head owns.  Nothing on the loss path needs a collective: PowSum's normaliser and every loss are
computed per GPU over that GPU's images (optimizer.py:62-69).
"""
import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, local_rank, world) as torchrun exports them; (0, 0, 1) for a plain `python` launch."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend=None, device=None):
    """Joins the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT.  Returns (rank, local_rank, world)."""
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs that are local to the GPU's PCIe root (sysfs local_cpulist) BEFORE host
    buffers are allocated and pinned, so their pages land on the GPU's own NUMA node.  With one process per GPU and
    every rank streaming host buffers over PCIe (the end-to-end path: 237 MB per step and GPU), leaving all ranks on
    one node funnels all traffic through that node's memory controllers and the inter-socket link.
    Returns a short description for the bench record; never fails (returns the reason instead)."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/local_cpulist") as f:
            local = _parse_cpulist(f.read())
        node = "?"
        try:
            with open(base + "/numa_node") as f:
                node = f.read().strip()
        except OSError:
            pass
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        if not cpus:
            return "numa: no local cpu of %s is in this process's affinity mask" % bdf
        os.sched_setaffinity(0, cpus)
        return "numa node %s (%d local cpus of %s)" % (node, len(cpus), bdf)
    except Exception as e:  # sysfs not exposed in a container, no such attribute, ...
        return "numa: not bound (%s: %s)" % (type(e).__name__, e)


def shard_images(global_batch, world, rank):
    """[begin, end) of the images rank `rank` owns: contiguous, sizes differing by at most one (TRAIN.IMS_PER_BATCH
    images per GPU in the reference, config.py:96; the remainder goes to the lowest ranks)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, rem = divmod(int(global_batch), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def distill_loss_scale(temperature, world):
    """The `scale` argument of SigmoidAdaptiveDistillLoss(+Gradient): T^2 / NUM_GPUS (retinanet_heads.py:342)."""
    return float(temperature) ** 2 / int(world)


class GradientExchange:
    """SUM-allreduce of one flat fp32 gradient buffer: the only exchange of the data-parallel step."""

    def __init__(self, flat_grads, world=None, group=None):
        if flat_grads.dtype != torch.float32 or not flat_grads.is_contiguous() or flat_grads.dim() != 1:
            raise ValueError("the gradient buffer must be a contiguous 1-D fp32 tensor")
        self.flat, self.group = flat_grads, group
        self.world = (dist.get_world_size(group) if dist.is_initialized() else 1) if world is None else int(world)
        self.nbytes = flat_grads.numel() * 4
        self.calls = 0

    def allreduce(self, async_op=False):
        """In place.  With the losses pre-scaled by 1 / world the sum is the mean gradient of the global batch."""
        self.calls += 1
        if self.world == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def bus_bytes(self):
        """Bytes each rank moves for one ring/NVLS allreduce: 2 (world - 1) / world x payload (NCCL's busbw convention)."""
        return 0 if self.world == 1 else 2.0 * (self.world - 1) / self.world * self.nbytes


def make_exchange(flat_grads, world=1, rank=0, group=None):
    """The exchange object of a step: on CUDA buffers the native one (libsad_exchange.so: host C++ over NCCL, bucketed and
    event-ordered so that it overlaps the backward pass — include/sad_exchange.h); on CPU buffers (the gloo tests) the
    torch.distributed form above.  Same interface: allreduce(), nbytes, bus_bytes(); the native one adds reduce_bucket / join."""
    if flat_grads.is_cuda:
        from .exchange import NativeGradientExchange
        return NativeGradientExchange(flat_grads, world=world, rank=rank, group=group)
    return GradientExchange(flat_grads, world=world, group=group)
