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
