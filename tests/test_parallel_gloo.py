"""Host-side data-parallel logic on CPU: world_size-2 gloo processes (no GPU).  Mirrors
caffe2/caffe2/contrib/nccl/nccl_ops_test.py:56-79 (allreduce compared bit-exactly with the numpy sum)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from sad_b200 import parallel
    r, lr, w = parallel.init_process_group(backend="gloo")
    assert (r, w) == (rank, world)
    rng = np.random.default_rng(100 + rank)
    flat = torch.from_numpy(rng.standard_normal(n).astype(np.float32))
    ex = parallel.GradientExchange(flat)
    assert ex.world == world and ex.nbytes == 4 * n
    ex.allreduce()
    ref = sum(np.random.default_rng(100 + k).standard_normal(n).astype(np.float32) for k in range(world))
    ok = flat.numpy().tobytes() == ref.astype(np.float32).tobytes()     # bit exact (2 ranks: one fp32 add per element)
    # loss pre-scaling makes the sum a mean: per-rank gradient of (scale * local loss) summed == gradient of the global mean
    scale = parallel.distill_loss_scale(1.0, world)
    g = torch.full((8,), float(rank + 1)) * scale
    parallel.GradientExchange(g).allreduce()
    ok = ok and torch.allclose(g, torch.full((8,), sum(range(1, world + 1)) / world))
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_exchange_world2_gloo():
    world, n = 2, 6463220 // 8     # an eighth of the head's parameter count keeps the CPU test quick
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_images():
    from sad_b200 import parallel
    for gb, world in ((16, 8), (8, 8), (16, 1), (10, 4), (3, 4)):
        spans = [parallel.shard_images(gb, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        parallel.shard_images(8, 2, 2)


def test_loss_scale_and_single_process_exchange():
    from sad_b200 import parallel
    assert parallel.distill_loss_scale(1.0, 8) == 0.125          # retinanet_heads.py:342 with T = 1, 8 GPUs
    assert parallel.distill_loss_scale(2.0, 4) == 1.0
    g = torch.arange(5, dtype=torch.float32)
    ex = parallel.GradientExchange(g, world=1)
    assert ex.allreduce() is None and ex.bus_bytes() == 0 and torch.equal(g, torch.arange(5, dtype=torch.float32))
    with pytest.raises(ValueError):
        parallel.GradientExchange(torch.zeros(4, dtype=torch.float64), world=1)
    assert parallel.GradientExchange(torch.zeros(1024), world=8).bus_bytes() == 2 * 7 / 8 * 4096


def test_cpulist_parser_and_numa_binding_is_harmless_without_a_gpu():
    from sad_b200 import parallel
    assert parallel._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert parallel._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    msg = parallel.bind_to_gpu_numa_node(0)       # no CUDA device here: must report, not raise, and change nothing
    assert isinstance(msg, str) and os.sched_getaffinity(0) == before
