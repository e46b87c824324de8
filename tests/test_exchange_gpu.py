"""GPU: the native gradient exchange (libsad_exchange.so, include/sad_exchange.h) on one device — stream / event ordering of the
bucketed form, capture into a CUDA graph, and the bucket layout of the full step.  The multi-rank sum itself is checked under
torchrun by scripts/exchange_check.py (2+ GPUs: gpurun --gpus 2) and by bench.py's multi_gpu_check, and on CPU by the world-2
gloo tests (tests/test_parallel_gloo.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_nccl_is_resolved_at_run_time():
    from sad_b200 import exchange
    v = exchange.nccl_version()
    assert v >= 20000, v


def test_single_rank_buckets_are_ordered_behind_the_producer_and_ahead_of_the_consumer():
    from sad_b200 import exchange
    n = 1 << 22
    flat = torch.zeros(n, device="cuda")
    ex = exchange.NativeGradientExchange(flat, world=1, rank=0)
    big = torch.randn(4096, 4096, device="cuda")
    for it in range(3):
        flat.zero_()
        (big @ big).sum()                       # keep the producer stream busy so that ordering, not luck, decides
        flat[: n // 2].fill_(float(it + 1))     # bucket 0 is produced ...
        ex.reduce_bucket(0, n // 2)             # ... and handed over
        flat[n // 2:].fill_(2.0 * (it + 1))
        ex.reduce_bucket(n // 2, n)
        ex.join()                               # consumer (this stream) waits for both
        total = flat.sum()
        torch.cuda.synchronize()
        assert float(total) == (n // 2) * (it + 1) + (n // 2) * 2.0 * (it + 1)
    assert ex.stats()["buckets"] == 6 and ex.stats()["bytes"] == 3 * 4 * n
    ex.allreduce()                              # un-overlapped form: identity at world 1
    ex.close()


def test_buckets_announced_inside_a_cuda_graph_run_beside_it():
    """reduce_bucket on a stream that is being captured leaves an external event-record node in the graph and a planned bucket;
    flush() after every replay runs the exchange beside the graph, ordered behind that replay's events; join() orders the consumer."""
    from sad_b200 import exchange
    n = 1 << 20
    flat = torch.zeros(n, device="cuda")
    src = torch.arange(n, device="cuda", dtype=torch.float32)
    ex = exchange.NativeGradientExchange(flat, world=1, rank=0)
    big = torch.randn(2048, 2048, device="cuda")

    def step():
        flat.copy_(src)
        (big @ big).sum()
        ex.reduce_bucket(0, n // 4)
        flat[n // 4:].mul_(2.0)
        ex.reduce_bucket(n // 4, n)
        ex.join()

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()                      # eager: the buckets go out at once
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    eager_buckets = ex.stats()["buckets"]
    assert eager_buckets == 2 and ex.planned() == 0
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    assert ex.planned() == 2 and ex.stats()["buckets"] == eager_buckets      # nothing was sent during capture
    for it in range(3):
        flat.zero_()
        g.replay()
        ex.flush()
        ex.join()
        flat.add_(1.0)              # consumer on this stream, behind the join
        torch.cuda.synchronize()
        ref = src.clone()
        ref[n // 4:] *= 2.0
        ref += 1.0
        assert torch.equal(flat, ref)
    assert ex.stats()["buckets"] == eager_buckets + 6
    ex.plan_reset()
    assert ex.planned() == 0
    del g
    ex.close()


def test_full_step_buckets_cover_the_flat_buffer_and_overlap_changes_nothing():
    from sad_b200.full_step import FullDistillStep
    kw = dict(n_images=1, scale_px=(128, 256), student_blocks=(1, 1, 1, 1), teacher_blocks=(1, 1, 1, 1), seed=7)
    a = FullDistillStep(overlap_exchange=True, **kw)
    n = a.flat_grads.numel()
    spans = sorted(r for rs in a.buckets.values() for r in rs)
    assert spans[0][0] == 0 and spans[-1][1] == n and all(x[1] == y[0] for x, y in zip(spans, spans[1:]))
    assert a.buckets["head"] == [(0, a.n_head)]
    a.forward_backward()
    torch.cuda.synchronize()
    ga = a.flat_grads.clone()
    assert a.exchange.stats()["buckets"] == 0      # world 1: nothing is sent ...
    b = FullDistillStep(overlap_exchange=False, **kw)
    b.forward_backward()
    torch.cuda.synchronize()
    # ... and closing the buckets early (per-stage multi-tensor folds) gives the same gradients as one fold at the end
    assert float((ga - b.flat_grads).abs().max()) <= 1e-5 * float(b.flat_grads.abs().max())
    # graph capture with the hooks in place
    assert a.capture(), getattr(a, "capture_error", None)
    a.run()
    torch.cuda.synchronize()
    assert float((a.flat_grads - ga).abs().max()) <= 2e-3 * float(ga.abs().max())


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("n,offset", [(1 << 20, 0), (100003, 0), (4099, 1), (3, 0), (70001, 3)])
def test_slot_sum_adds_the_ranks_slots_in_rank_order(world, n, offset):
    """The local half of the copy-engine exchange (csrc/exchange/slot_sum.cu): bit-identical to ((s0 + s1) + s2) + ... in fp32, for
    vectorised, ragged and misaligned buckets; elements beyond the bucket stay untouched."""
    from sad_b200 import exchange
    g = torch.Generator(device="cuda").manual_seed(17 * world + n)
    stride = (n + 127) // 128 * 128
    slots = torch.randn(world, stride, device="cuda", generator=g) * 100.0
    backing = torch.full((n + 8,), -7.0, device="cuda")
    out = backing[offset: offset + n]
    exchange.slot_sum(slots, world, out)
    ref = slots[0, :n].clone()
    for r in range(1, world):
        ref = ref + slots[r, :n]
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    assert bool((backing[:offset] == -7.0).all()) and bool((backing[offset + n:] == -7.0).all())
    # and against the fp64 sum: fp32 round-off only
    ref64 = slots[:, :n].double().sum(0)
    assert float((out.double() - ref64).abs().max()) <= 1e-6 * 100.0 * world * 4


def test_copy_engine_form_is_reported_not_substituted():
    """The copy-engine form is a property of the resolved NCCL (>= 2.28); at world 1 there is nothing to gather and the flag is off."""
    from sad_b200 import exchange
    assert isinstance(exchange.gather_supported(), bool)
    flat = torch.ones(1024, device="cuda")
    ex = exchange.NativeGradientExchange(flat, world=1, rank=0, gather=True)
    assert ex.gather is False and ex.mode == "ncclAllReduce"
    ex.reduce_bucket(0, 1024)
    ex.join()
    torch.cuda.synchronize()
    assert float(flat.sum()) == 1024.0 and ex.stats()["gathered_buckets"] == 0
    ex.close()
