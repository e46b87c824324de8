"""Multi-GPU check of the native gradient exchange in the style of caffe2/caffe2/contrib/nccl/nccl_ops_test.py:56-79:
every rank's result equals the sum over ranks (fp64 reference, fp32 round-off; bit-exact at 2 ranks) and all ranks hold
bit-identical results, for the whole-buffer call, for the bucketed overlapped form, and inside a CUDA graph.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/exchange_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sad_b200 import exchange
    n = 6463220 + 31233280 // 8          # the head's parameters + a slice of the body's
    n -= n % 4

    def mine(r, salt):
        g = torch.Generator(device="cuda").manual_seed(1000 * salt + r)
        return torch.randn(n, device="cuda", generator=g)

    def reference(salt):
        acc = torch.zeros(n, device="cuda", dtype=torch.float64)
        for r in range(world):
            acc += mine(r, salt).double()
        return acc

    def rank_ordered(salt):
        acc = mine(0, salt)
        for r in range(1, world):
            acc = acc + mine(r, salt)
        return acc

    ex = None
    results = {}
    gather = os.environ.get("SAD_EXCHANGE_GATHER", "0") == "1"    # the copy-engine form (NativeGradientExchange reads the same variable)
    parts = os.environ.get("SAD_EXCHANGE_CHECK_PARTS", "whole,buckets,graph,ragged,activity").split(",")
    for name in parts:
        print("[rank %d] part %s" % (rank, name), flush=True)
        salt = {"whole": 1, "buckets": 2, "graph": 3, "ragged": 4, "activity": 5}[name]
        flat = mine(rank, salt)
        if ex is None:
            print("[rank %d] creating the exchange" % rank, flush=True)
            ex = exchange.NativeGradientExchange(flat, world=world, rank=rank)
            print("[rank %d] exchange created (NCCL %d)" % (rank, exchange.nccl_version()), flush=True)
        else:
            ex.flat = flat
        cuts = [0, n // 5, n // 2, n]
        if name == "ragged":
            cuts = [0, n // 5 + 1, n // 2 + 3, n - 2]     # bucket starts / lengths that are not multiples of 4 floats; the last 2 stay local
        if name == "whole":
            ex.allreduce()
        elif name == "activity":
            # what ran on the GPU for one bucketed exchange: NCCL kernels (SMs) or copy-engine memcpys + the slot-sum kernel
            from torch.profiler import ProfilerActivity, profile
            for lo, hi in zip(cuts, cuts[1:]):
                ex.reduce_bucket(lo, hi)
            ex.join()
            torch.cuda.synchronize()
            flat.copy_(mine(rank, salt))
            with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
                for lo, hi in zip(cuts, cuts[1:]):
                    ex.reduce_bucket(lo, hi)
                ex.join()
                torch.cuda.synchronize()
            seen = {}
            try:
                for e in prof.events():
                    if e.device_type == torch.autograd.DeviceType.CUDA:
                        k = seen.setdefault(e.name[:90], [0, 0.0])
                        k[0] += 1
                        k[1] += getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0.0)
                results["activity_rank0"] = {k: {"count": v[0], "us": round(v[1], 1)} for k, v in sorted(seen.items(), key=lambda kv: -kv[1][1])}
            except Exception as exc:   # the listing is evidence, not a check
                results["activity_rank0"] = {"error": repr(exc)}
        elif name in ("buckets", "ragged"):
            for lo, hi in zip(cuts, cuts[1:]):
                ex.reduce_bucket(lo, hi)
            ex.join()
        else:
            keep = flat.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for lo, hi in zip(cuts, cuts[1:]):
                    ex.reduce_bucket(lo, hi)
                ex.join()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            flat.copy_(keep)
            ex.plan_reset()
            big = torch.randn(4096, 4096, device="cuda")
            (big @ big).sum()                   # cuBLAS initialises its handle / workspace outside the capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(4):
                    (big @ big).sum()           # ~1 ms of captured work in front of the producer ...
                flat.copy_(keep)                # ... which only then writes the buckets' contents
                for lo, hi in zip(cuts, cuts[1:]):
                    ex.reduce_bucket(lo, hi)
                ex.join()
            assert ex.planned() == 3            # captured buckets stay out of the graph: event-record nodes + a plan
            for _ in range(2):
                flat.zero_()                    # an exchange that started before its bucket's event would sum these zeros
                g.replay()
                ex.flush()                      # the exchanges run beside the graph, behind the events of this launch
                ex.join()
                torch.cuda.synchronize()
            del g
        torch.cuda.synchronize()
        ref = reference(salt)
        err = float((flat.double() - ref).abs().max() / ref.abs().max())
        if name == "ragged":
            assert torch.equal(flat[n - 2:], mine(rank, salt)[n - 2:])       # outside every bucket: untouched
            flat[n - 2:] = rank_ordered(salt)[n - 2:]
            ref[n - 2:] = flat[n - 2:].double()
            err = float((flat.double() - ref).abs().max() / ref.abs().max())
        # the copy-engine form adds the ranks' slots in rank order at every world size; ncclAllReduce is only known to at 2 ranks
        exact2 = bool(torch.equal(flat, rank_ordered(salt))) if (world == 2 or (gather and name != "whole")) else None
        # all ranks bit-identical: MAX and MIN over ranks of the int32 view agree
        bits = flat.view(torch.int32)
        hi_, lo_ = bits.clone(), bits.clone()
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        same = bool(torch.equal(hi_, lo_))
        results[name] = {"rel_err_vs_fp64_sum": err, "bit_exact_vs_rank_ordered_sum": exact2, "ranks_bit_identical": same}
        assert err < 1e-6 and same and exact2 in (None, True), (name, results[name])
    if rank == 0:
        print(json.dumps({"exchange_check": "ok", "world": world, "elements": n, "nccl": exchange.nccl_version(), "mode": ex.mode, "results": results,
                          "stats": ex.stats()}))
    ex.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
