#!/usr/bin/env python
"""bench.py — distillation-loss+grad M-anchors/s on BASELINE.json configs[1].

A "step" is one pass of the hot path over one batch of synthetic input: PowSum over the teacher
probabilities of all 5 FPN levels (-> adaptive normaliser) followed by the fused
SigmoidAdaptiveDistillLoss + Gradient over all 5 levels (bs = 2, 600 px: 245 520 anchors,
19 641 600 logits), through the C ABI of include/sad_b200.h.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); the path shards by image, so every rank runs its
own bs = 2 batch (weak scaling) and no data-path collective is involved in the loss step.
`--impl reference` times the CPU restatement of the reference operators (oracle/, all host
threads) on a bounded sample of the same workload; rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "distillation-loss+grad M-anchors/sec"
UNIT = "M-anchors/s"
HEAD = dict(gamma=2.0, alpha=0.5, beta=0.0, scale=1.0, num_classes=80, ignored_label=-1)
POWER = 1.8
BYTES_PER_ELEMENT = 12.05 + 4.0  # loss+grad X 4 + T 4 + dX 4 + label 4/80, plus PowSum T 4 (same launch)
WORKLOAD = "configs[1]: PowSum(5 levels, power 1.8) + fused SigmoidAdaptiveDistillLoss+Gradient, 5 FPN levels, bs=2, 600px (245520 anchors), one launch"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic_per_launch():
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "distill_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period_s=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if not self.nv:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(self.period)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def cpu_reference_pass(levels, cpu_oracle):
    """One pass of the reference algorithm on the CPU: PowSum, then per level loss forward +
    gradient (the reference runs them as separate operators)."""
    wp = cpu_oracle.pow_sum([l[1] for l in levels], POWER)
    out = []
    for (x, t, g) in levels:
        lo = cpu_oracle.distill_loss(x, t, g, wp, **HEAD)
        gr = cpu_oracle.distill_grad(x, t, g, wp, d_loss=1.0, **HEAD)
        out.append((lo, gr))
    return wp, out


def operator_net_ms(lib_path, host_levels, fuse=False, runs=5):
    """the reference's OWN CUDA operators (pow_sum_op.cu and
    sigmoid_adaptive_distillation_loss_op.cu compiled unmodified into oracle/_ref/libref_ops.so,
    with its single-block math::Sum, temp buffers and extra Scale passes) run as the reference
    graph builds them (PowSum, 5 x loss, 5 x gradient) on the same inputs on this GPU.  Wall time
    around RunNet with a device synchronise on both sides (the reference ops synchronise
    themselves).  The same NetDef is also run through the product's operator library
    (lib_path=None), as built and after its FuseAdaptiveDistillOps graph pass."""
    import torch
    from sad_b200 import c2, retinanet_heads
    lib = c2.OperatorLibrary(lib_path) if lib_path else c2.OperatorLibrary()
    net, _, _ = retinanet_heads.add_distill_loss(gpu_id=0, num_gpus=1)
    text = net.to_text()
    if fuse:
        text, _ = lib.FuseAdaptiveDistillOps(text)
    ws = lib.Workspace()
    dev = [tuple(torch.from_numpy(a).cuda() for a in l) for l in host_levels]
    retinanet_heads.feed_level_blobs(ws, 0, dev)
    ws.CreateNet(text)
    for _ in range(2):
        ws.RunNet(net.name)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(runs):
        ws.RunNet(net.name)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / runs * 1e3


def run_reference_arm(args, rank):
    if rank != 0:
        return
    import numpy as np
    from oracle import cpu_oracle
    from sad_b200 import synthetic
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the reference arm runs on rank 0 alone and takes every host core it may use
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    cpu_oracle.set_num_threads(max(1, avail))
    cores = cpu_oracle.num_threads()
    full = synthetic.make_pyramid(1234, 2, 600)
    # bounded sample: the largest suffix of the pyramid (P3..P7, P4..P7, ...) whose pass keeps the
    # whole K + W run within ~2.5 minutes on this host
    t0 = time.perf_counter()
    cpu_reference_pass(full[2:], cpu_oracle)
    t_small = time.perf_counter() - t0
    per_anchor = t_small / synthetic.anchors_in(full[2:])
    budget = 150.0 / max(1, args.steps + args.warmup)
    first = 0
    while first < 4 and per_anchor * synthetic.anchors_in(full[first:]) > budget:
        first += 1
    sample = full[first:]
    anchors = synthetic.anchors_in(sample)
    for _ in range(args.warmup):
        cpu_reference_pass(sample, cpu_oracle)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_pass(sample, cpu_oracle)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = anchors / dt / 1e6
    sample_desc = "FPN levels P%d..P7 of configs[1] (bs=2, 600px): %d anchors per step, PowSum + loss fwd + grad" % (3 + first, anchors)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample_desc,
                   "note": "the reference has no CPU implementation of these ops (CAFFE_NOT_IMPLEMENTED); this is the "
                           "CPU restatement of its CUDA kernels in oracle/, OpenMP over all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def multi_gpu_check(flat_params, exchange, world, rank):
    """caffe2/caffe2/contrib/nccl/nccl_ops_test.py:56-79 on the live job: (1) the exchange's SUM equals the fp64 sum over ranks of
    rank-seeded buffers (fp32 round-off; bit-exact against the rank-ordered fp32 sum at 2 ranks) and every rank holds the same
    bits; (2) after the K optimiser steps just timed, all replicas' parameters are bit-identical."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return "n/a (1 GPU)"
    n = min(exchange.flat.numel(), 1 << 22)

    def seeded(r):
        g = torch.Generator(device="cuda").manual_seed(4242 + r)
        return torch.randn(n, device="cuda", generator=g)

    keep = exchange.flat[:n].clone()
    exchange.flat[:n].copy_(seeded(rank))
    if hasattr(exchange, "reduce_bucket"):
        exchange.reduce_bucket(0, n // 3)
        exchange.reduce_bucket(n // 3, n)
        exchange.join()
    else:
        exchange.allreduce()
    got = exchange.flat[:n].clone()
    exchange.flat[:n].copy_(keep)
    ref = torch.zeros(n, device="cuda", dtype=torch.float64)
    for r in range(world):
        ref += seeded(r).double()
    err = float((got.double() - ref).abs().max() / ref.abs().max())

    def same_on_all_ranks(t):
        bits = t.contiguous().view(torch.int32)
        hi, lo = bits.clone(), bits.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        return bool(torch.equal(hi, lo))

    ok = err < 1e-6 and same_on_all_ranks(got)
    if world == 2:
        ok = ok and bool(torch.equal(got, seeded(0) + seeded(1)))
    replicas = same_on_all_ranks(flat_params)
    flag = torch.tensor([1 if (ok and replicas) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return "ok" if int(flag.item()) == 1 else "FAILED (sum rel err %.3g, sum identical on ranks / replicas identical: %s / %s)" % (err, ok, replicas)


def _tensor_peak():
    """tf32 tensor peak for a kernel timed inside a long step: half the measured sustained bf16 rate (tf32 runs at
    half the bf16 rate on this part: 1.1 vs 2.25 PFLOP/s nominal, B200_PROFILING.md); MEASURED_PEAKS.json holds no tf32 figure."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d.get("bf16_tflops_sustained", d["bf16_tflops"])) / 2.0, "measured bf16_tflops_sustained / 2 (tf32 = half the bf16 rate)"
        except Exception:
            pass
    return 1400.0 / 2.0, "fallback: 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md) / 2"


def run_head_step(args, rank, world, barrier, native, f16=False, x3=False):
    """the widened path on every rank's own 2-image shard — student head forward, PowSum, fused
    distillation loss + gradient, head backward (device work replayed from ONE CUDA graph), then the step's only
    collective (SUM-allreduce of the 25.9 MB flat head-gradient buffer) and the momentum-SGD update (one launch)."""
    import torch
    import torch.distributed as dist
    from sad_b200.step import DistillHeadStep

    K = args.head_steps or min(args.steps, 100)
    st = DistillHeadStep(n_images=2, scale_px=600, world=world, rank=rank, compute_f16=f16, compute_f32x3=x3)
    n0 = native.lib().sad_launch_count()
    st.forward_backward()
    per_step = int(native.lib().sad_launch_count() - n0)
    st.capture()

    def one():
        st.run()
        st.allreduce()
        st.sgd()

    for _ in range(5):
        one()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ar = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    ev0.record()
    for i in range(K):
        st.run()
        ar[i][0].record()
        st.allreduce()
        ar[i][1].record()
        st.sgd()
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1) / K, sum(a.elapsed_time(b) for a, b in ar) / K], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_ms = float(t[0].item()), float(t[1].item())
    losses = st.losses()
    assert all(l == l and abs(l) < 1e30 for l in losses), ("non-finite distillation loss", losses)
    mgc = multi_gpu_check(st.head.flat_params, st.exchange, world, rank)
    assert not mgc.startswith("FAILED"), mgc
    fwd_f, bwd_f = st.flops()
    dev_ms = ms - ar_ms
    peak, src = _tensor_peak()
    if f16:   # 16-bit operands: the measured sustained bf16 rate itself
        peak, src = peak * 2.0, src.replace(" / 2 (tf32 = half the bf16 rate)", " (16-bit operands)").replace(" / 2", "")
    achieved = (fwd_f + bwd_f) / (ms * 1e-3) / 1e12
    kind = "f16" if f16 else ("tf32 x3 passes (3xTF32, fp32-accurate)" if x3 else "tf32")
    if x3:   # three tensor-core passes per product: the tensor work is 3x the algorithmic flops
        achieved *= 3.0
    out = {
        "metric": "RetinaNet head distill-step imgs/sec", "value": world * st.images / (ms * 1e-3), "unit": "imgs/s",
        "ms_per_step": ms, "steps": K, "images_per_gpu": st.images, "scaling": "weak",
        "workload": "student head fwd (10 convs, 5 levels) + PowSum + fused distill loss+grad + head bwd (dgrad + wgrad) at bs=2/GPU, "
                    "600px, then ONE allreduce of %d head-gradient bytes + momentum SGD" % st.exchange.nbytes,
        "allreduce_ms": ar_ms, "allreduce_bytes": st.exchange.nbytes,
        "allreduce_busbw_gbs": (st.exchange.bus_bytes() / (ar_ms * 1e-3) / 1e9) if world > 1 and ar_ms > 0 else None,
        "exchange": "libsad_exchange.so (host C++ over NCCL), whole flat buffer on the step's stream", "multi_gpu_check": mgc,
        "conv_gflop_per_step": (fwd_f + bwd_f) / 1e9,
        "roofline": {"bound": "tensor", "kernel": "conv3x3_tf32_kernel + conv3x3_wgrad_tf32_kernel (tcgen05 kind::%s, 30 launches/step)" % kind,
                     "achieved": achieved, "peak": peak, "peak_source": src, "unit": "TFLOP/s", "frac": achieved / peak,
                     "note": "achieved = algorithmic conv flops of the step / WHOLE step time (loss kernels, layout passes, allreduce and "
                             "SGD included); device-only step %.3f ms" % dev_ms},
        "gpu_launches": per_step * (K + 5), "launches_per_step": per_step, "cuda_graph": True,
        "dtype": "%s operands, fp32 accumulate (convs); f32 (loss)" % kind, "distill_losses": losses,
    }
    st.close()
    return out


def run_full_step(args, rank, world, barrier, native, n_images=2, config5=False, teacher_f16=False, student_f16=False):
    """BASELINE.json configs[3] geometry (n_images = 2 per GPU) or configs[2] (n_images = 16 on one GPU): R-50-FPN student <-
    R-101-FPN teacher, full distillation training step, 600 px, one allreduce of the flat [head | body] gradient buffer.
    Heads, every loss, the exchange and the optimiser step are this repository's kernels; the ResNet/FPN bodies are
    PyTorch/cuDNN scaffolding (full_step.py)."""
    import torch
    import torch.distributed as dist
    from sad_b200.full_step import FullDistillStep

    K = args.full_steps or min(args.steps, 20)
    if n_images > 2:
        K = min(K, 10)
    if config5:   # configs[4]: R-101 student <- ResNeXt-101-64x4d teacher, 500 px, one image per GPU
        st = FullDistillStep(n_images=n_images, scale_px=500, world=world, rank=rank, student_blocks=(3, 4, 23, 3),
                             teacher_blocks=(3, 4, 23, 3), teacher_body=dict(groups=64, width_per_group=4, stride_1x1=False),
                             teacher_head_f16=teacher_f16, student_head_f16=student_f16)
    else:
        st = FullDistillStep(n_images=n_images, scale_px=600, world=world, rank=rank)
    for _ in range(3):
        st.step()
    n0 = native.lib().sad_launch_count()
    st.forward_backward()
    per_step = int(native.lib().sad_launch_count() - n0)

    def timed(overlap):
        """K steps of the captured step; returns (ms per step, ms between the end of the device work and the end of the
        exchange — 0 by construction when the exchange is inside the graph), max over ranks."""
        st.overlap_exchange = overlap
        graphed_ = st.capture()
        for _ in range(2):
            st.step()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ar = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        ev0.record()
        for i in range(K):
            st.run()
            ar[i][0].record()
            st.allreduce()
            ar[i][1].record()
            st.sgd()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1) / K, sum(a.elapsed_time(b) for a, b in ar) / K], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item()), graphed_

    # the reference's order first (optimizer.py:72-92: every allreduce after the whole backward), then the overlapped form
    ms_seq, ar_ms, _ = timed(False)
    ms, _, graphed = timed(True)
    launches = per_step * 2 * (K + 2)
    exposed = max(0.0, ms - (ms_seq - ar_ms))
    mgc = multi_gpu_check(st.flat_params, st.exchange, world, rank)
    assert not mgc.startswith("FAILED"), mgc
    losses = st.losses()
    assert all(v == v for v in [losses["normalizer"]] + losses["bbox"] + losses["distill"] + losses["focal"]), ("non-finite loss", losses)
    line = {
        "metric": "RetinaNet-R50 distill-step imgs/sec", "value": world * st.images / (ms * 1e-3), "unit": "imgs/s", "ms_per_step": ms,
        "steps": K, "images_per_gpu": st.images, "scaling": "weak",
        "workload": "R-50-FPN student <- R-101-FPN teacher (random init), 3x640x1024 synthetic images, bs=%d/GPU: teacher fwd, student " % n_images +
                    "fwd+bwd, focal + box + adaptive distillation losses, ONE allreduce of %d gradient bytes, momentum SGD" % st.exchange.nbytes,
        "baseline_config": "configs[3] (bs=2 per GPU)" if n_images == 2 else "configs[2] (bs=%d on one GPU)" % n_images,
        "native": "both RetinaNet heads forward + backward (tcgen05 tf32), PowSum + distillation loss/gradient (one cooperative launch), "
                  "SigmoidFocalLoss + gradient accumulated into the same d(logits), SelectSmoothL1Loss + gradient, teacher Sigmoid fused into its "
                  "prediction convolution, gradient exchange, momentum-SGD update (one launch over the flat buffers)",
        "scaffolding": "ResNet/FPN bodies on cuDNN (TF32) under autograd (SURVEY.md 8f rank 3)",
        "allreduce_ms": ar_ms, "allreduce_bytes": st.exchange.nbytes,
        "allreduce_busbw_gbs": (st.exchange.bus_bytes() / (ar_ms * 1e-3) / 1e9) if world > 1 and ar_ms > 0 else None,
        "allreduce_exposed_ms": exposed, "ms_per_step_exchange_after_backward": ms_seq,
        "exchange": "libsad_exchange.so (host C++ over NCCL): 4 buckets (head | res5 + FPN | res4 | res3 + biases) enqueued on the exchange's "
                    "stream as the backward pass completes them, inside the captured graph; value / ms_per_step are this overlapped form, "
                    "ms_per_step_exchange_after_backward + allreduce_ms the reference's order (one allreduce after the whole backward); "
                    "allreduce_exposed_ms = ms_per_step - (ms_per_step_exchange_after_backward - allreduce_ms)",
        "exchange_mode": getattr(st.exchange, "mode", None), "exchange_stats": st.exchange.stats() if hasattr(st.exchange, "stats") else None,
        "multi_gpu_check": mgc,
        "params": st.param_count(), "gpu_launches": launches, "native_launches_per_step": per_step, "cuda_graph": bool(graphed),
        "cuda_graph_error": getattr(st, "capture_error", None), "losses": losses,
    }
    if config5:
        line["metric"] = "RetinaNet-R101 distill-step imgs/sec"
        line["workload"] = ("R-101-FPN student <- ResNeXt-101-64x4d-FPN teacher (random init), 3x512x896 synthetic images (500 px scale), "
                            "bs=%d/GPU: teacher fwd, student fwd+bwd, focal + box + adaptive distillation losses, ONE allreduce of %d "
                            "gradient bytes, momentum SGD" % (n_images, st.exchange.nbytes))
        line["baseline_config"] = "configs[4] geometry and models (bs=1 per GPU); computed in tf32 / fp32, i.e. at higher precision than the fp16 the config names"
        line["teacher_head_dtype"] = "f16 operands, fp32 accumulate (tcgen05 kind::f16)" if teacher_f16 else "tf32"
        line["student_head_dtype"] = "f16 operands forward + backward, fp32 accumulate, loss-scaled fp16 gradient tensors" if student_f16 else "tf32"
        if teacher_f16 and student_f16:
            line["baseline_config"] = ("configs[4] geometry and models (bs=1 per GPU); both RetinaNet heads in mixed fp16 (fp16 operands, fp32 "
                                       "accumulation, fp32 losses and parameters) as the config names; the cuDNN bodies (scaffolding) stay tf32")
    st.close()     # graph first, then the exchange's communicator (NCCL waits for graphs that captured its collectives), then the heads
    del st
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="host-buffer steps (0 = min(steps, 20))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="env", choices=["env", "allreduce", "gather"],
                    help="gradient exchange of the step objects: ncclAllReduce buckets, or the copy-engine all-gather + local sum "
                         "(include/sad_exchange.h); env = SAD_EXCHANGE_GATHER decides (default allreduce)")
    ap.add_argument("--head-steps", type=int, default=0, help="head distillation steps (0 = min(steps, 100); -1 = skip)")
    ap.add_argument("--full-steps", type=int, default=0, help="full R-50 <- R-101 distillation steps (0 = min(steps, 20); -1 = skip)")
    ap.add_argument("--no-heads-f16", dest="teacher_f16", action="store_false", default=True,
                    help="skip configs[4]'s step with both RetinaNet heads on fp16 operands (object full_step_config5_heads_f16)")
    ap.add_argument("--teacher-f16", "--heads-f16", dest="teacher_f16", action="store_true", help="(default) run that object")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.exchange != "env":
        os.environ["SAD_EXCHANGE_GATHER"] = "1" if args.exchange == "gather" else "0"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU path for --impl b200"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    from sad_b200 import native, ops, parallel, synthetic
    numa = parallel.bind_to_gpu_numa_node(local_rank)  # before any pinned host allocation (the e2e path streams host buffers)

    # ---- inputs: this rank's shard (2 images), NSETS rotating copies so no step finds its inputs in L2
    NSETS = 3
    host = synthetic.make_pyramid(1234 + 1000 * rank, 2, 600)
    anchors = synthetic.anchors_in(host)
    elements = int(sum(l[0].size for l in host))
    sets = []
    for s in range(NSETS):
        if s == 0:
            dev = [tuple(torch.from_numpy(a).cuda() for a in l) for l in host]
        else:  # different values per set (a cheap device-side perturbation), same shapes
            dev = [(x + 0.01 * s, t.clone(), g.clone()) for (x, t, g) in sets[0]]
        sets.append(dev)
    plans = [ops.DistillPlan(dev, power=POWER, **HEAD) for dev in sets]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for i in range(args.warmup):
        plans[i % NSETS].run()
    barrier()

    # ---- timed region: K steps; the dominant kernel is bracketed by its own event pair
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.sample()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = native.lib().sad_launch_count()
    ev0.record()
    for i in range(args.steps):
        p = plans[i % NSETS]
        kev[i][0].record()
        p.run()  # ONE cooperative launch: PowSum -> grid barrier -> fused loss + gradient of all 5 levels
        kev[i][1].record()
    ev1.record()
    sampler.sample()
    barrier()
    sampler.stop_flag = True
    launches = native.lib().sad_launch_count() - launches0
    total_ms = ev0.elapsed_time(ev1)
    kernel_times = sorted(a.elapsed_time(b) for a, b in kev)
    kernel_ms = sum(kernel_times) / max(1, args.steps)
    kernel_ms_median = kernel_times[len(kernel_times) // 2]
    kernel_ms_min = kernel_times[0]
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / max(1, args.steps)
    value = world * anchors / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: host (pinned) buffers through sad_distill_step_host, copies inside the timed region
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    cpu = [tuple(torch.from_numpy(a).pin_memory() for a in l) for l in host]
    outs = [torch.empty_like(l[0]).pin_memory() for l in cpu]
    step = ops.HostStep(local_rank)
    step.bind(cpu, outs, power=POWER, **HEAD)
    for _ in range(3):
        step.run()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_losses, e2e_norm = step.run()
    torch.cuda.synchronize()
    e2e_dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_value = world * anchors / float(e2e_dt.item()) / 1e6
    h2d = int(sum(l[0].nbytes + l[1].nbytes + l[2].nbytes for l in host))
    d2h = int(sum(l[0].nbytes for l in host)) + 4 * (len(host) * 2 + 1)
    # the e2e path and the device path must agree (same kernels)
    plans[0].run()
    torch.cuda.synchronize()
    for a, b in zip(e2e_losses, plans[0].losses):
        assert abs(a - b.item()) <= 1e-5 * abs(b.item()), ("e2e/device loss mismatch", a, b.item())
    step.close()
    # the same call with the gradients LEFT ON THE DEVICE (d_logits = NULL: what a training step does — ConvGradient consumes
    # them there); only the losses and the normaliser are read back
    step2 = ops.HostStep(local_rank)
    step2.bind(cpu, None, power=POWER, **HEAD)
    for _ in range(3):
        step2.run()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step2.run()
    torch.cuda.synchronize()
    e2e_dt2 = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_dt2, op=dist.ReduceOp.MAX)
    e2e_value2 = world * anchors / float(e2e_dt2.item()) / 1e6
    step2.close()

    # ---- copy-only baseline of the e2e path: the SAME bytes over the same two directions (all inputs up on one stream, all
    # gradients down on another, concurrently), no kernels: what this host's PCIe / memory path allows for that step
    cp_in, cp_out = torch.cuda.Stream(), torch.cuda.Stream()
    dev_in = [tuple(torch.empty_like(a, device="cuda") for a in l) for l in cpu]
    dev_out = [torch.empty_like(l[0], device="cuda") for l in cpu]

    def copy_only():
        with torch.cuda.stream(cp_in):
            for l, d in zip(cpu, dev_in):
                for a, b in zip(l, d):
                    b.copy_(a, non_blocking=True)
        with torch.cuda.stream(cp_out):
            for o, d in zip(outs, dev_out):
                o.copy_(d, non_blocking=True)

    for _ in range(3):
        copy_only()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        copy_only()
    torch.cuda.synchronize()
    copy_dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(copy_dt, op=dist.ReduceOp.MAX)
    copy_only_ms = float(copy_dt.item()) * 1e3
    del dev_in, dev_out

    head_line = head_f16_line = head_x3_line = None
    if args.head_steps >= 0:
        head_line = run_head_step(args, rank, world, barrier, native)
        if args.teacher_f16:
            head_f16_line = run_head_step(args, rank, world, barrier, native, f16=True)
        if world == 1:   # the cost of the fp32-accurate mode (what matches the reference's fp32 convolution to 1e-4)
            head_x3_line = run_head_step(args, rank, world, barrier, native, x3=True)

    full_line = full16_line = full5_line = full5_f16_line = None
    if args.full_steps >= 0:
        full_line = run_full_step(args, rank, world, barrier, native)
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        full5_line = run_full_step(args, rank, world, barrier, native, n_images=1, config5=True)
        if args.teacher_f16:   # the same step with the forward-only teacher head on fp16 operands (configs[4]: mixed fp16 compute)
            gc.collect()
            torch.cuda.empty_cache()
            full5_f16_line = run_full_step(args, rank, world, barrier, native, n_images=1, config5=True, teacher_f16=True, student_f16=True)
        if world == 1:   # BASELINE.json configs[2]: the same step at bs = 16 on one GPU
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            full16_line = run_full_step(args, rank, world, barrier, native, n_images=16)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    achieved = BYTES_PER_ELEMENT * elements / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "anchors_per_gpu_step": anchors, "elements_per_gpu_step": elements,
                   "args": HEAD, "power": POWER,
                   "l2": "%d rotating input sets (%.0f MB) > 126 MB L2; one step touches 236 MB" % (NSETS, NSETS * 3 * elements * 4 / 1e6),
                   "parallelism": "image-sharded x%d, no data-path collective" % world, "host_binding": numa},
        "roofline": {"bound": "hbm", "kernel": "distill_fused_kernel<alpha=.5> (cooperative persistent TMA-ring: PowSum, grid barrier, 5-level loss+grad)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": _traffic_per_launch(), "algorithmic_bytes_per_launch": BYTES_PER_ELEMENT * elements,
                     "algorithmic_bytes_per_element": "16.05 = PowSum 4 (T) + loss+grad 12.05 (X 4 + T 4 + dX 4 + labels 4/80), SURVEY.md 8(d); "
                                                      "the second read of T can be served by L2, so achieved may exceed the DRAM copy peak",
                     "kernel_ms": kernel_ms, "kernel_ms_median": kernel_ms_median, "kernel_ms_min": kernel_ms_min,
                     "frac_median": BYTES_PER_ELEMENT * elements / (kernel_ms_median * 1e-3) / 1e9 / peak,
                     "timing": "one CUDA event pair around every launch of the timed region (%d steps): achieved / frac use the mean, "
                               "kernel_ms_median / frac_median the median" % args.steps,
                     "kernel_share_of_step": kernel_ms / ms_per_step},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(e2e_dt.item()) * 1e3, "steps": e2e_steps,
                "api": "sad_distill_step_host (pinned host buffers in, losses + normaliser + gradients out)",
                "copy_only": {"ms_per_step": copy_only_ms, "value": world * anchors / (copy_only_ms * 1e-3) / 1e6, "unit": UNIT,
                              "e2e_over_copy_only": float(e2e_dt.item()) * 1e3 / copy_only_ms,
                              "what": "the same %d H2D + %d D2H bytes per step and rank as two concurrent copy streams with NO kernels "
                                      "(max over ranks): the bound this host's PCIe / memory path sets for the e2e step" % (h2d, d2h)},
                "gradients_on_device": {"value": e2e_value2, "unit": UNIT, "ms_per_step": float(e2e_dt2.item()) * 1e3,
                                        "d2h_bytes_per_step": 4 * (len(host) * 2 + 1),
                                        "note": "same call with d_logits = NULL: gradients stay in HBM for ConvGradient, only losses + normaliser return"}},
        "gpu_launches": int(launches) + (head_line["gpu_launches"] if head_line else 0),
        "clocks": sampler.result(),
    }
    if head_f16_line:
        line["head_step_f16"] = head_f16_line
        line["gpu_launches"] += head_f16_line["gpu_launches"]
    if head_x3_line:
        head_x3_line["roofline"]["note"] = ("achieved = 3 x algorithmic conv flops (W_hi X_hi + W_hi X_lo + W_lo X_hi) / whole step time; "
                                            "imgs/s is the algorithmic rate.  " + head_x3_line["roofline"]["note"])
        line["head_step_f32x3"] = head_x3_line
        line["gpu_launches"] += head_x3_line["gpu_launches"]
    if head_line:
        line["head_step"] = head_line
    if full_line:
        line["full_step"] = full_line
        line["gpu_launches"] += full_line["gpu_launches"]
    if full16_line:
        line["full_step_bs16"] = full16_line
        line["gpu_launches"] += full16_line["gpu_launches"]
    if full5_line:
        line["full_step_config5"] = full5_line
        line["gpu_launches"] += full5_line["gpu_launches"]
    if full5_f16_line:
        line["full_step_config5_heads_f16"] = full5_f16_line
        line["gpu_launches"] += full5_f16_line["gpu_launches"]
    step_imgs = {}
    for key, obj in (("head_step", head_line), ("head_step_f16", head_f16_line), ("head_step_f32x3", head_x3_line), ("full_step", full_line), ("full_step_bs16", full16_line),
                     ("full_step_config5", full5_line), ("full_step_config5_heads_f16", full5_f16_line)):
        if obj:
            step_imgs[key] = {"imgs_s": obj["value"], "ms_per_step": obj["ms_per_step"], "allreduce_ms": obj.get("allreduce_ms"),
                              "allreduce_exposed_ms": obj.get("allreduce_exposed_ms"), "multi_gpu_check": obj.get("multi_gpu_check")}
    line["config"]["step_imgs_s"] = step_imgs
    line["config"]["e2e_ms_per_step"] = float(e2e_dt.item()) * 1e3
    line["config"]["e2e_copy_only_ms_per_step"] = copy_only_ms
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_oracle
        sample = host  # the full configs[1] batch; passes are repeated until >= 10 core-seconds of CPU work were timed
        cpu_reference_pass(sample[2:], cpu_oracle)  # warm the pages / thread pool
        cores = cpu_oracle.num_threads()
        passes, t0 = 0, time.perf_counter()
        while passes < 3 or (time.perf_counter() - t0) * cores < 10.0:
            cpu_reference_pass(sample, cpu_oracle)
            passes += 1
        dt = (time.perf_counter() - t0) / passes
        line["cpu_baseline"] = {"value": anchors / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d passes over the full configs[1] batch (245520 anchors each): PowSum + loss fwd + grad, %.2f s per "
                                          "pass, %.1f core-seconds in total" % (passes, dt, dt * passes * cores)}
        # one thread, on a bounded sample (levels P5..P7), as SURVEY.md 8(d) asks
        cpu_oracle.set_num_threads(1)
        small = sample[2:]
        t0 = time.perf_counter()
        cpu_reference_pass(small, cpu_oracle)
        line["cpu_baseline"]["single_thread_value"] = synthetic.anchors_in(small) / (time.perf_counter() - t0) / 1e6
        cpu_oracle.set_num_threads(cores)
        # the reference's own CUDA operators on this GPU, same inputs (checker library, timed beside the product)
        try:
            if os.path.exists(cpu_oracle.REF_GPU_LIB):
                ref_ms = operator_net_ms(cpu_oracle.REF_GPU_LIB, host)
                line["cpu_baseline"]["reference_cuda_on_this_gpu"] = {
                    "value": anchors / (ref_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ref_ms,
                    "what": "oracle/_ref/libref_ops.so: the reference's unmodified pow_sum_op.cu + "
                            "sigmoid_adaptive_distillation_loss_op.cu run as the PowSum + 5 x (loss, gradient) operator net, "
                            "wall time per RunNet"}
            own_ms, own_fused_ms = operator_net_ms(None, host), operator_net_ms(None, host, fuse=True)
            line["operator_net"] = {
                "value": anchors / (own_fused_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": own_fused_ms,
                "unfused_ms_per_step": own_ms,
                "what": "the same NetDef through libcaffe2_detectron_ops_gpu.so (the drop-in operators), wall time per RunNet: "
                        "as built (11 operators) and after the FuseAdaptiveDistillOps graph pass (one operator)"}
        except Exception as e:  # a checker failure must not take the bench line down
            line["cpu_baseline"]["reference_cuda_error"] = str(e)[:200]
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
