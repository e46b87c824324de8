"""GPU parity of the one-launch momentum-SGD update against the CPU oracle restating
detectron/lib/modeling/optimizer.py:115-130 (Scale 2x for biases, WeightedSum weight decay) + MomentumSGDKernel
(caffe2/caffe2/sgd/momentum_sgd_op_gpu.cu:23-54).  Tolerance: 2 ulp-level (rtol 1e-6): the same handful of fp32
operations, possibly contracted into FMAs on the device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nesterov", [False, True])
@pytest.mark.parametrize("sizes", [[(1000, 1.0, 1e-4), (40, 2.0, 0.0)], [(7, 1.0, 1e-4), (5, 2.0, 0.0), (1026, 1.0, 5e-4), (3, 2.0, 0.0)]])
def test_momentum_sgd_matches_oracle(oracle, sizes, nesterov):
    from sad_b200 import ops
    rng = np.random.default_rng(len(sizes) + int(nesterov))
    n = sum(c for c, _, _ in sizes)
    p, g, m = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
    pd, gd, md = (torch.from_numpy(a.copy()).cuda() for a in (p, g, m))
    lr = torch.tensor(0.0137, device="cuda")
    for step in range(2):     # two steps: the momentum written by the first feeds the second
        ops.momentum_sgd(pd, gd, md, sizes, lr, momentum=0.9, nesterov=nesterov)
        off = 0
        for cnt, mult, wd in sizes:
            sl = slice(off, off + cnt)
            p[sl], g[sl], m[sl] = oracle.momentum_sgd(p[sl], g[sl], m[sl], 0.0137, 0.9, nesterov, mult, wd)
            off += cnt
        torch.cuda.synchronize()
        for got, ref, what in ((pd, p, "param"), (gd, g, "grad"), (md, m, "momentum")):
            np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=2e-6, atol=2e-6, err_msg="%s step %d" % (what, step))  # atol: (1 + mu) * m_new - mu * m cancels (Nesterov)
        g = rng.standard_normal(n).astype(np.float32)
        gd.copy_(torch.from_numpy(g))


def test_momentum_sgd_rejects_bad_segments():
    from sad_b200 import ops
    t = torch.zeros(16, device="cuda")
    with pytest.raises(ValueError):
        ops.momentum_sgd(t, t.clone(), t.clone(), [(8, 1.0, 0.0)], torch.tensor(0.1, device="cuda"))


# ---------------------------------------------------------------------------------------------
# the optimiser OPERATORS under the reference's names (csrc/ops/sgd_ops.cc), op for op against the reference's own
# MomentumSGDUpdate / MomentumSGD (caffe2/caffe2/sgd/momentum_sgd_op.{h,cc}, momentum_sgd_op_gpu.cu compiled unmodified into
# oracle/_ref/libref_ops.so) and, for WeightedSum (utility_ops.h:333-378; its .cu is not in _ref), against the oracle restatement
# ---------------------------------------------------------------------------------------------
def _libs():
    import os
    from oracle import cpu_oracle
    from sad_b200 import c2
    ours = c2.OperatorLibrary()
    ref = c2.OperatorLibrary(cpu_oracle.REF_GPU_LIB) if os.path.exists(cpu_oracle.REF_GPU_LIB) else None
    return c2, ours, ref


@pytest.mark.parametrize("nesterov", [0, 1])
@pytest.mark.parametrize("n", [1, 1000, 589824 + 3])
def test_momentum_sgd_update_operator_is_bit_identical_to_the_reference_operator(n, nesterov):
    c2, ours, ref = _libs()
    if ref is None:
        pytest.skip("oracle/_ref/libref_ops.so not present")
    assert ours.SchemaArity("MomentumSGDUpdate") == (4, 4, 3, 3) == ref.SchemaArity("MomentumSGDUpdate")   # momentum_sgd_op.cc:56-58
    rng = np.random.default_rng(n + nesterov)
    g, m, p = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
    dev = c2.DeviceOption(c2.CUDA, 0)
    out = {}
    for name, lib in (("ours", ours), ("ref", ref)):
        ws = lib.Workspace()
        for blob, a in (("g", g), ("m", m), ("p", p)):
            ws.FeedBlob(blob, torch.from_numpy(a.copy()).cuda())
        ws.FeedBlob("lr", torch.tensor([0.0173], device="cuda"))
        # in place, as optimizer.py:125-130 emits it
        op = c2.CreateOperator("MomentumSGDUpdate", ["g", "m", "lr", "p"], ["g", "m", "p"], momentum=0.9, nesterov=nesterov, device_option=dev)
        ws.RunOperatorOnce(op)
        ws.RunOperatorOnce(op)       # a second step on the updated blobs
        out[name] = [ws.FetchBlob(b) for b in ("g", "m", "p")]
    for a, b, what in zip(out["ours"], out["ref"], ("grad", "momentum", "param")):
        assert a.tobytes() == b.tobytes(), what


def test_momentum_sgd_operator_without_parameter_and_out_of_place():
    c2, ours, ref = _libs()
    rng = np.random.default_rng(5)
    n = 4099
    g, m, p = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
    dev = c2.DeviceOption(c2.CUDA, 0)
    ws = ours.Workspace()
    for blob, a in (("g", g), ("m", m), ("p", p)):
        ws.FeedBlob(blob, torch.from_numpy(a.copy()).cuda())
    ws.FeedBlob("lr", torch.tensor([0.05], device="cuda"))
    ws.RunOperatorOnce(c2.CreateOperator("MomentumSGD", ["g", "m", "lr"], ["g2", "m2"], momentum=0.9, device_option=dev))
    adj = np.float32(0.05) * g + np.float32(0.9) * m
    np.testing.assert_allclose(ws.FetchBlob("g2"), adj, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(ws.FetchBlob("m2"), adj, rtol=2e-6, atol=1e-7)
    # out of place: inputs untouched, param_out = param - adjusted
    ws.RunOperatorOnce(c2.CreateOperator("MomentumSGDUpdate", ["g", "m", "lr", "p"], ["g3", "m3", "p3"], momentum=0.9, device_option=dev))
    assert ws.FetchBlob("p").tobytes() == p.tobytes() and ws.FetchBlob("g").tobytes() == g.tobytes()
    np.testing.assert_allclose(ws.FetchBlob("p3"), p - adj, rtol=2e-6, atol=1e-6)
    if ref is not None:   # the gradient of an optimiser operator is an error in both libraries (SHOULD_NOT_DO_GRADIENT)
        for lib in (ours, ref):
            with pytest.raises(c2.EnforceNotMet):
                lib.GetGradientDefs(c2.CreateOperator("MomentumSGDUpdate", ["g", "m", "lr", "p"], ["g", "m", "p"], device_option=dev), ["g_grad", "", ""])


def test_weighted_sum_operator_matches_oracle(oracle):
    c2, ours, _ = _libs()
    rng = np.random.default_rng(9)
    n = 100003
    grad, param = (rng.standard_normal(n).astype(np.float32) for _ in range(2))
    dev = c2.DeviceOption(c2.CUDA, 0)
    ws = ours.Workspace()
    ws.FeedBlob("grad", torch.from_numpy(grad.copy()).cuda())
    ws.FeedBlob("param", torch.from_numpy(param.copy()).cuda())
    ws.FeedBlob("one", torch.tensor([1.0], device="cuda"))
    ws.FeedBlob("wd", torch.tensor([1e-4], device="cuda"))
    # optimizer.py:122-124: WeightedSum([param_grad, one, param, wd], param_grad), in place with input 0
    ws.RunOperatorOnce(c2.CreateOperator("WeightedSum", ["grad", "one", "param", "wd"], ["grad"], device_option=dev))
    got = ws.FetchBlob("grad")
    ref = oracle.weighted_sum([grad, param], [1.0, 1e-4])
    assert got.tobytes() == ref.tobytes()
    # arity: an odd number of inputs is refused by the schema; in place with an input other than 0 returns false
    with pytest.raises(c2.EnforceNotMet):
        ws.RunOperatorOnce(c2.CreateOperator("WeightedSum", ["grad", "one", "param"], ["out"], device_option=dev))
    with pytest.raises(c2.EnforceNotMet):
        ws.RunOperatorOnce(c2.CreateOperator("WeightedSum", ["grad", "one", "param", "wd"], ["param"], device_option=dev))


# ---------------------------------------------------------------------------------------------
# overflow guard of mixed-precision training: non-finite flag + guarded optimiser launch + host-side loss scaler
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bad", [None, float("inf"), float("-inf"), float("nan")])
@pytest.mark.parametrize("n", [3, 4096, 100003])
def test_nonfinite_flag_and_guarded_sgd(n, bad):
    from sad_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n)
    p, gr, m = (torch.randn(n, device="cuda", generator=g) for _ in range(3))
    if bad is not None:
        gr[n - 2] = bad                 # in the scalar tail for n % 4 != 0, in a float4 otherwise
    flag = torch.zeros((), dtype=torch.int32, device="cuda")
    ops.nonfinite_flag(gr, flag)
    assert int(flag.item()) == (0 if bad is None else 1)
    before = (p.clone(), gr.clone(), m.clone())
    lr = torch.tensor(0.01, device="cuda")
    ops.momentum_sgd(p, gr, m, [(n, 1.0, 1e-4)], lr, skip_flag=flag)
    torch.cuda.synchronize()
    if bad is None:   # the guarded launch with a clear flag is the plain launch
        p2, g2, m2 = (t.clone() for t in before)
        ops.momentum_sgd(p2, g2, m2, [(n, 1.0, 1e-4)], lr)
        assert torch.equal(p, p2) and torch.equal(gr, g2) and torch.equal(m, m2)
    else:             # skipped on the device: nothing moved, nothing was poisoned
        for a, b in zip((p, gr, m), before):
            assert a.tobytes() == b.tobytes() if hasattr(a, "tobytes") else torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
        assert bool(torch.isfinite(p).all()) and bool(torch.isfinite(m).all())


def test_loss_scaler_halves_on_overflow_and_grows_after_clean_steps():
    from sad_b200 import solver
    from sad_b200.head import RetinaNetHead
    head = RetinaNetHead(1, [(8, 32)], dim=32, num_convs=1, num_anchors=3, num_classes=4, compute_f16=True)
    flag = torch.zeros((), dtype=torch.int32, device="cuda")
    sc = solver.LossScaler(head, flag, init_scale=4096.0, growth_interval=4, check_every=2)
    assert head.f16_grad_scale() == 4096.0
    assert sc.update() is False                 # step 1: not a check step
    flag.fill_(1)
    assert sc.update() is True                  # step 2: overflow seen -> halved, flag cleared
    assert head.f16_grad_scale() == 2048.0 and int(flag.item()) == 0
    changed = [sc.update() for _ in range(4)]   # 4 clean steps = 2 checks -> grows back
    assert changed == [False, False, False, True] and head.f16_grad_scale() == 4096.0
    head.close()


def test_fp16_head_overflow_is_caught_and_the_step_skipped():
    """A loss scale far too large for the incoming gradient overflows the fp16 gradient tensors: the head's parameter gradients come
    out non-finite, the flag goes up, the guarded optimiser launch leaves the parameters alone, and the scaler's next look halves
    the scale until a step goes through."""
    from sad_b200 import ops, solver
    from sad_b200.head import RetinaNetHead
    shapes = [(8, 32), (4, 16)]
    head = RetinaNetHead(1, shapes, dim=64, num_convs=1, num_anchors=3, num_classes=8, compute_f16=True, seed=3)
    g = torch.Generator(device="cuda").manual_seed(1)
    fpn = [torch.randn(1, 64, h, w, device="cuda", generator=g) for h, w in shapes]
    d_cls = [torch.randn(1, head.cls_out, h, w, device="cuda", generator=g) * 30.0 for h, w in shapes]   # 30 * 4096 > 65504
    d_box = [torch.randn(1, head.bbox_out, h, w, device="cuda", generator=g) for h, w in shapes]
    flag = torch.zeros((), dtype=torch.int32, device="cuda")
    sc = solver.LossScaler(head, flag, init_scale=4096.0, check_every=1)
    mom = torch.zeros_like(head.flat_params)
    lr = torch.tensor(0.01, device="cuda")
    before = head.flat_params.clone()
    skipped = 0
    for step in range(12):
        head.forward(fpn)
        head.backward(d_cls, d_box)
        ops.nonfinite_flag(head.flat_grads, flag)
        ops.momentum_sgd(head.flat_params, head.flat_grads, mom, head.sgd_segments(1e-4), lr, skip_flag=flag)
        if int(flag.item()):
            skipped += 1
            assert torch.equal(head.flat_params, before), "an overflowed step must not touch the parameters"
        else:
            break
        sc.update()
    assert skipped >= 1 and head.f16_grad_scale() < 4096.0
    assert not torch.equal(head.flat_params, before) and bool(torch.isfinite(head.flat_params).all())
    head.close()
