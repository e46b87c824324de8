"""CPU: host-side logic that needs no device — parameter naming / layout of the head mirror, synthetic geometry, bench arithmetic."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_head_parameter_names_follow_the_reference_blobs():
    from sad_b200 import head
    names = head.param_names(4)
    assert len(names) == 20 and len(set(names)) == 20
    # weights first, biases second: the optimiser's two parameter classes are two contiguous segments (optimizer.py:115-124)
    assert all(n.endswith("_w") for n in names[:10]) and all(n.endswith("_b") for n in names[10:])
    # names as detectron/lib/modeling/retinanet_heads.py:101-152,188-245 creates them at level k_min = 3
    for n in ("retnet_cls_conv_n0_fpn3_w", "retnet_cls_conv_n3_fpn3_b", "retnet_bbox_conv_n2_fpn3_w", "retnet_cls_pred_fpn3_w",
              "retnet_bbox_pred_fpn3_b"):
        assert n in names
    assert head.param_names(1, k_min=4)[0] == "retnet_cls_conv_n0_fpn4_w"


def test_synthetic_geometry_matches_survey():
    from sad_b200 import synthetic
    s600, s500 = synthetic.level_shapes(600), synthetic.level_shapes(500)
    assert s600 == [(80, 128), (40, 64), (20, 32), (10, 16), (5, 8)]       # 3 x 640 x 1024 input, strides 8..128
    assert s500 == [(64, 112), (32, 56), (16, 28), (8, 14), (4, 7)]        # 3 x 512 x 896
    assert sum(h * w for h, w in s600) * 9 == 122760                       # anchors per image
    assert sum(h * w for h, w in s500) * 9 == 85932
    lv = synthetic.make_pyramid(1, 1, 600)
    assert synthetic.anchors_in(lv) == 122760 and lv[0][0].shape == (1, 720, 80, 128) and lv[0][2].dtype.name == "int32"


def test_bench_constants_and_peaks():
    sys.path.insert(0, ROOT)
    import bench
    assert abs(bench.BYTES_PER_ELEMENT - 16.05) < 1e-9          # PowSum 4 + loss/grad 12.05 B per logit
    peak, src = bench._peaks()
    assert 1000.0 < peak < 10000.0 and isinstance(src, str)
    tpeak, tsrc = bench._tensor_peak()
    assert 300.0 < tpeak < 1500.0 and "bf16" in tsrc
    p = os.path.join(ROOT, "profiles", "distill_kernel_traffic.json")
    if os.path.exists(p):
        t = json.load(open(p))
        assert "distill" in t["kernel"] and t["dram_bytes_per_launch"] > 1e8


def test_bench_default_line_includes_the_fp16_objects(monkeypatch):
    """The driver runs `python bench.py` without flags: the configs[4] step with both heads in fp16 and head_step_f16 are part of
    the default line; --no-heads-f16 drops them.  (Argument parsing only: no device needed.)"""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    seen = {}

    def fake_parse(self, *a, **k):
        ns = real(self, *a, **k)
        seen["ns"] = ns
        raise SystemExit(0)

    real = argparse.ArgumentParser.parse_args
    monkeypatch.setattr(argparse.ArgumentParser, "parse_args", fake_parse)
    for argv, want in ((["bench.py"], True), (["bench.py", "--no-heads-f16"], False), (["bench.py", "--heads-f16"], True)):
        monkeypatch.setattr(sys, "argv", argv)
        try:
            bench.main()
        except SystemExit:
            pass
        assert seen["ns"].teacher_f16 is want and seen["ns"].gpus == 1 and seen["ns"].impl == "b200"
        assert seen["ns"].exchange == "env"        # the exchange form is the environment's / the library default (ncclAllReduce buckets)
    for form in ("allreduce", "gather"):
        monkeypatch.setattr(sys, "argv", ["bench.py", "--exchange", form])
        try:
            bench.main()
        except SystemExit:
            pass
        assert seen["ns"].exchange == form
