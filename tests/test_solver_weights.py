"""CPU: learning-rate policy, momentum-correction rule and weight-file IO against golden vectors produced
by the reference's OWN detectron/lib/utils/lr_policy.py (tests/golden/make_lr_policy_golden.py imports it unmodified) and against
the rules of detectron/lib/utils/net.py:50-181 stated one by one."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import ROOT
from sad_b200 import solver as S
from sad_b200 import weights_io as W

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "lr_policy_golden.json")))


@pytest.mark.parametrize("case", sorted(GOLDEN["cases"]))
def test_lr_policy_bit_exact_against_reference_module(case):
    g = GOLDEN["cases"][case]
    cfg = S.SolverConfig(**g["solver"])
    for it, want_hex, want_ind in zip(g["iters"], g["lr_f32_hex"], g["step_index"]):
        got = S.get_lr_at_iter(cfg, it)
        assert isinstance(got, np.float32)
        assert float(got).hex() == want_hex, (case, it)
        if want_ind is not None:
            assert S.get_step_index(cfg, it) == want_ind


def test_lr_change_ratio_matches_reference():
    for row in GOLDEN["change_ratio"]:
        got = S.get_lr_change_ratio(np.float32(row["cur"]), np.float32(row["new"]))
        assert float(got).hex() == row["ratio_hex"], row


def test_policy_errors_like_the_reference():
    with pytest.raises(NotImplementedError, match="Unknown LR policy"):
        S.get_lr_at_iter(S.SolverConfig(LR_POLICY="cosine"), 0)
    with pytest.raises(KeyError, match="WARM_UP_METHOD"):
        S.get_lr_at_iter(S.SolverConfig(WARM_UP_METHOD="exp"), 0)
    with pytest.raises(AssertionError, match="first step"):
        S.get_lr_at_iter(S.SolverConfig(LR_POLICY="steps_with_decay", STEPS=[10, 20]), 0)
    with pytest.raises(KeyError, match="Non-existent config key"):
        S.SolverConfig(LEARNING_RATE=1.0)


def test_momentum_correction_rule():
    cfg = S.SolverConfig()
    # detector.py:623-626: only above the 1.1 threshold and for a current rate above 1e-7
    assert S.momentum_correction(cfg, np.float32(0.01), np.float32(0.001)) == pytest.approx(0.1)
    assert S.momentum_correction(cfg, np.float32(0.01), np.float32(0.0105)) is None       # warm-up sized change
    assert S.momentum_correction(cfg, np.float32(0.0), np.float32(0.01)) is None          # first iteration: lr blob starts at 0
    assert S.momentum_correction(S.SolverConfig(SCALE_MOMENTUM=False), np.float32(0.01), np.float32(0.001)) is None
    # the distillation schedule: one correction per decay step, none during warm-up (ratios 1.002 .. 1.0007 per iteration)
    cfg = S.SolverConfig.retinanet_r50_distillation(8)
    assert (cfg.MAX_ITER, cfg.STEPS, cfg.WARM_UP_ITERS) == (270000, [0, 180000, 240000], 1000)
    n, cur = 0, S.get_lr_at_iter(cfg, 0)
    for it in list(range(1, 1200)) + [179999, 180000, 180001, 239999, 240000, 240001]:
        new = S.get_lr_at_iter(cfg, it)
        if new != cur and S.momentum_correction(cfg, cur, new) is not None:
            n += 1
        cur = new
    assert n == 2
    half = S.SolverConfig.retinanet_r50_distillation(4)
    assert half.BASE_LR == pytest.approx(0.005) and half.MAX_ITER == 540000


def _params():
    g = torch.Generator().manual_seed(0)
    shapes = {"gpu_0/retnet_cls_conv_n0_fpn3_w": (4, 4, 3, 3), "gpu_0/retnet_cls_conv_n0_fpn3_b": (4,),
              "gpu_0/_[mask]_fcn1_w": (2, 2), "gpu_0/retnet_bbox_pred_fpn3_w": (36, 4, 3, 3), "teacher/conv1_w": (3, 3)}
    return {k: torch.randn(s, generator=g) for k, s in shapes.items()}


def test_unscope_name():
    assert W.unscope_name("gpu_0/conv1_w") == "conv1_w"
    assert W.unscope_name("gpu_3/teacher/conv1_w") == "teacher/conv1_w"   # utils/c2.py:101-102: cut at the FIRST separator
    assert W.unscope_name("teacher/conv1_w") == "teacher/conv1_w"
    assert W.unscope_name("conv1_w") == "conv1_w"


def test_weight_file_rules(tmp_path):
    rng = np.random.default_rng(0)
    blobs = {
        "retnet_cls_conv_n0_fpn3_w": rng.normal(size=(4, 4, 3, 3)),                       # float64 in the file: cast to float32
        "retnet_cls_conv_n0_fpn3_w_momentum": rng.normal(size=(4, 4, 3, 3)).astype(np.float32),
        "fcn1_w": rng.normal(size=(2, 2)).astype(np.float32),                             # source of `_[mask]_fcn1_w`
        "retnet_bbox_pred_fpn3_w": rng.normal(size=(36, 8, 3, 3)).astype(np.float32),     # wrong shape: skipped, not fatal
        "unused_blob": rng.normal(size=(5,)).astype(np.float32),                          # preserved
        "unused_blob_momentum": rng.normal(size=(5,)).astype(np.float32),                 # momentum of unused: dropped
    }
    student = tmp_path / "student.pkl"
    with open(student, "wb") as f:
        pickle.dump({"blobs": blobs, "cfg": "NUM_GPUS: 8\n"}, f, 2)
    teacher = tmp_path / "teacher.pkl"
    tw = rng.normal(size=(3, 3)).astype(np.float32)
    with open(teacher, "wb") as f:
        pickle.dump({"conv1_w": tw}, f, 2)                                                # old layout: the dictionary IS the blobs
    params = _params()
    before_b = params["gpu_0/retnet_cls_conv_n0_fpn3_b"].clone()
    before_bbox = params["gpu_0/retnet_bbox_pred_fpn3_w"].clone()
    mom = {k: torch.zeros_like(v) for k, v in params.items()}
    rep = W.initialize_from_weights_file(params, str(student), momentum=mom, teacher_weights_file=str(teacher))
    assert rep.loaded == ["retnet_cls_conv_n0_fpn3_w", "_[mask]_fcn1_w", "teacher/conv1_w"]
    assert rep.with_momentum == ["retnet_cls_conv_n0_fpn3_w"]
    assert rep.not_found == ["retnet_cls_conv_n0_fpn3_b"] and rep.shape_mismatch == ["retnet_bbox_pred_fpn3_w"]
    # unused blobs are preserved; so is the source of a `_[xyz]_` parameter (net.py:136-143 compares against the model's names only)
    assert list(rep.preserved) == ["fcn1_w", "unused_blob"]
    assert params["gpu_0/retnet_cls_conv_n0_fpn3_w"].dtype == torch.float32
    assert np.array_equal(params["gpu_0/retnet_cls_conv_n0_fpn3_w"].numpy(), blobs["retnet_cls_conv_n0_fpn3_w"].astype(np.float32))
    assert np.array_equal(mom["gpu_0/retnet_cls_conv_n0_fpn3_w"].numpy(), blobs["retnet_cls_conv_n0_fpn3_w_momentum"])
    assert np.array_equal(params["gpu_0/_[mask]_fcn1_w"].numpy(), blobs["fcn1_w"])
    assert np.array_equal(params["teacher/conv1_w"].numpy(), tw)
    assert torch.equal(params["gpu_0/retnet_cls_conv_n0_fpn3_b"], before_b) and torch.equal(params["gpu_0/retnet_bbox_pred_fpn3_w"], before_bbox)

    # save: unscoped names, momentum of the given blobs, preserved blobs, cfg text; readable by a Python-2 unpickler (protocol 2)
    out = tmp_path / "model_iter9.pkl"
    trainable_mom = {k: v for k, v in mom.items() if not k.startswith("teacher/")}
    names = W.save_to_weights_file(str(out), params, trainable_mom, rep.preserved, cfg_text="NUM_GPUS: 8\n")
    raw = open(out, "rb").read()
    assert raw[:2] == b"\x80\x02"
    back, cfg_text = W.load_blobs(str(out))
    assert cfg_text == "NUM_GPUS: 8\n" and sorted(back) == names
    assert "retnet_cls_conv_n0_fpn3_w_momentum" in back and "unused_blob" in back and "teacher/conv1_w" in back
    assert "teacher/conv1_w_momentum" not in back and "unused_blob_momentum" not in back
    again = _params()
    rep2 = W.initialize_from_weights_file(again, str(out))
    assert rep2.not_found == [] and rep2.shape_mismatch == []
    for k in params:
        assert torch.equal(again[k], params[k]), k


def test_head_parameter_names_round_trip_through_a_weight_file(tmp_path):
    """The head's parameter views carry the reference's blob names (retinanet_heads.py:101-152): a file written from them loads
    back by name.  Built without a device: the name / shape tables only."""
    from sad_b200 import head
    names = head.param_names(4)
    assert names[0] == "retnet_cls_conv_n0_fpn3_w" and "retnet_bbox_pred_fpn3_b" in names and len(names) == 20
    params = {n: torch.full((2, 2), float(i)) for i, n in enumerate(names)}
    W.save_to_weights_file(str(tmp_path / "h.pkl"), params)
    fresh = {n: torch.zeros(2, 2) for n in names}
    rep = W.initialize_from_weights_file(fresh, str(tmp_path / "h.pkl"))
    assert rep.loaded == names and all(torch.equal(fresh[n], params[n]) for n in names)
