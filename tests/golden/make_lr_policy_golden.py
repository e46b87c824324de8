"""Generates tests/golden/lr_policy_golden.json by IMPORTING the reference's own detectron/lib/utils/lr_policy.py in this
container (it cannot travel to the GPU box) and evaluating get_lr_at_iter over several SOLVER settings.  The module only
needs `cfg.SOLVER` from core.config; core.config itself pulls in the whole Python-2 Detectron tree, so a stand-in module
`core.config` exposing a plain `cfg.SOLVER` namespace is installed in sys.modules first — lr_policy.py runs unmodified.
Also records _get_lr_change_ratio (detector.py:673-678, a 5-line pure function re-evaluated here from the reference source text).

    python tests/golden/make_lr_policy_golden.py
"""
import importlib.util
import json
import os
import re
import sys
import types

import numpy as np

REF = "/root/reference/detectron/lib"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "distill_r50_8gpu": dict(BASE_LR=0.01, LR_POLICY="steps_with_decay", GAMMA=0.1, MAX_ITER=270000, STEPS=[0, 180000, 240000],
                             WARM_UP_ITERS=1000, WARM_UP_FACTOR=1.0 / 3.0, WARM_UP_METHOD="linear", STEP_SIZE=30000, LRS=[]),
    "defaults_step": dict(BASE_LR=0.001, LR_POLICY="step", GAMMA=0.1, MAX_ITER=40000, STEPS=[], WARM_UP_ITERS=500,
                          WARM_UP_FACTOR=1.0 / 3.0, WARM_UP_METHOD="linear", STEP_SIZE=30000, LRS=[]),
    "steps_with_lrs_constant_warmup": dict(BASE_LR=0.02, LR_POLICY="steps_with_lrs", GAMMA=0.1, MAX_ITER=90, STEPS=[0, 60, 80],
                                           LRS=[0.02, 0.002, 0.0002], WARM_UP_ITERS=10, WARM_UP_FACTOR=0.25,
                                           WARM_UP_METHOD="constant", STEP_SIZE=30000),
    "no_warmup": dict(BASE_LR=0.02, LR_POLICY="steps_with_decay", GAMMA=0.5, MAX_ITER=100, STEPS=[0, 10, 20, 70], WARM_UP_ITERS=0,
                      WARM_UP_FACTOR=1.0 / 3.0, WARM_UP_METHOD="linear", STEP_SIZE=30000, LRS=[]),
}
ITERS = {
    "distill_r50_8gpu": [0, 1, 2, 499, 500, 999, 1000, 1001, 90000, 179999, 180000, 239999, 240000, 269999, 270000, 300000],
    "defaults_step": [0, 1, 250, 499, 500, 29999, 30000, 59999, 60000, 39999],
    "steps_with_lrs_constant_warmup": [0, 5, 9, 10, 59, 60, 79, 80, 89, 90, 1000],
    "no_warmup": [0, 9, 10, 19, 20, 69, 70, 99, 100, 150],
}


def load_reference_lr_policy(solver):
    core = types.ModuleType("core")
    config = types.ModuleType("core.config")
    config.cfg = types.SimpleNamespace(SOLVER=solver)
    core.config = config
    sys.modules["core"], sys.modules["core.config"] = core, config
    spec = importlib.util.spec_from_file_location("ref_lr_policy", os.path.join(REF, "utils", "lr_policy.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_change_ratio():
    src = open(os.path.join(REF, "modeling", "detector.py")).read()
    m = re.search(r"def _get_lr_change_ratio\(cur_lr, new_lr\):\n(?:    .*\n)+", src)
    ns = {"np": np}
    exec(m.group(0), ns)
    return ns["_get_lr_change_ratio"]


def main():
    out = {"cases": {}, "change_ratio": []}
    for name, kw in CASES.items():
        solver = types.SimpleNamespace(**kw)
        mod = load_reference_lr_policy(solver)
        vals = [mod.get_lr_at_iter(it) for it in ITERS[name]]
        assert all(isinstance(v, np.float32) for v in vals)
        out["cases"][name] = {"solver": kw, "iters": ITERS[name], "lr_f32_hex": [float(v).hex() for v in vals],
                              "step_index": [mod.get_step_index(it) if kw["STEPS"] else None for it in ITERS[name]]}
    ratio = reference_change_ratio()
    for cur, new in [(0.01, 0.001), (0.001, 0.01), (0.0033333334, 0.0033400002), (0.0, 0.01), (0.01, 0.01), (1e-8, 0.02)]:
        out["change_ratio"].append({"cur": cur, "new": new, "ratio_hex": float(ratio(np.float32(cur), np.float32(new))).hex()})
    with open(os.path.join(HERE, "lr_policy_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
