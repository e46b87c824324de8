"""Learning-rate policy and the workspace-side learning-rate update of the training step.

Host mirror of detectron/lib/utils/lr_policy.py:28-114 (`get_lr_at_iter`, the three `lr_func_*` policies, `get_step_index`,
warm-up) and of detectron/lib/modeling/detector.py:598-648,673-678 (`UpdateWorkspaceLr`, `_SetNewLr`, `_CorrectMomentum`,
`_get_lr_change_ratio`), over this package's state: the learning rate is ONE device fp32 scalar (the blob `gpu_k/lr` the
momentum-SGD launch reads) and the update history is ONE flat buffer, so the momentum correction is one `Scale` launch
(`sad_scale_f32`) instead of one operator per `<param>_momentum` blob.  Same names, argument meaning and error behaviour
as the reference functions; the SOLVER defaults are detectron/lib/core/config.py:580-642.  Pure host logic except
`LearningRate.update`, which needs the CUDA library (no CPU path).
"""
import logging

import numpy as np

logger = logging.getLogger(__name__)


class SolverConfig:
    """cfg.SOLVER (detectron/lib/core/config.py:580-642), same field names and defaults."""
    BASE_LR = 0.001
    LR_POLICY = "step"
    GAMMA = 0.1
    STEP_SIZE = 30000
    STEPS = []
    LRS = []
    MAX_ITER = 40000
    MOMENTUM = 0.9
    WEIGHT_DECAY = 0.0005
    WARM_UP_ITERS = 500
    WARM_UP_FACTOR = 1.0 / 3.0
    WARM_UP_METHOD = "linear"
    SCALE_MOMENTUM = True
    SCALE_MOMENTUM_THRESHOLD = 1.1
    LOG_LR_CHANGE_THRESHOLD = 1.1

    def __init__(self, **kw):
        for k, v in kw.items():
            if not hasattr(type(self), k):
                raise KeyError("Non-existent config key: SOLVER.%s" % k)
            setattr(self, k, list(v) if isinstance(v, (list, tuple)) else v)

    @classmethod
    def retinanet_r50_distillation(cls, num_gpus=8):
        """SOLVER block of configs/focal_distillation/retinanet_R-50-FPN_distillation.yaml:6-13 (NUM_GPUS 8, 2 images per GPU;
        the 3x-long schedule covers the labelled + unlabelled alternation); for other GPU counts the linear scaling rule of
        detectron/GETTING_STARTED.md (lr x k, iterations / k)."""
        k = num_gpus / 8.0
        return cls(BASE_LR=0.01 * k, LR_POLICY="steps_with_decay", GAMMA=0.1, WEIGHT_DECAY=0.0001, WARM_UP_ITERS=1000,
                   MAX_ITER=int(round(270000 / k)), STEPS=[0, int(round(180000 / k)), int(round(240000 / k))])


def get_step_index(solver, cur_iter):
    """lr_policy.py:94-102: which learning-rate step `cur_iter` falls into."""
    assert solver.STEPS[0] == 0, "The first step should always start at 0."
    steps = list(solver.STEPS) + [solver.MAX_ITER]
    ind = 0
    for ind, step in enumerate(steps):
        if cur_iter < step:
            break
    return ind - 1


def lr_func_steps_with_lrs(solver, cur_iter):
    """lr_policy.py:50-64."""
    return solver.LRS[get_step_index(solver, cur_iter)]


def lr_func_steps_with_decay(solver, cur_iter):
    """lr_policy.py:67-82: base_lr * gamma ** step index."""
    return solver.BASE_LR * solver.GAMMA ** get_step_index(solver, cur_iter)


def lr_func_step(solver, cur_iter):
    """lr_policy.py:85-90."""
    return solver.BASE_LR * solver.GAMMA ** (cur_iter // solver.STEP_SIZE)


_POLICIES = {"steps_with_lrs": lr_func_steps_with_lrs, "steps_with_decay": lr_func_steps_with_decay, "step": lr_func_step}


def get_lr_func(solver):
    """lr_policy.py:105-111."""
    if solver.LR_POLICY not in _POLICIES:
        raise NotImplementedError("Unknown LR policy: {}".format(solver.LR_POLICY))
    return _POLICIES[solver.LR_POLICY]


def get_lr_at_iter(solver, it):
    """lr_policy.py:28-43: policy value, times the warm-up factor while it < WARM_UP_ITERS; returned as np.float32."""
    lr = get_lr_func(solver)(solver, it)
    if it < solver.WARM_UP_ITERS:
        method = solver.WARM_UP_METHOD
        if method == "constant":
            warmup_factor = solver.WARM_UP_FACTOR
        elif method == "linear":
            alpha = it / solver.WARM_UP_ITERS  # true division (the module imports division from __future__)
            warmup_factor = solver.WARM_UP_FACTOR * (1 - alpha) + alpha
        else:
            raise KeyError("Unknown SOLVER.WARM_UP_METHOD: {}".format(method))
        lr *= warmup_factor
    return np.float32(lr)


def get_lr_change_ratio(cur_lr, new_lr):
    """detector.py:673-678."""
    eps = 1e-10
    return np.max((new_lr / np.max((cur_lr, eps)), cur_lr / np.max((new_lr, eps))))


def momentum_correction(solver, cur_lr, new_lr):
    """The factor _SetNewLr hands to _CorrectMomentum (detector.py:616-626), or None when no correction applies."""
    ratio = get_lr_change_ratio(cur_lr, new_lr)
    if solver.SCALE_MOMENTUM and cur_lr > 1e-7 and ratio > solver.SCALE_MOMENTUM_THRESHOLD:
        return new_lr / cur_lr
    return None


class LearningRate:
    """The `lr` blob and its update (detector.py:598-648).  `lr_blob`: CUDA fp32 scalar tensor read by the optimiser launch;
    `momentum_buffers`: the flat update-history buffers to rescale when the rate jumps (student only — the teacher has none)."""

    def __init__(self, solver, lr_blob, momentum_buffers=()):
        self.solver, self.lr_blob, self.momentum_buffers = solver, lr_blob, list(momentum_buffers)
        self.corrections = 0
        # host copy of the blob's value: the blob is only ever written here, so reading it back every iteration (a full device
        # synchronisation per training step) is not needed.  One read at construction picks up a value loaded from a checkpoint.
        self._cur_lr = np.float32(lr_blob.item())

    def update(self, cur_iter, new_lr=None):
        """UpdateWorkspaceLr(cur_iter, new_lr): the workspace is the one source of truth for the current rate."""
        from . import ops
        if new_lr is None:
            new_lr = get_lr_at_iter(self.solver, cur_iter)
        new_lr = np.float32(new_lr)
        cur_lr = self._cur_lr
        if cur_lr != new_lr:
            self._cur_lr = new_lr
            ratio = get_lr_change_ratio(cur_lr, new_lr)
            if ratio > self.solver.LOG_LR_CHANGE_THRESHOLD:
                logger.info("Changing learning rate {:.6f} -> {:.6f} at iter {:d}".format(cur_lr, new_lr, cur_iter))
            self.lr_blob.fill_(float(new_lr))
            correction = momentum_correction(self.solver, cur_lr, new_lr)
            if correction is not None:
                logger.info("Scaling update history by {:.6f} (new lr / old lr)".format(correction))
                for buf in self.momentum_buffers:
                    ops.scale_(buf, correction)
                self.corrections += 1
        return new_lr


class LossScaler:
    """Dynamic loss scaling for the mixed-fp16 head (BASELINE.json configs[4]) without a host round trip per step.

    On the device, every step: ops.nonfinite_flag over the REDUCED head gradient sets `flag`, and the guarded optimiser launch
    (ops.momentum_sgd(..., skip_flag=flag)) leaves parameters and update history untouched when it is set — an overflowed step
    is skipped by all ranks alike.  On the host, every `check_every` steps: `update()` reads the flag (one 4-byte copy); if it is
    set the scale is halved and the flag cleared, and after `growth_interval` clean steps it is doubled (the usual policy).
    Returns True when the scale changed: the head takes the scale by value, so a captured step graph has to be captured again."""

    def __init__(self, head, flag, init_scale=4096.0, growth_interval=2000, check_every=50, min_scale=1.0, max_scale=65536.0):
        self.head, self.flag = head, flag
        self.scale, self.growth_interval, self.check_every = float(init_scale), int(growth_interval), int(check_every)
        self.min_scale, self.max_scale = float(min_scale), float(max_scale)
        self.clean_steps, self.steps, self.skipped_windows = 0, 0, 0
        head.set_f16_grad_scale(self.scale)

    def update(self, force=False):
        self.steps += 1
        if not force and self.steps % self.check_every:
            return False
        overflowed = bool(int(self.flag.item()))
        if overflowed:
            self.flag.zero_()
            self.skipped_windows += 1
            self.clean_steps = 0
            new = max(self.min_scale, self.scale * 0.5)
        else:
            self.clean_steps += self.check_every
            new = self.scale
            if self.clean_steps >= self.growth_interval:
                self.clean_steps = 0
                new = min(self.max_scale, self.scale * 2.0)
        if new != self.scale:
            self.scale = new
            self.head.set_f16_grad_scale(new)
            return True
        return False
