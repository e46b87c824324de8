"""Weight files of the reference (pickled `{'blobs': {name: ndarray}, 'cfg': yaml}` dictionaries) for the blobs on the path
(SURVEY.md §8f rank 4): host mirror of detectron/lib/utils/net.py:50-143 (`initialize_gpu_from_weights_file`) and :146-181
(`save_model_to_weights_file`), over this package's `{blob name: tensor}` parameter views (head.RetinaNetHead.params, whose
names are the reference's blob names) and their momentum views.

Kept from the reference, rule by rule:
  * the dictionary may be the blobs themselves or hold them under 'blobs' (net.py:66-69);
  * with distillation on, the teacher's file is merged under the prefix 'teacher/' (net.py:70-77);
  * a parameter `_[xyz]_foo` missing from the file initialises from `foo` (net.py:85-96);
  * a missing source blob is reported and skipped (net.py:97-99); a shape mismatch is reported and skipped — this fork's change
    (net.py:112-115; stock Detectron asserts);
  * values are cast to float32 (net.py:123-125); `<name>_momentum` is loaded when present (net.py:126-130);
  * blobs of the file that the model does not use are preserved and written back on save (net.py:132-143, 171-179);
  * saving writes every parameter, the momentum of the trainable ones, the preserved blobs and the cfg text (net.py:152-181),
    pickled with protocol 2 so the Python-2 reference can read the file back.
Pure host logic: tensors are filled with `copy_`, no kernel of this repository runs here.
"""
import logging
import pickle
from collections import OrderedDict

import numpy as np
import torch

logger = logging.getLogger(__name__)


def unscope_name(name):
    """utils/c2.py:95-102: 'gpu_0/foo' -> 'foo'.  This fork cuts at the FIRST separator (stock Detectron: the last) and passes names
    that do not start with 'gpu' through, so 'gpu_0/teacher/foo' -> 'teacher/foo' and 'teacher/foo' stays."""
    if name[:3] != "gpu":
        return name
    return name[name.find("/") + 1:]


def load_blobs(weights_file):
    """(blobs, cfg_text or None) of a reference weight file.  Python-2 pickles hold byte strings: latin1 keeps ndarrays intact."""
    with open(weights_file, "rb") as f:
        src = pickle.load(f, encoding="latin1")
    cfg_text = None
    if "cfg" in src:
        cfg_text = src["cfg"]
    if "blobs" in src:
        src = src["blobs"]
    return {(k.decode() if isinstance(k, bytes) else k): v for k, v in src.items()}, cfg_text


def merge_teacher_blobs(src_blobs, teacher_blobs):
    """net.py:70-77: the teacher's parameters live under 'teacher/<name>' in the same dictionary."""
    for k, v in teacher_blobs.items():
        src_blobs["teacher/{}".format(k)] = v
    return src_blobs


class LoadReport:
    def __init__(self):
        self.loaded, self.with_momentum, self.not_found, self.shape_mismatch = [], [], [], []
        self.preserved = OrderedDict()

    def __repr__(self):
        return "LoadReport(loaded=%d, momentum=%d, not_found=%d, shape_mismatch=%d, preserved=%d)" % (
            len(self.loaded), len(self.with_momentum), len(self.not_found), len(self.shape_mismatch), len(self.preserved))


def initialize_from_blobs(params, src_blobs, momentum=None):
    """Fill `params` ({possibly scoped blob name: tensor}, model order) from `src_blobs`; `momentum`: {same names: tensor}."""
    report = LoadReport()
    unscoped = OrderedDict((unscope_name(str(n)), n) for n in params)
    for name, key in unscoped.items():
        if name.find("]_") >= 0 and name not in src_blobs:
            src_name = name[name.find("]_") + 2:]
        else:
            src_name = name
        if src_name not in src_blobs:
            logger.info("{:s} not found".format(src_name))
            report.not_found.append(src_name)
            continue
        src = np.asarray(src_blobs[src_name])
        dst = params[key]
        if tuple(dst.shape) != tuple(src.shape):
            logger.info("Shape missmatch: name: {} src: {}, dst: {}".format(name, tuple(dst.shape), tuple(src.shape)))
            report.shape_mismatch.append(name)
            continue
        dst.copy_(torch.from_numpy(np.ascontiguousarray(src.astype(np.float32, copy=False))))
        report.loaded.append(name)
        if src_name + "_momentum" in src_blobs and momentum is not None and key in momentum:
            m = np.asarray(src_blobs[src_name + "_momentum"]).astype(np.float32, copy=False)
            momentum[key].copy_(torch.from_numpy(np.ascontiguousarray(m)).view(momentum[key].shape))
            report.with_momentum.append(name)
    for src_name, v in src_blobs.items():
        if src_name not in unscoped and not src_name.endswith("_momentum") and v is not None:
            report.preserved[src_name] = v
    return report


def initialize_from_weights_file(params, weights_file, momentum=None, teacher_weights_file=None):
    """initialize_gpu_from_weights_file (net.py:50-143) for one replica; every rank loads the same file, which is what
    broadcast_parameters (net.py:184-215) achieves in the reference."""
    logger.info("Loading weights from: {}".format(weights_file))
    src_blobs, _ = load_blobs(weights_file)
    if teacher_weights_file is not None:
        logger.info("Loading teacher weights from: {}".format(teacher_weights_file))
        merge_teacher_blobs(src_blobs, load_blobs(teacher_weights_file)[0])
    return initialize_from_blobs(params, src_blobs, momentum)


def save_to_weights_file(weights_file, params, momentum=None, preserved=None, cfg_text=""):
    """save_model_to_weights_file (net.py:146-181): unscoped names, parameters, then momentum, then preserved blobs."""
    blobs = {}
    for n, t in params.items():
        u = unscope_name(str(n))
        if u not in blobs:
            blobs[u] = t.detach().cpu().numpy().copy()
    for n, t in (momentum or {}).items():
        u = unscope_name(str(n)) + "_momentum"
        if u not in blobs:
            blobs[u] = t.detach().cpu().numpy().copy()
    for n, v in (preserved or {}).items():
        if n not in blobs:
            blobs[n] = v
    with open(weights_file, "wb") as f:
        pickle.dump(dict(blobs=blobs, cfg=cfg_text), f, 2)
    return sorted(blobs)
