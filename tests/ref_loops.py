"""Harness that runs the reference's OWN training / test loops (models/pangu_sample.py train() :96-389, test() :391-575)
unchanged, with everything around them stubbed:

  * `era5_data.config.cfg`        -- the handful of fields the loops read (EPOCHS, ACCUMULATION_STEPS, intervals);
  * `era5_data.utils_data`        -- `loadAllConstants` returns the constants handed to `install()`, `normData` /
                                     `normBackData` are restated from era5_data/utils_data.py:531-546 (three lines each);
  * `era5_data.utils`             -- `mkdirs`, `save_errorScores` (records its arguments), no plotting;
  * `era5_data.score`             -- the PRODUCT's mirror `pangu_b200.score` on CUDA tensors, or a plain-torch
                                     restatement (oracle functions) for CPU runs.

`pangu_sample.py` itself is loaded from a file path: /root/reference in the build container (golden generation),
the git-ignored baseline/_ref copy elsewhere (it travels with the gpurun snapshot; tests skip when it is absent).
Test infrastructure only.
"""
import importlib.util
import logging
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ("/root/reference/models/pangu_sample.py", os.path.join(ROOT, "baseline", "_ref", "models", "pangu_sample.py"))


def find_pangu_sample():
    for p in CANDIDATES:
        if os.path.exists(p):
            return p
    return None


class Recorder:
    def __init__(self):
        self.saved, self.logs = [], []


def install(constants, score_module, epochs=1, accumulation_steps=1):
    """Register the stub modules; returns a Recorder that collects save_errorScores calls."""
    rec = Recorder()
    ns = types.SimpleNamespace
    cfg = ns(PG=ns(TRAIN=ns(EPOCHS=epochs, ACCUMULATION_STEPS=accumulation_steps, SAVE_INTERVAL=10 ** 6, EARLY_STOP=20),
                   VAL=ns(INTERVAL=10 ** 6)), PG_INPUT_PATH="/nonexistent")
    era5 = types.ModuleType("era5_data")
    era5.__path__ = []
    config = types.ModuleType("era5_data.config")
    config.cfg = cfg
    utils_data = types.ModuleType("era5_data.utils_data")

    def normData(upper, surface, statistics):                 # era5_data/utils_data.py:531-537
        sm, ss, um, us = statistics
        return (upper - um) / us, (surface - sm) / ss

    def normBackData(upper, surface, statistics):             # era5_data/utils_data.py:540-546
        sm, ss, um, us = statistics
        return upper * us + um, surface * ss + sm

    utils_data.normData, utils_data.normBackData = normData, normBackData
    utils_data.loadAllConstants = lambda device: {k: (v.to(device) if torch.is_tensor(v) else
                                                      tuple(t.to(device) for t in v) if isinstance(v, tuple) else v)
                                                  for k, v in constants.items()}
    utils = types.ModuleType("era5_data.utils")
    utils.mkdirs = lambda p: os.makedirs(p, exist_ok=True)
    utils.save_errorScores = lambda *a: rec.saved.append(a)
    era5.config, era5.utils_data, era5.utils, era5.score = config, utils_data, utils, score_module
    sys.modules.update({"era5_data": era5, "era5_data.config": config, "era5_data.utils_data": utils_data,
                        "era5_data.utils": utils, "era5_data.score": score_module})
    return rec


def load_pangu_sample(path=None):
    path = path or find_pangu_sample()
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location("reference_pangu_sample", path)
    mod = importlib.util.module_from_spec(spec)
    keep = list(sys.path)
    try:
        spec.loader.exec_module(mod)            # its sys.path.append(parent_dir) is undone below
    finally:
        sys.path[:] = keep
    return mod


class ListLogger:
    """logger.info sink that keeps the formatted lines (train() reports the epoch loss through it)."""

    def __init__(self):
        self.lines = []

    def info(self, msg, *a):
        self.lines.append(msg % a if a else str(msg))


def variable_weights(device="cpu"):
    """era5_data/utils_data.py:505-512 with the weights of era5_data/config.py:52-55."""
    uw = torch.tensor([3.00, 0.60, 1.50, 0.77, 0.54]).reshape(1, 5, 1, 1, 1)
    sw = torch.tensor([1.50, 0.77, 0.66, 3.00]).reshape(1, 4, 1, 1)
    return uw.to(device), sw.to(device), torch.tensor(1.0).to(device), torch.tensor(0.25).to(device)


def statistics_last(stats):
    """era5_data/utils_data.py:395-421 applied to tensors shaped like the .npy files."""
    sm, ss, um, us = stats
    f = lambda t: t.flip(0).permute(1, 3, 0, 2).unsqueeze(-1).contiguous()
    return sm.view(1, 4, 1, 1), ss.view(1, 4, 1, 1), f(um), f(us)


logging.getLogger(__name__)
