#!/usr/bin/env python
"""CPU arm: the UNMODIFIED reference PanguModel (models/pangu_model.py:61-104 over models/layers.py) timed on
the host cores.

    python baseline/run_reference.py --steps K --warmup W        -> one JSON object on stdout

The two reference files live in the git-ignored `baseline/_ref/models/` (copied there, byte for byte, by
`__graft_entry__.build()` in the build container where /root/reference exists; git-ignored files still travel
to the GPU box with the gpurun snapshot).  They are imported as they are; the two imports that cannot be
satisfied offline are stubbed exactly as SURVEY Appendix C describes (`timm.models.layers.{DropPath,
trunc_normal_}`, an empty `era5_data.utils_data`).  Weights: `torch.manual_seed(0)` + the reference's own
`_init_weights`; inputs: the seeded ERA5-shaped tensors of SURVEY 8(d) — generated here, not imported from
the oracle, so this file depends on nothing but torch and the reference.

If `baseline/_ref` is missing the caller (bench.py) falls back to the oracle port and says so.
This script runs in its own process: the reference package is called `models`, like the product's.
"""
import argparse
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return all(os.path.exists(os.path.join(REF, "models", f)) for f in ("layers.py", "pangu_model.py"))


def import_reference():
    import torch
    from torch import nn

    class DropPath(nn.Module):                                  # timm.models.layers.DropPath semantics
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                m.div_(keep)
            return x * m

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return torch.nn.init.trunc_normal_(t, mean, std, a, b)

    timm, timm_models = types.ModuleType("timm"), types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")
    timm_layers.DropPath, timm_layers.trunc_normal_ = DropPath, trunc_normal_
    timm.models, timm_models.layers = timm_models, timm_layers
    era5, era5_utils = types.ModuleType("era5_data"), types.ModuleType("era5_data.utils_data")
    era5.__path__, era5.utils_data = [], era5_utils
    sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers,
                        "era5_data": era5, "era5_data.utils_data": era5_utils})
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    import models.pangu_model as M
    assert os.path.realpath(M.__file__).startswith(os.path.realpath(REF)), M.__file__
    return M


def synth_inputs(torch, seed=1):
    """SURVEY 8(d) synthetic inputs (the same draws as oracle.synth_inputs; order matters)."""
    g = torch.Generator().manual_seed(seed)
    inp = torch.randn(1, 5, 13, 721, 1440, generator=g)
    inp_s = torch.randn(1, 4, 721, 1440, generator=g)
    s_mean, s_std = torch.randn(4, generator=g), torch.rand(4, generator=g) + 0.5
    u_mean, u_std = torch.randn(13, 1, 1, 5, generator=g), torch.rand(13, 1, 1, 5, generator=g) + 0.5
    maps = torch.randn(1, 3, 724, 1440, generator=g)
    const_h = torch.randn(1, 1, 1, 13, 721, 1440, generator=g)
    return inp, inp_s, (s_mean, s_std, u_mean, u_std), maps, const_h


def run(steps, warmup, threads=None):
    import torch
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    M = import_reference()
    torch.manual_seed(0)
    t0 = time.perf_counter()
    model = M.PanguModel(device="cpu").eval()
    t_init = time.perf_counter() - t0
    inp, inp_s, stats, maps, const_h = synth_inputs(torch)
    times, out = [], None
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = model(inp, inp_s, stats, maps, const_h)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    o, os_ = out
    try:
        cpu = [ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
    except (OSError, IndexError):
        cpu = "unknown"
    return {"kind": "reference", "cores": cores, "cpu_model": cpu, "torch": torch.__version__, "steps": steps,
            "warmup": warmup, "times_s": times, "seconds_per_forward": sum(times) / len(times),
            "best_s": min(times), "init_s": t_init, "params": sum(p.numel() for p in model.parameters()),
            "out_shapes": [list(o.shape), list(os_.shape)],
            "out_abs_mean": [float(o.abs().mean()), float(os_.abs().mean())]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    if not available():
        print(json.dumps({"unavailable": "baseline/_ref/models/{layers,pangu_model}.py missing "
                                         "(run __graft_entry__.build() where /root/reference exists)"}))
        sys.exit(0)
    print(json.dumps(run(a.steps, a.warmup, a.threads or None)))
