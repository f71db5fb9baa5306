"""Generate tests/golden/reference_loss_goldens.npz by running the reference's OWN train() loop
(models/pangu_sample.py:96-235) for one iteration on a tiny stand-in model whose two parameters ARE the model
outputs, so that after `loss.backward()` their `.grad` is d loss / d output; the epoch loss is read from the
logger line train() prints.  All four loss branches: default / custom mask / wind speed / wind speed + custom mask
(:183-204), plus get_wind_speed (:74-93) itself.

Run in the build container only (needs /root/reference):   python tests/golden/make_loss_golden.py
"""
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_loops  # noqa: E402

H, W = 32, 64


def synth(seed=5):
    g = torch.Generator().manual_seed(seed)
    d = {"out": torch.randn(1, 5, 13, H, W, generator=g), "out_s": torch.randn(1, 4, H, W, generator=g),
         "tgt": torch.randn(1, 5, 13, H, W, generator=g) * 3 + 1, "tgt_s": torch.randn(1, 4, H, W, generator=g) * 2 - 1,
         "um": torch.randn(1, 5, 13, 1, 1, generator=g), "us": torch.rand(1, 5, 13, 1, 1, generator=g) + 0.5,
         "sm": torch.randn(1, 4, 1, 1, generator=g), "ss": torch.rand(1, 4, 1, 1, generator=g) + 0.5,
         "mask": (torch.rand(H, W, generator=g) > 0.35).float()}
    return d


class OutputsAsParameters(torch.nn.Module):
    def __init__(self, out, out_s):
        super().__init__()
        self.out, self.out_s = torch.nn.Parameter(out.clone()), torch.nn.Parameter(out_s.clone())

    def forward(self, *a):
        return self.out, self.out_s


def main():
    d = synth()
    consts = {"weather_statistics": None, "weather_statistics_last": (d["sm"], d["ss"], d["um"], d["us"]),
              "constant_maps": None, "const_h": None, "variable_weights": ref_loops.variable_weights(), "custom_mask": d["mask"]}
    ref_loops.install(consts, score_module=None)
    ps = ref_loops.load_pangu_sample("/root/reference/models/pangu_sample.py")
    out = {}
    ws = ps.get_wind_speed(d["out_s"], d["tgt_s"], d["out"], d["tgt"])
    for name, t in zip(("ws_out_s", "ws_tgt_s", "ws_out", "ws_tgt"), ws):
        out["wind." + name] = t.numpy()
    for wind in (False, True):
        for masked in (False, True):
            model = OutputsAsParameters(d["out"], d["out_s"])
            opt = torch.optim.SGD(model.parameters(), lr=0.0)
            sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[25, 50], gamma=0.5)
            log = ref_loops.ListLogger()
            loader = [(torch.zeros(1), torch.zeros(1), d["tgt"], d["tgt_s"], [["2018010100"], ["2018010200"]])]
            ps.train(model, loader, loader, opt, sched, "/tmp/ref_loss_golden", "cpu", None, log, 1,
                     only_use_wind_speed_loss=wind, use_custom_mask=masked)
            loss = float(re.search(r"loss=([0-9.eE+-]+)", log.lines[0]).group(1))
            tag = f"loss.wind{int(wind)}.mask{int(masked)}"
            out[tag + ".value"] = np.float64(loss)
            out[tag + ".d_out"] = model.out.grad.numpy().copy()
            out[tag + ".d_out_s"] = model.out_s.grad.numpy().copy()
            print(tag, loss, float(model.out.grad.abs().sum()), float(model.out_s.grad.abs().sum()))
    out["torch_version"] = np.array(torch.__version__)
    path = os.path.join(HERE, "reference_loss_goldens.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
