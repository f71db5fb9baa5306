"""Generate tests/golden/reference_rollout_goldens.npz: the UNMODIFIED reference PanguModel chained 7 times
(BASELINE configs[2]; inference/inference_mix_multiOutput.py:201-238: `best_model(...)` then
`utils_data.normBackData(output, output_surface, weather_statistics_last)` fed back as the next input).

Run in the build container only (needs /root/reference):

    python tests/golden/make_rollout_golden.py [--steps 7]

Weights: oracle.synth_params(0); inputs / statistics: oracle.synth_inputs(1); `weather_statistics_last` is derived
from the same statistics exactly as era5_data/utils_data.py:395-421 derives it from the .npy files (surface
view(1,4,1,1); upper [13,1,1,5] level-reversed, transposed to [1,5,13,1,1]).  Stored per step: a fixed pseudo-random
subsample (positions + fp32 values) and the L2 norm of `output` and `output_surface` in physical units.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import digest, import_reference, orc  # noqa: E402


def statistics_last(stats):
    sm, ss, um, us = stats
    f = lambda t: t.flip(0).permute(1, 3, 0, 2).unsqueeze(-1).contiguous()        # [13,1,1,5] -> [1,5,13,1,1], levels reversed
    return sm.view(1, 4, 1, 1), ss.view(1, 4, 1, 1), f(um), f(us)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=7)
    a = ap.parse_args()
    L, M = import_reference()
    params = orc.synth_params(seed=0)
    model = M.PanguModel(device="cpu")
    model.load_state_dict(params, strict=True)
    model.eval()
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    last = statistics_last(stats)
    out = {"steps": np.int64(a.steps)}
    with torch.no_grad():
        for k in range(a.steps):
            t0 = time.perf_counter()
            o, os_ = model(inp, inp_s, stats, maps, const_h)
            o = o * last[3] + last[2]                      # normBackData, era5_data/utils_data.py:540-546
            os_ = os_ * last[1] + last[0]
            digest(f"step{k}.output", o, out, count=8192)
            digest(f"step{k}.output_surface", os_, out, count=8192)
            inp, inp_s = o, os_
            print(f"step {k}: {time.perf_counter() - t0:.1f} s, |o| mean {float(o.abs().mean()):.4f}", flush=True)
    out["torch_version"] = np.array(torch.__version__)
    path = os.path.join(HERE, "reference_rollout_goldens.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
