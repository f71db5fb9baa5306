"""GPU: BASELINE configs[2] -- the 24 h model chained 7 times on the device (pangu_b200.rollout.Rollout) against the
UNMODIFIED reference chained 7 times on the CPU (tests/golden/reference_rollout_goldens.npz, written by
make_rollout_golden.py: best_model(...) then normBackData fed back, inference/inference_mix_multiOutput.py:201-238).

fp32 mode must stay within 1e-4 rel-L2 at every one of the 7 steps (measured 4.9e-7 ... 5.9e-7); bf16 mode within
north_star's 2e-2 at every step; the growth curve is printed (measured: flat, 3.0e-3 -> 3.7e-3)."""
import os

import numpy as np
import pytest
import torch

import pangu_oracle as orc
import ref_loops
from util_gpu import check_digest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
BF16_STEP_TOL = (2e-2,) * 7       # north_star's bf16 bound at EVERY lead time (measured 3.0e-3 at step 1 -> 3.7e-3 / 4.1e-3 at step 7)


@pytest.fixture(scope="module")
def rg():
    p = os.path.join(HERE, "golden", "reference_rollout_goldens.npz")
    if not os.path.exists(p):
        pytest.skip("reference_rollout_goldens.npz not generated")
    return np.load(p, allow_pickle=False)


@pytest.mark.parametrize("mode,graph", [("fp32", False), ("bf16", True)])
def test_seven_step_rollout_vs_reference_chain(rg, mode, graph):
    from models.pangu_model import PanguModel
    from pangu_b200.rollout import Rollout
    steps = int(rg["steps"])
    assert steps == 7
    model = PanguModel(device="cpu")
    model.load_state_dict(orc.synth_params(seed=0), strict=True)
    model = model.cuda().eval().set_compute_dtype(mode)
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    last = ref_loops.statistics_last(stats)
    ro = Rollout(model, stats, last, maps, const_h, graph=graph)
    curve = []
    for k, (o, os_) in enumerate(ro.run(inp, inp_s, steps=steps)):
        tol = 1e-4 if mode == "fp32" else BF16_STEP_TOL[k]
        curve.append((check_digest(rg, f"step{k}.output", o, tol), check_digest(rg, f"step{k}.output_surface", os_, tol)))
    print(f"7-step rollout rel-L2 vs the reference chain ({mode}): " + "  ".join(f"{a:.2e}/{b:.2e}" for a, b in curve))
    assert len(curve) == 7
