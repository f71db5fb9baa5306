"""GPU: BASELINE configs[2] -- the 24 h model chained 7 times on the device (pangu_b200.rollout.Rollout) against the
UNMODIFIED reference chained 7 times on the CPU (tests/golden/reference_rollout_goldens.npz, written by
make_rollout_golden.py: best_model(...) then normBackData fed back, inference/inference_mix_multiOutput.py:201-238).

fp32 mode must stay within 1e-4 rel-L2 at every one of the 7 steps; bf16 mode must start inside north_star's 2e-2 and its
growth curve is printed and bounded (the chain feeds each step's rounding back through a random-init network)."""
import os

import numpy as np
import pytest
import torch

import pangu_oracle as orc
import ref_loops
from util_gpu import check_digest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
BF16_STEP_TOL = (2e-2, 3e-2, 4e-2, 5e-2, 6e-2, 7e-2, 8e-2)       # step 1 is north_star's bound; later steps: bounded growth


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
