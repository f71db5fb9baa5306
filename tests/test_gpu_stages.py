"""GPU: the stage-level goldens of the reference's full-resolution forward (embed, layer0, down, layer1, layer2, up,
layer3 in tests/golden/reference_goldens.npz, written by make_golden.py from the UNMODIFIED reference) replayed on the
CUDA path module by module, exactly as make_golden.py calls the reference's modules -- so a full-model failure
localises.  fp32 mode: rel-L2 <= 1e-5; bf16 mode: <= 2e-2 (north_star's tolerances) at every stage."""
import pytest
import torch

import pangu_oracle as orc
from util_gpu import check_digest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_stage_goldens_on_the_cuda_path(goldens, mode, tol):
    if "layer3.val" not in goldens:
        pytest.skip("goldens were generated with --skip-full")
    from models.pangu_model import PanguModel
    model = PanguModel(device="cpu")
    model.load_state_dict(orc.synth_params(seed=0), strict=True)
    model = model.cuda().eval().set_compute_dtype(mode)
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    errs = {}
    with torch.no_grad():
        x0 = model._input_layer(inp.cuda(), inp_s.cuda(), tuple(s.cuda() for s in stats), maps.cuda(), const_h.cuda())
        errs["embed"] = check_digest(goldens, "embed", x0, tol)
        x1 = model.layers[0](x0, 8, 181, 360)
        errs["layer0"] = check_digest(goldens, "layer0", x1, tol)
        x2 = model.downsample(x1, 8, 181, 360)
        errs["down"] = check_digest(goldens, "down", x2, tol)
        x3 = model.layers[1](x2, 8, 91, 180)
        errs["layer1"] = check_digest(goldens, "layer1", x3, tol)
        x4 = model.layers[2](x3, 8, 91, 180)
        errs["layer2"] = check_digest(goldens, "layer2", x4, tol)
        x5 = model.upsample(x4)
        errs["up"] = check_digest(goldens, "up", x5, tol)
        x6 = model.layers[3](x5, 8, 181, 360)
        errs["layer3"] = check_digest(goldens, "layer3", x6, tol)
        o, os_ = model._output_layer(torch.cat((x1, x6), dim=-1), 8, 181, 360)
        errs["output"] = check_digest(goldens, "output", o, tol)
        errs["output_surface"] = check_digest(goldens, "output_surface", os_, tol)
    print(f"stage rel-L2 vs reference goldens ({mode}): " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
