"""Chained forecasts (BASELINE configs[2]): Rollout must equal model() + normBackData fed back by hand."""
import pytest
import torch

import pangu_oracle as orc

pytestmark = pytest.mark.gpu


def test_rollout_matches_manual_chain():
    from models.pangu_model import PanguModel
    from pangu_b200.rollout import Rollout
    model = PanguModel(device="cpu")
    model.load_state_dict(orc.synth_params(seed=0), strict=True)
    model = model.cuda().eval().set_compute_dtype("bf16")
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    g = torch.Generator().manual_seed(11)
    # statistics_last in the layout of era5_data.utils_data.weatherStatistics_output
    last = (torch.randn(1, 4, 1, 1, generator=g), torch.rand(1, 4, 1, 1, generator=g) + 0.5,
            torch.randn(1, 5, 13, 1, 1, generator=g), torch.rand(1, 5, 13, 1, 1, generator=g) + 0.5)
    stats_d = tuple(s.cuda() for s in stats)
    last_d = tuple(s.cuda() for s in last)
    maps_d, ch_d = maps.cuda(), const_h.cuda()

    def manual_step(a, b):                                  # the reference's step: model() then normBackData
        with torch.no_grad():
            o, os_ = model(a, b, stats_d, maps_d, ch_d)
            return o * last_d[3] + last_d[2], os_ * last_d[1] + last_d[0]

    for graph in (False, True):
        ro = Rollout(model, stats, last, maps, const_h, graph=graph)
        prev = (inp.cuda(), inp_s.cuda())
        n = 0
        for o, os_ in ro.run(inp, inp_s, steps=3):
            # each step is checked on the rollout's OWN previous state: the bf16 network amplifies the 1-ulp
            # difference between fmaf(x, std, mean) and x*std+mean to ~1e-3 after another forward, which is
            # path noise, not a rollout error
            want, want_s = manual_step(*prev)
            e0, e1 = orc.rel_l2(o.cpu(), want.cpu()), orc.rel_l2(os_.cpu(), want_s.cpu())
            assert e0 <= 1e-6 and e1 <= 1e-6, (graph, n, e0, e1)
            prev = (o.clone(), os_.clone())
            n += 1
        assert n == 3
