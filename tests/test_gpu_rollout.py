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


def test_ort_like_session_equals_rollout_step_and_chains_on_device():
    """InferenceSession.run (the onnxruntime call shape of inference/inference_singleOutput.py:146-147) must return what
    Rollout.step returns for the same numpy fields, and feeding its outputs back (the scripts' chained-forecast loop) must
    equal re-uploading them."""
    import numpy as np
    from models.pangu_model import PanguModel
    from pangu_b200.rollout import Rollout
    from pangu_b200.session import InferenceSession
    model = PanguModel(device="cpu")
    model.load_state_dict(orc.synth_params(seed=0), strict=True)
    model = model.cuda().eval().set_compute_dtype("bf16")
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    g = torch.Generator().manual_seed(11)
    last = (torch.randn(1, 4, 1, 1, generator=g), torch.rand(1, 4, 1, 1, generator=g) + 0.5,
            torch.randn(1, 5, 13, 1, 1, generator=g), torch.rand(1, 5, 13, 1, 1, generator=g) + 0.5)
    x = inp.numpy().astype(np.float32).squeeze()              # as the scripts prepare them: [5,13,721,1440], [4,721,1440]
    xs = inp_s.numpy().astype(np.float32).squeeze()
    sess = InferenceSession(model, stats, last, maps, const_h)
    out, out_s = sess.run(None, {"input": x, "input_surface": xs})
    assert out.shape == (5, 13, 721, 1440) and out_s.shape == (4, 721, 1440) and out.dtype == np.float32
    assert sess.h2d_bytes == x.nbytes + xs.nbytes
    want, want_s = Rollout(model, stats, last, maps, const_h, graph=False).step(inp.cuda()[0], inp_s.cuda()[0])
    assert np.array_equal(out, want[0].cpu().numpy()) and np.array_equal(out_s, want_s[0].cpu().numpy())
    # chained call on the returned arrays: stays on the device ...
    out2, out2_s = sess.run(None, {"input": out, "input_surface": out_s})
    assert sess.h2d_bytes == 0
    # ... and equals a call on copies of them (which goes through the host)
    out2b, out2b_s = sess.run(["output_surface", "output"], {"input": out.copy(), "input_surface": out_s.copy()})[::-1]
    assert sess.h2d_bytes > 0
    assert np.array_equal(out2, out2b) and np.array_equal(out2_s, out2b_s)
    assert np.isfinite(out2).all() and not np.array_equal(out2, out)
    # the chained fast path may never drop an edit: outputs are read-only, an edited copy is re-uploaded
    assert not out.flags.writeable and not out_s.flags.writeable
    with pytest.raises(ValueError):
        out[0, 0, 0, 0] = 0.0
    edited = np.clip(out, None, float(np.median(out)))
    out3, _ = sess.run(None, {"input": edited, "input_surface": out_s})
    assert sess.h2d_bytes > 0 and not np.array_equal(out3, out2)
    plain = InferenceSession(model, stats, last, maps, const_h, chain_on_device=False)
    o1, o1_s = plain.run(None, {"input": x, "input_surface": xs})
    assert o1.flags.writeable and np.array_equal(o1, out)
    plain.run(None, {"input": o1, "input_surface": o1_s})
    assert plain.h2d_bytes > 0                               # plain ORT behaviour: every call uploads
