"""CPU: the oracle's restatement of the training losses (models/pangu_sample.py:163-218) against goldens produced by
the reference's OWN train() loop (tests/golden/make_loss_golden.py), value and gradient, all four branches; and the
reference loops themselves run unchanged against stubs (boundary proof on the CPU box)."""
import os

import numpy as np
import pytest
import torch

import pangu_oracle as orc
import ref_loops

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lg():
    return np.load(os.path.join(HERE, "golden", "reference_loss_goldens.npz"), allow_pickle=False)


def _synth():
    from golden.make_loss_golden import synth
    return synth()


def test_wind_speed_matches_reference(lg):
    d = _synth()
    got = orc.wind_speed(d["out_s"], d["tgt_s"], d["out"], d["tgt"])
    for name, t in zip(("ws_out_s", "ws_tgt_s", "ws_out", "ws_tgt"), got):
        assert np.array_equal(t.numpy(), lg["wind." + name]), name


@pytest.mark.parametrize("wind", [False, True])
@pytest.mark.parametrize("masked", [False, True])
def test_training_loss_matches_reference_train_loop(lg, wind, masked):
    d = _synth()
    uw, sw, ulw, slw = ref_loops.variable_weights()
    o, os_ = d["out"].clone().requires_grad_(), d["out_s"].clone().requires_grad_()
    loss = orc.training_loss(o, os_, d["tgt"], d["tgt_s"], (d["sm"], d["ss"], d["um"], d["us"]), uw, sw, ulw, slw,
                             only_use_wind_speed_loss=wind, custom_mask=d["mask"] if masked else None)
    loss.backward()
    tag = f"loss.wind{int(wind)}.mask{int(masked)}"
    assert abs(float(loss) - float(lg[tag + ".value"])) <= 1e-6 * max(1.0, abs(float(loss)))     # logged with 6 decimals
    assert np.allclose(o.grad.numpy(), lg[tag + ".d_out"], rtol=1e-6, atol=1e-9)
    assert np.allclose(os_.grad.numpy(), lg[tag + ".d_out_s"], rtol=1e-6, atol=1e-9)


def test_reference_loops_import_and_run_unchanged_on_stubs(tmp_path):
    """models/pangu_sample.py loads from its file (build container: /root/reference; GPU box: baseline/_ref) with the
    era5_data stubs and its train() / test() run unchanged on a stand-in model -- the same harness the GPU test uses with
    the B200 PanguModel (tests/test_gpu_reference_loops.py)."""
    if ref_loops.find_pangu_sample() is None:
        pytest.skip("neither /root/reference nor baseline/_ref/models/pangu_sample.py is present")
    H, W = 16, 32
    g = torch.Generator().manual_seed(3)
    stats_last = (torch.randn(1, 4, 1, 1, generator=g), torch.rand(1, 4, 1, 1, generator=g) + 0.5,
                  torch.randn(1, 5, 13, 1, 1, generator=g), torch.rand(1, 5, 13, 1, 1, generator=g) + 0.5)
    consts = {"weather_statistics": None, "weather_statistics_last": stats_last, "constant_maps": None, "const_h": None,
              "variable_weights": ref_loops.variable_weights(), "custom_mask": None}

    class Score:                                                          # era5_data.score surface used by test()
        weighted_rmse_torch_channels = staticmethod(orc.weighted_rmse_channels)
        weighted_acc_torch_channels = staticmethod(orc.weighted_acc_channels)

    rec = ref_loops.install(consts, Score)
    ps = ref_loops.load_pangu_sample()
    assert all(hasattr(ps, n) for n in ("train", "test", "get_wind_speed"))

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Parameter(torch.ones(()))

        def forward(self, inp, inp_s, stats, maps, const_h):
            return inp * self.a, inp_s * self.a

    model = Tiny()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[25, 50], gamma=0.5)
    sample = lambda: (torch.randn(1, 5, 13, H, W, generator=g), torch.randn(1, 4, H, W, generator=g),
                      torch.randn(1, 5, 13, H, W, generator=g), torch.randn(1, 4, H, W, generator=g),
                      [["2018010100"], ["2018010200"]])
    log = ref_loops.ListLogger()
    ps.train(model, [sample(), sample()], [sample()], opt, sched, str(tmp_path), "cpu", None, log, 1)
    assert float(model.a) != 1.0 and "loss=" in log.lines[0]
    ps.test([sample()], model, "cpu", str(tmp_path))
    assert len(rec.saved) == 2 and rec.saved[0][-1] == "rmse" and rec.saved[1][-1] == "acc"
    rmse_z = rec.saved[0][1]["2018010200"]
    assert rmse_z.shape == (13,) and np.isfinite(rmse_z).all()
