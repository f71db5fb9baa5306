"""GPU: (1) the fused loss kernels (pangu_b200.loss) against goldens produced by the reference's own train() loop, all
four branches of models/pangu_sample.py:183-204, value and gradient; (2) boundary proof: the reference's train() and
test() (models/pangu_sample.py:96-235, :391-575), loaded from the git-ignored baseline/_ref copy and run UNCHANGED on the
B200 PanguModel with a synthetic loader."""
import copy
import os
import re

import numpy as np
import pytest
import torch

import pangu_oracle as orc
import ref_loops

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lg():
    return np.load(os.path.join(HERE, "golden", "reference_loss_goldens.npz"), allow_pickle=False)


@pytest.mark.parametrize("wind", [False, True])
@pytest.mark.parametrize("masked", [False, True])
def test_fused_loss_matches_reference_train_loop(lg, wind, masked):
    """Tolerance: fp32 sums in a different order -> value 2e-6 relative; gradients are products of the same few fp32
    factors -> 1e-6 relative (sign(0) = 0 on both sides)."""
    from golden.make_loss_golden import synth
    from pangu_b200.loss import training_loss
    d = {k: v.cuda() for k, v in synth().items()}
    o, os_ = d["out"].clone().requires_grad_(), d["out_s"].clone().requires_grad_()
    loss = training_loss(o, os_, d["tgt"], d["tgt_s"], (d["sm"], d["ss"], d["um"], d["us"]),
                         only_use_wind_speed_loss=wind, custom_mask=d["mask"] if masked else None)
    (loss * 3.0).backward()
    tag = f"loss.wind{int(wind)}.mask{int(masked)}"
    want = float(lg[tag + ".value"])
    assert abs(float(loss) - want) <= 2e-6 * abs(want) + 1e-6, (float(loss), want)
    for got, name in ((o.grad, ".d_out"), (os_.grad, ".d_out_s")):
        ref = torch.from_numpy(lg[tag + name]).cuda() * 3.0
        assert float((got - ref).abs().max()) <= 1e-6 * float(ref.abs().max()), (tag, name)


def test_fused_losses_full_resolution_vs_oracle():
    """721 x 1440 (BASELINE sizes): every branch against the oracle's plain-torch restatement evaluated on the GPU."""
    from pangu_b200.loss import training_loss
    g = torch.Generator().manual_seed(9)
    dev = "cuda"
    o = torch.randn(1, 5, 13, 721, 1440, generator=g).to(dev)
    os_ = torch.randn(1, 4, 721, 1440, generator=g).to(dev)
    t = (torch.randn(1, 5, 13, 721, 1440, generator=g) * 3 + 1).to(dev)
    ts = (torch.randn(1, 4, 721, 1440, generator=g) * 2 - 1).to(dev)
    last = tuple(x.to(dev) for x in (torch.randn(1, 4, 1, 1, generator=g), torch.rand(1, 4, 1, 1, generator=g) + 0.5,
                                     torch.randn(1, 5, 13, 1, 1, generator=g), torch.rand(1, 5, 13, 1, 1, generator=g) + 0.5))
    mask = (torch.rand(721, 1440, generator=g) > 0.4).float().to(dev)
    uw, sw, ulw, slw = ref_loops.variable_weights(dev)
    for wind in (False, True):
        for m in (None, mask):
            a, b = o.clone().requires_grad_(), os_.clone().requires_grad_()
            got = training_loss(a, b, t, ts, last, only_use_wind_speed_loss=wind, custom_mask=m)
            got.backward()
            a2, b2 = o.clone().requires_grad_(), os_.clone().requires_grad_()
            want = orc.training_loss(a2, b2, t, ts, last, uw, sw, ulw, slw, only_use_wind_speed_loss=wind, custom_mask=m)
            want.backward()
            assert abs(float(got) - float(want)) <= 2e-5 * abs(float(want)), (wind, m is not None)
            assert orc.rel_l2(a.grad, a2.grad) <= 1e-6 and orc.rel_l2(b.grad, b2.grad) <= 1e-6


def _b200_model(drop_path=True):
    from models.pangu_model import PanguModel
    model = PanguModel(device="cpu")
    model.load_state_dict(orc.synth_params(seed=0), strict=True)
    if not drop_path:
        for m in model.modules():
            if hasattr(m, "drop_prob"):
                m.drop_prob = 0.0
    return model.cuda().set_compute_dtype("bf16")


def _sample(seed):
    g = torch.Generator().manual_seed(seed)
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=seed)
    last = ref_loops.statistics_last(stats)
    tgt = torch.randn(1, 5, 13, 721, 1440, generator=g) * last[3] + last[2]          # physical units, like the dataset
    tgt_s = torch.randn(1, 4, 721, 1440, generator=g) * last[1] + last[0]
    return (inp, inp_s, tgt, tgt_s, [["2018010100"], ["2018010200"]]), stats, last, maps, const_h


def test_reference_train_and_test_loops_run_unchanged_on_the_b200_model(tmp_path):
    if ref_loops.find_pangu_sample() is None:
        pytest.skip("baseline/_ref/models/pangu_sample.py did not travel (run __graft_entry__.build() where /root/reference exists)")
    from pangu_b200 import score as b200_score
    from pangu_b200.loss import weighted_l1_loss
    s1, stats, last, maps, const_h = _sample(1)
    s2 = _sample(2)[0]
    consts = {"weather_statistics": stats, "weather_statistics_last": last, "constant_maps": maps, "const_h": const_h,
              "variable_weights": ref_loops.variable_weights(), "custom_mask": None}
    rec = ref_loops.install(consts, b200_score)                           # era5_data.score -> the product's mirror
    ps = ref_loops.load_pangu_sample()
    dev = torch.device("cuda:0")

    # (a) one iteration with lr = 0 and DropPath off: the loss train() logs and the gradients its loss.backward() leaves
    #     must equal the product's fused loss + backward on the same model
    model = _b200_model(drop_path=False)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[25, 50], gamma=0.5)
    log = ref_loops.ListLogger()
    ps.train(model, [s1], [s1], opt, sched, str(tmp_path), dev, None, log, 1)
    ref_loss = float(re.search(r"loss=([0-9.eE+-]+)", log.lines[0]).group(1))
    ref_grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    assert len(ref_grads) == 223 and all(torch.isfinite(g).all() for g in ref_grads.values())
    model.zero_grad(set_to_none=True)
    model.train()
    d_stats = tuple(s.to(dev) for s in stats)
    o, os_ = model(s1[0].to(dev), s1[1].to(dev), d_stats, maps.to(dev), const_h.to(dev))
    loss = weighted_l1_loss(o, os_, s1[2].to(dev), s1[3].to(dev), tuple(t.to(dev) for t in last))
    loss.backward()
    assert abs(float(loss) - ref_loss) <= 2e-5 * abs(ref_loss) + 1e-6, (float(loss), ref_loss)
    worst = max(orc.rel_l2(p.grad, ref_grads[k]) for k, p in model.named_parameters())
    assert worst <= 5e-3, worst            # same backward kernels: 1-ulp differences of the two loss gradients flip a few bf16
                                           # roundings, split-K wgrad atomics reorder fp32 sums (measured 1.3e-3)
    del o, os_, loss, ref_grads
    model.zero_grad(set_to_none=True)

    # (b) two iterations of the real thing: train() mode with DropPath, Adam as in era5_data/config.py:45-46
    model = _b200_model()
    before = copy.deepcopy(model.layers[1].blocks[0].attention.linear1.weight.detach())
    opt = torch.optim.Adam(model.parameters(), lr=2e-5, weight_decay=3e-6)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[25, 50], gamma=0.5)
    log = ref_loops.ListLogger()
    ps.train(model, [s1, s2], [s1], opt, sched, str(tmp_path), dev, None, log, 1)
    m = re.search(r"loss=([0-9.eE+-]+)", log.lines[0])
    assert m and np.isfinite(float(m.group(1))) and float(m.group(1)) > 0
    assert not torch.equal(before, model.layers[1].blocks[0].attention.linear1.weight.detach())
    del opt

    # (c) test() on one batch: the RMSE / ACC tables it saves, against the oracle's scores of the same forecast
    ps.test([s1], model, dev, str(tmp_path))
    assert len(rec.saved) == 2 and rec.saved[0][-1] == "rmse" and rec.saved[1][-1] == "acc"
    model.eval()
    o, os_ = model(s1[0].to(dev), s1[1].to(dev), d_stats, maps.to(dev), const_h.to(dev))     # grad mode on, like test()
    o = (o.detach() * last[3].to(dev) + last[2].to(dev)).squeeze()
    os_ = (os_.detach() * last[1].to(dev) + last[0].to(dev)).squeeze()
    tgt, tgt_s = s1[2].squeeze(), s1[3].squeeze()
    key = "2018010200"
    want_z = orc.weighted_rmse_channels(o[0].cpu(), tgt[0]).numpy()
    want_sfc = orc.weighted_rmse_channels(os_.cpu(), tgt_s).numpy()
    assert np.allclose(rec.saved[0][1][key], want_z, rtol=1e-4) and np.allclose(rec.saved[0][7][key], want_sfc, rtol=1e-4)
    want_acc_t = orc.weighted_acc_channels(o[2].cpu() - last[2][0, 2], tgt[2] - last[2][0, 2]).numpy()
    assert np.allclose(rec.saved[1][3][key], want_acc_t, rtol=1e-4, atol=1e-5)
