"""Latitude-band sharding on the GPU: all bands of a plan run in ONE process (LocalComm) and must reproduce
the un-sharded forward -- same kernels, same per-token arithmetic, only the tiling over rows differs."""
import pytest
import torch

import pangu_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model_and_inputs():
    from models.pangu_model import PanguModel
    model = PanguModel(device="cpu")
    model.load_state_dict(orc.synth_params(seed=0), strict=True)
    model = model.cuda().eval().set_compute_dtype("bf16")
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    args = (inp.cuda(), inp_s.cuda(), tuple(s.cuda() for s in stats), maps.cuda(), const_h.cuda())
    with torch.no_grad():
        want = model(*args)
    return model, args, want


@pytest.mark.parametrize("world,scheme", [(1, "redundant"), (2, "redundant"), (4, "redundant"), (8, "redundant"),
                                          (2, "sendback"), (8, "sendback")])
def test_banded_forward_equals_unsharded(model_and_inputs, world, scheme):
    from pangu_b200.dist import emulate_bands
    model, args, want = model_and_inputs
    out, out_s = emulate_bands(model, world, *args, scheme=scheme)
    assert out.shape == want[0].shape and out_s.shape == want[1].shape
    e0, e1 = orc.rel_l2(out.cpu(), want[0].cpu()), orc.rel_l2(out_s.cpu(), want[1].cpu())
    print(f"bands={world}: rel-L2 vs un-sharded: output {e0:.2e} surface {e1:.2e}")
    # the -100 shift mask is finite: the un-sharded kernel lets masked keys contribute exp(-100) ~ 4e-44,
    # the band of rank 0 sees the same keys, so the results agree to fp32 rounding
    assert e0 <= 1e-6 and e1 <= 1e-6


def test_band_attention_rejects_inconsistent_band():
    from pangu_b200 import ops
    from pangu_b200.abi import Band, PanguError
    qkv = torch.zeros(8 * 24 * 24, 3 * 192, dtype=torch.bfloat16, device="cuda")
    qb = torch.zeros(3 * 192, device="cuda")
    eb = torch.zeros(124, 6, 144, 144, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(PanguError):                     # rolled windows 0..3 need 3 halo rows
        ops.window_attention_band(qkv, None, qb, eb, 8, 181, 24, 6, Band(0, 24, 0, 4, 0, 0, 0), 1)


def test_cuda_graph_replay_equals_eager(model_and_inputs):
    """pangu_b200.graph.GraphedForward: the captured step replays to the same bits, for new inputs too."""
    from pangu_b200.graph import GraphedForward
    model, args, want = model_and_inputs
    fwd = GraphedForward(model, args)
    out, out_s = fwd(args[0], args[1])
    torch.cuda.synchronize()
    assert torch.equal(out, want[0]) and torch.equal(out_s, want[1])
    inp2, inp_s2 = args[0] * 0.5 + 0.1, args[1] * 0.5 - 0.2
    with torch.no_grad():
        want2 = model(inp2, inp_s2, *args[2:])
    out2, out_s2 = fwd(inp2, inp_s2)
    torch.cuda.synchronize()
    assert torch.equal(out2, want2[0]) and torch.equal(out_s2, want2[1])
    assert fwd.launches_per_replay > 50


def test_streamed_forecaster_matches_eager(model_and_inputs):
    """pangu_b200.pipeline.StreamedForecaster: overlapped H2D / forward / D2H returns each sample's own result."""
    from pangu_b200.pipeline import StreamedForecaster
    model, args, want = model_and_inputs
    sf = StreamedForecaster(model, args)
    hosts = []
    for k in range(4):
        a = (args[0] * (1.0 - 0.1 * k)).cpu().pin_memory()
        b = (args[1] * (1.0 + 0.05 * k)).cpu().pin_memory()
        hosts.append((a, b))
    got = []
    for a, b in hosts:
        r = sf.submit(a, b)
        if r is not None:
            got.append(tuple(t.clone() for t in r))
    got += [tuple(t.clone() for t in r) for r in sf.flush()]
    assert len(got) == 4
    for (a, b), (o, os_) in zip(hosts, got):
        with torch.no_grad():
            w, ws = model(a.cuda(), b.cuda(), *args[2:])
        assert torch.equal(o, w.cpu()) and torch.equal(os_, ws.cpu())
