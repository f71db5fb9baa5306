"""Either side of the forward: latitude-weighted RMSE / ACC: oracle pinned to outputs of the reference's era5_data/score.py
(tests/golden/make_score_golden.py), CUDA one-pass kernel against the oracle and the goldens."""
import os

import numpy as np
import pytest
import torch

import pangu_oracle as orc

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_score_goldens.npz"))
T = lambda k: torch.from_numpy(G[k])


def test_oracle_scores_match_reference_goldens():
    assert np.array_equal(orc.latitude_weights(33).numpy(), G["lat_weight"])               # same fp32 expression
    for p, t, r, rm, a in (("pred3", "targ3", "rmse3", "rmse3_masked", "acc3"), ("pred4", "targ4", "rmse4", "rmse4_masked", "acc4")):
        np.testing.assert_allclose(orc.weighted_rmse_channels(T(p), T(t)).numpy(), G[r], rtol=2e-6)
        np.testing.assert_allclose(orc.weighted_rmse_channels(T(p), T(t), T("mask")).numpy(), G[rm], rtol=2e-6)
        np.testing.assert_allclose(orc.weighted_acc_channels(T(p), T(t)).numpy(), G[a], rtol=2e-6)
    np.testing.assert_allclose(orc.weighted_rmse_channels(T("pred4"), T("targ4")).mean(0).numpy(), G["rmse4_mean"], rtol=2e-6)


def test_score_module_refuses_cpu_tensors():
    from pangu_b200 import score
    from pangu_b200.abi import PanguError
    with pytest.raises(PanguError, match="CUDA"):
        score.weighted_rmse_torch_channels(T("pred3"), T("targ3"))


@pytest.mark.gpu
def test_cuda_scores_match_reference_goldens():
    from pangu_b200 import score
    assert torch.equal(score.latitude_weights(33, "cuda").cpu(), T("lat_weight"))
    for p, t, r, rm, a in (("pred3", "targ3", "rmse3", "rmse3_masked", "acc3"), ("pred4", "targ4", "rmse4", "rmse4_masked", "acc4")):
        pc, tc = T(p).cuda(), T(t).cuda()
        np.testing.assert_allclose(score.weighted_rmse_torch_channels(pc, tc).cpu().numpy(), G[r], rtol=1e-5)   # tolerance: fp32 sums
        np.testing.assert_allclose(score.weighted_rmse_torch_channels(pc, tc, T("mask").cuda()).cpu().numpy(), G[rm], rtol=1e-5)
        np.testing.assert_allclose(score.weighted_acc_torch_channels(pc, tc).cpu().numpy(), G[a], rtol=1e-5)
    np.testing.assert_allclose(score.weighted_rmse_torch(T("pred4").cuda(), T("targ4").cuda()).cpu().numpy(), G["rmse4_mean"], rtol=1e-5)


@pytest.mark.gpu
def test_cuda_scores_full_resolution_against_the_oracle():
    """The 69 planes of one forecast (5 x 13 upper-air + 4 surface) at 721 x 1440, with climatology and mask."""
    from pangu_b200 import score
    g = torch.Generator().manual_seed(5)
    pred = torch.randn(69, 721, 1440, generator=g) * 2.0 + 3.0
    targ = pred + 0.3 * torch.randn(69, 721, 1440, generator=g)
    clim = torch.randn(69, generator=g) + 3.0
    mask = (torch.rand(721, 1440, generator=g) > 0.25).float()
    pc, tc = pred.cuda(), targ.cuda()
    rmse, acc = score.scores(pc, tc, clim=clim.cuda())
    rmse_m, _ = score.scores(pc, tc, mask=mask.cuda())
    a, b = pred.double() - clim.double()[:, None, None], targ.double() - clim.double()[:, None, None]
    np.testing.assert_allclose(rmse.cpu().numpy(), orc.weighted_rmse_channels(pred.double(), targ.double()).numpy(), rtol=1e-5)
    np.testing.assert_allclose(rmse_m.cpu().numpy(), orc.weighted_rmse_channels(pred.double(), targ.double(), mask.double()).numpy(), rtol=1e-5)
    np.testing.assert_allclose(acc.cpu().numpy(), orc.weighted_acc_channels(a, b).numpy(), rtol=1e-5)
    # size-independent properties: a perfect forecast has RMSE 0 and ACC 1; scaling the error scales the RMSE
    r0, a0 = score.scores(pc, pc, clim=clim.cuda())
    assert float(r0.abs().max()) == 0.0 and float((a0 - 1).abs().max()) <= 1e-6
    r2, _ = score.scores(pc, pc + 2.0 * (tc - pc))
    np.testing.assert_allclose(r2.cpu().numpy(), 2.0 * rmse.cpu().numpy(), rtol=1e-5)
    # one pass: 2 x 286 MB read
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        score.score_sums(pc, tc, clim=clim.cuda())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("lat_weighted_score_sums: %.3f ms, %.0f GB/s" % (ms, 2 * pred.numel() * 4 / ms / 1e6))


# ------------------------------------------------------------------------------------------ input pipeline
def test_prefetcher_refuses_to_run_without_cuda():
    from pangu_b200.abi import PanguError
    from pangu_b200.prefetch import DataPrefetcher
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(PanguError, match="CUDA"):
        DataPrefetcher([(torch.zeros(1),) * 5])


@pytest.mark.gpu
def test_prefetcher_returns_every_batch_in_order_and_wraps_around():
    """Contract of era5_data/utils_data.py:20-57 (next() -> 5 device tensors, len(), wrap-around), minus its skipped batch."""
    from pangu_b200.prefetch import DataPrefetcher
    g = torch.Generator().manual_seed(11)
    batches = [(torch.randn(1, 5, 13, 40, 64, generator=g), torch.randn(1, 4, 40, 64, generator=g),
                torch.randn(1, 5, 13, 40, 64, generator=g), torch.randn(1, 4, 40, 64, generator=g),
                torch.tensor([[2018010100 + i, 2018010200 + i]])) for i in range(3)]
    pf = DataPrefetcher(batches)
    assert len(pf) == 3
    side = torch.cuda.Stream()
    for i in range(7):                                     # 7 > 2 * len: wraps around twice, pinned slots are reused
        with torch.cuda.stream(side if i % 2 else torch.cuda.current_stream()):
            got = pf.next()
            assert len(got) == 5 and all(t.is_cuda for t in got)
            acc = [t.clone() for t in got]                 # consumed on the caller's stream
        torch.cuda.synchronize()
        for a, want in zip(acc, batches[i % 3]):
            assert torch.equal(a.cpu(), want)


# ------------------------------------------------------------------------------------------ weights
def test_oracle_compact_bias_gather_is_the_reference_formula():
    """models/layers.py:442-449 (commented there): table[position_index].view(144, 144, T, heads).permute(2, 3, 0, 1)[None]."""
    rng = np.random.default_rng(3)
    table = rng.standard_normal((3312, 5, 2)).astype(np.float32)
    full = orc.expand_bias_table(table)
    assert full.shape == (1, 5, 2, 144, 144)
    idx = torch.from_numpy(orc.position_index())
    want = torch.from_numpy(table)[idx].view(144, 144, 5, 2).permute(2, 3, 0, 1).unsqueeze(0)
    assert torch.equal(torch.from_numpy(full), want)
    # every table row is used, by 12 - |d| pairs
    cnt = np.bincount(orc.position_index(), minlength=3312)
    assert cnt.min() == 1 and cnt.max() == 12 and cnt.sum() == 144 * 144


def test_load_named_weights_follows_the_onnx2torch_shape_rules():
    """models/onnx2torch.py:124-161: 1-D / 3-D / 5-D copied, 2-D transposed, rows without a source name skipped, frozen."""
    from pangu_b200.abi import PanguError
    from pangu_b200.weights import load_named_weights

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(3, 5)
            self.conv = torch.nn.Conv1d(4, 6, kernel_size=1)
            self.bias5 = torch.nn.Parameter(torch.zeros(1, 2, 3, 4, 4))
            self.untouched = torch.nn.Parameter(torch.ones(2))

    m = Tiny()
    rng = np.random.default_rng(0)
    w = {"a": rng.standard_normal((3, 5)).astype(np.float32), "b": rng.standard_normal(5).astype(np.float32),
         "c": rng.standard_normal((6, 4, 1)).astype(np.float32), "d": rng.standard_normal(6).astype(np.float32),
         "e": rng.standard_normal((1, 2, 3, 4, 4)).astype(np.float32)}
    name_map = [("lin.weight", "a"), ("lin.bias", "b"), ("conv.weight", "c"), ("conv.bias", "d"), ("bias5", "e"), ("untouched", float("nan"))]
    assert load_named_weights(m, w, name_map) == 5
    assert np.array_equal(m.lin.weight.detach().numpy(), w["a"].T) and np.array_equal(m.conv.weight.detach().numpy(), w["c"])
    assert np.array_equal(m.bias5.detach().numpy(), w["e"]) and float(m.untouched.sum()) == 2.0
    assert not m.lin.weight.requires_grad and m.untouched.requires_grad
    with pytest.raises(PanguError, match="gives"):
        load_named_weights(Tiny(), {**w, "b": np.zeros(4, np.float32)}, name_map)


@pytest.mark.gpu
@pytest.mark.parametrize("T,heads", [(124, 6), (64, 12), (3, 1)])
def test_cuda_compact_bias_expand_is_bit_exact_and_reduce_is_its_adjoint(T, heads):
    from pangu_b200.weights import expand_bias_table, reduce_bias_grad
    g = torch.Generator().manual_seed(T)
    table = torch.randn(3312, T, heads, generator=g)
    full = expand_bias_table(table.cuda())
    assert np.array_equal(full.cpu().numpy(), orc.expand_bias_table(table.numpy()))                      # index work: bit-exact
    d_full = torch.randn(1, T, heads, 144, 144, generator=g)
    d_table = reduce_bias_grad(d_full.cuda())
    want = torch.zeros(3312, T * heads, dtype=torch.float64)
    want.index_add_(0, torch.from_numpy(orc.position_index()), d_full[0].double().permute(2, 3, 0, 1).reshape(144 * 144, T * heads))
    np.testing.assert_allclose(d_table.cpu().double().reshape(3312, -1).numpy(), want.numpy(), rtol=0, atol=2e-5)   # <= 12 fp32 terms
    lhs = float((full.cpu().double() * d_full.double()).sum())                                           # <expand(t), g> = <t, reduce(g)>
    rhs = float((table.double() * d_table.cpu().double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * max(1.0, abs(lhs))
    acc = reduce_bias_grad(d_full.cuda(), d_table.clone())                                               # accumulates
    np.testing.assert_allclose(acc.cpu().numpy(), 2 * d_table.cpu().numpy(), rtol=1e-6)
