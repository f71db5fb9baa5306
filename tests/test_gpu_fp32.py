"""fp32 CUDA path (FFMA kernels) against the CPU oracle and the reference-generated goldens.
Tolerance: relative L2 <= 1e-5 (north_star)."""
import numpy as np
import pytest
import torch

import pangu_oracle as orc
from util_gpu import check_digest, load_params

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def L():
    import models.layers as layers
    return layers


def test_linear_matches_torch(L):
    from pangu_b200 import ops
    from pangu_b200.abi import ACT_GELU
    g = torch.Generator().manual_seed(0)
    for M, K, N in ((1000, 192, 576), (333, 112, 192), (257, 768, 160), (129, 1536, 64)):
        a, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05, torch.randn(N, generator=g)
        want = torch.nn.functional.gelu(a.double() @ w.double().t() + b.double())
        got = ops.linear(a.cuda(), w.cuda(), b.cuda(), act=ACT_GELU).cpu()
        assert orc.rel_l2(got, want) <= 2e-6
        want = a.double() @ w.double().t()
        got = ops.linear(a.cuda(), w.cuda(), None).cpu()
        assert orc.rel_l2(got, want) <= 2e-6


@pytest.mark.parametrize("tag,dim,heads,Z,H,W,pfx", [
    ("blockA", 192, 6, 8, 181, 24, "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."),
    ("blockB", 384, 12, 8, 91, 24, "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock3."),
])
def test_block_vs_oracle_and_golden(L, goldens, tag, dim, heads, Z, H, W, pfx):
    params = orc.synth_params(seed=0, only_prefix=pfx)
    blk = load_params(L.EarthSpecificBlock(dim, 0.0, heads, "cpu"), params, pfx)
    L.set_compute_dtype(blk, "fp32")
    g = torch.Generator().manual_seed(7)
    if tag == "blockB":
        torch.randn(1, 8 * 181 * 24, 192, generator=g)
    x = torch.randn(1, Z * H * W, dim, generator=g)
    for roll in (False, True):
        with torch.no_grad():
            y = blk(x.cuda(), Z, H, W, roll)
        want = orc.earth_block(x, Z, H, W, roll, params, pfx, heads)
        assert orc.rel_l2(y.cpu(), want) <= TOL
        check_digest(goldens, f"{tag}.roll{int(roll)}", y, TOL)


def test_attention_mlp_downsample_modules(L, goldens):
    pfx = "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."
    params = orc.synth_params(seed=0, only_prefix=pfx)
    params.update(orc.synth_params(seed=0, only_prefix="downsample."))
    g = torch.Generator().manual_seed(7)
    torch.randn(1, 8 * 181 * 24, 192, generator=g)
    torch.randn(1, 8 * 91 * 24, 384, generator=g)
    x = torch.randn(1, 8 * 181 * 24, 192, generator=g)
    ds = L.set_compute_dtype(load_params(L.DownSample(192), params, "downsample."), "fp32")
    with torch.no_grad():
        check_digest(goldens, "down24", ds(x.cuda(), 8, 181, 24), TOL)
        xw = torch.randn(2, 124, 144, 192, generator=g)
        att = L.set_compute_dtype(load_params(L.EarthAttention3D(192, 6, 0, (2, 6, 12), "cpu"), params, pfx + "attention."), "fp32")
        mask = torch.from_numpy(orc.shift_mask(8, 181, 24)).unsqueeze(0).expand(2, -1, -1, -1).cuda()
        check_digest(goldens, "attnA.nomask", att(xw.cuda(), None), TOL)
        check_digest(goldens, "attnA.mask", att(xw.cuda(), mask), TOL)
        ml = L.set_compute_dtype(load_params(L.Mlp(192, 0), params, pfx + "linear."), "fp32")
        check_digest(goldens, "mlpA", ml(xw[0, :8].cuda()), TOL)


def test_full_model_fp32_vs_reference_golden(goldens):
    """Full 721x1440 forward on the GPU vs the stored subsample of the reference's CPU fp32 output."""
    if "output.val" not in goldens:
        pytest.skip("goldens were generated with --skip-full")
    from models.pangu_model import PanguModel
    params = orc.synth_params(seed=0)
    model = PanguModel(device="cpu")
    model.load_state_dict(params, strict=True)
    model = model.cuda().eval().set_compute_dtype("fp32")
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    assert abs(float(inp.double().sum() + inp_s.double().sum()) - float(goldens["inputs.sum"])) < 1e-3
    stats = tuple(s.cuda() for s in stats)
    with torch.no_grad():
        out, out_s = model(inp.cuda(), inp_s.cuda(), stats, maps.cuda(), const_h.cuda())
    assert out.shape == (1, 5, 13, 721, 1440) and out_s.shape == (1, 4, 721, 1440)
    e1 = check_digest(goldens, "output", out, TOL)
    e2 = check_digest(goldens, "output_surface", out_s, TOL)
    print(f"fp32 full model rel-L2: output {e1:.2e} surface {e2:.2e}")
