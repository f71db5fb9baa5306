"""Pin the CPU oracle (oracle/pangu_oracle.py) against goldens produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

import pangu_oracle as orc

GEOMS = {"A24": (8, 181, 24), "B24": (8, 91, 24), "A": (8, 181, 360), "B": (8, 91, 180)}
FP32_TOL = 1e-5      # north_star: relative L2 <= 1e-5 for the fp32 path


def _check_digest(goldens, name, t, tol=FP32_TOL):
    flat = t.detach().reshape(-1)
    assert flat.numel() == int(goldens[name + ".numel"])
    pos = torch.from_numpy(goldens[name + ".pos"])
    ref = torch.from_numpy(goldens[name + ".val"])
    got = flat[pos]
    err = orc.rel_l2(got, ref)
    assert err <= tol, f"{name}: rel-L2 {err:.3e} > {tol}"
    n = float(flat.double().norm())
    assert abs(n - float(goldens[name + ".norm"])) <= 1e-4 * float(goldens[name + ".norm"])


@pytest.mark.parametrize("tag", list(GEOMS))
@pytest.mark.parametrize("roll", [False, True])
def test_window_source_index_bit_exact(goldens, tag, roll):
    Z, H, W = GEOMS[tag]
    src = orc.window_source_index(Z, H, W, roll)
    key = f"index.{tag}.roll{int(roll)}"
    assert tuple(goldens[key + ".shape"]) == src.shape
    assert bool(goldens[key + ".reverse_is_inverse"])
    assert hashlib.sha256(src.tobytes()).digest() == goldens[key + ".sha256"].tobytes()
    if key + ".src" in goldens:
        assert np.array_equal(goldens[key + ".src"].astype(np.int64), src)
    # every real token appears exactly once (partition is a bijection onto real tokens + pads)
    real = src[src >= 0]
    assert real.size == Z * H * W and np.array_equal(np.sort(real), np.arange(Z * H * W))


@pytest.mark.parametrize("tag", list(GEOMS))
def test_shift_mask_bit_exact(goldens, tag):
    Z, H, W = GEOMS[tag]
    m = orc.shift_mask(Z, H, W)
    assert bool(goldens[f"mask.{tag}.all_lon_identical"])
    assert tuple(goldens[f"mask.{tag}.shape"]) == m.shape
    assert sorted(goldens[f"mask.{tag}.values"].tolist()) == [-100.0, 0.0]
    bits = np.packbits((m != 0).reshape(-1))
    assert np.array_equal(bits, goldens[f"mask.{tag}.bits"])
    # the compact closed form used in the kernels induces the same mask
    gid = orc.shift_group_ids_closed_form(Z, H, W)
    assert np.array_equal(gid[:, None, :] != gid[:, :, None], m != 0)


def test_position_index_bit_exact(goldens):
    idx = orc.position_index()
    assert idx.dtype == np.int64 and idx.shape == (20736,)
    assert np.array_equal(idx, goldens["position_index"].astype(np.int64))
    assert idx.min() == 0 and idx.max() == 3311 and np.unique(idx).size == 3312


def test_param_contract_matches_reference_state_dict(goldens):
    mine = list(orc.param_shapes().keys())
    assert len(mine) == 223
    if "state_dict.keys" in goldens:
        assert mine == [str(k) for k in goldens["state_dict.keys"]]


@pytest.mark.parametrize("tag,dim,heads,Z,H,W,pfx", [
    ("blockA", 192, 6, 8, 181, 24, "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."),
    ("blockB", 384, 12, 8, 91, 24, "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock3."),
])
def test_block_matches_reference(goldens, tag, dim, heads, Z, H, W, pfx):
    params = orc.synth_params(seed=0, only_prefix=pfx)
    g = torch.Generator().manual_seed(7)
    if tag == "blockB":                       # replay the generator stream of make_golden.py
        torch.randn(1, 8 * 181 * 24, 192, generator=g)
    x = torch.randn(1, Z * H * W, dim, generator=g)
    assert abs(float(x.double().sum()) - float(goldens[f"{tag}.x.sum"])) < 1e-6
    for roll in (False, True):
        y = orc.earth_block(x, Z, H, W, roll, params, pfx, heads)
        _check_digest(goldens, f"{tag}.roll{int(roll)}", y)


def test_downsample_attention_mlp_match_reference(goldens):
    pfx = "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."
    params = orc.synth_params(seed=0, only_prefix=pfx)
    params.update(orc.synth_params(seed=0, only_prefix="downsample."))
    g = torch.Generator().manual_seed(7)
    torch.randn(1, 8 * 181 * 24, 192, generator=g)
    torch.randn(1, 8 * 91 * 24, 384, generator=g)
    x = torch.randn(1, 8 * 181 * 24, 192, generator=g)
    assert abs(float(x.double().sum()) - float(goldens["down24.x.sum"])) < 1e-6
    _check_digest(goldens, "down24", orc.down_sample(x, 8, 181, 24, params))
    xw = torch.randn(2, 124, 144, 192, generator=g)
    mask = torch.from_numpy(orc.shift_mask(8, 181, 24))
    _check_digest(goldens, "attnA.nomask", orc.window_attention(xw, None, params, pfx + "attention.", 6))
    _check_digest(goldens, "attnA.mask", orc.window_attention(xw, mask, params, pfx + "attention.", 6))
    _check_digest(goldens, "mlpA", orc.mlp(xw[0, :8], params, pfx + "linear."))


@pytest.mark.slow
def test_full_model_matches_reference(goldens):
    """Full 721x1440 forward of the oracle vs the stored subsample of the reference's output
    (and of every intermediate stage).  About a minute on 8 cores, ~10 GB."""
    if "output.val" not in goldens:
        pytest.skip("goldens were generated with --skip-full")
    params = orc.synth_params(seed=0)
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    assert abs(float(inp.double().sum() + inp_s.double().sum()) - float(goldens["inputs.sum"])) < 1e-3
    with torch.no_grad():
        x0 = orc.patch_embed(inp, inp_s, stats, maps, const_h, params)
        _check_digest(goldens, "embed", x0)
        x1 = orc.earth_layer(x0, 8, 181, 360, 2, params, "layers.EarthSpecificLayer0.", 6)
        _check_digest(goldens, "layer0", x1)
        x2 = orc.down_sample(x1, 8, 181, 360, params)
        _check_digest(goldens, "down", x2)
        x3 = orc.earth_layer(x2, 8, 91, 180, 6, params, "layers.EarthSpecificLayer1.", 12)
        _check_digest(goldens, "layer1", x3)
        x4 = orc.earth_layer(x3, 8, 91, 180, 6, params, "layers.EarthSpecificLayer2.", 12)
        _check_digest(goldens, "layer2", x4)
        x5 = orc.up_sample(x4, params)
        _check_digest(goldens, "up", x5)
        x6 = orc.earth_layer(x5, 8, 181, 360, 2, params, "layers.EarthSpecificLayer3.", 6)
        _check_digest(goldens, "layer3", x6)
        o, os_ = orc.patch_recover(torch.cat((x1, x6), dim=-1), 8, 181, 360, params)
        _check_digest(goldens, "output", o)
        _check_digest(goldens, "output_surface", os_)
