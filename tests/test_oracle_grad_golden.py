"""The oracle's autograd (torch.autograd over oracle/pangu_oracle.py -- what the GPU backward tests compare against)
replayed against gradient digests of the REFERENCE itself (tests/golden/make_grad_golden.py imported
/root/reference's modules, ran forward + backward on the same seeded tensors).  fp32 CPU on both sides:
rel-L2 <= 2e-5 on the sampled values, 1e-4 relative on the norms (different summation orders only)."""
import os

import numpy as np
import pytest
import torch

import pangu_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 2e-5


@pytest.fixture(scope="module")
def gg():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_grad_goldens.npz"), allow_pickle=False)


def check(gg, name, t):
    flat = t.detach().reshape(-1)
    assert flat.numel() == int(gg[name + ".numel"]), name
    got = flat[torch.from_numpy(gg[name + ".pos"])].float()
    ref = torch.from_numpy(gg[name + ".val"])
    assert orc.rel_l2(got, ref) <= TOL, (name, orc.rel_l2(got, ref))
    assert abs(float(flat.double().norm()) - float(gg[name + ".norm"])) <= 1e-4 * float(gg[name + ".norm"]) + 1e-12, name


def same(a, b):
    return abs(float(a.double().sum()) - float(b)) < 1e-6


def test_oracle_autograd_reproduces_reference_gradients(gg):
    # the golden maker draws all its tensors from ONE generator in a fixed order; replay the same order
    g = torch.Generator().manual_seed(77)
    params = orc.synth_params(seed=0)
    for tag, dim, heads, Z, H, W, pfx in (
            ("blockA", 192, 6, 8, 181, 24, "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."),
            ("blockB", 384, 12, 8, 91, 24, "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock3.")):
        x = torch.randn(1, Z * H * W, dim, generator=g)
        r = torch.randn(1, Z * H * W, dim, generator=g)
        assert same(x, gg[f"{tag}.x.sum"]) and same(r, gg[f"{tag}.r.sum"])
        pd = {k: v for k, v in params.items() if k.startswith(pfx)}
        for roll in (False, True):
            got = orc.grads(lambda lv: orc.earth_block(lv["x"], Z, H, W, roll, lv, pfx, heads), {"x": x, **pd}, r)
            check(gg, f"{tag}.roll{int(roll)}.d.x", got["x"])
            for k in pd:
                check(gg, f"{tag}.roll{int(roll)}.d.{k[len(pfx):]}", got[k])
    # down-sample (W = 24)
    x = torch.randn(1, 8 * 181 * 24, 192, generator=g)
    r = torch.randn(1, 8 * 91 * 12, 384, generator=g)
    assert same(x, gg["down24.x.sum"])
    pd = {k: v for k, v in params.items() if k.startswith("downsample.")}
    got = orc.grads(lambda lv: orc.down_sample(lv["x"], 8, 181, 24, lv), {"x": x, **pd}, r)
    check(gg, "down24.d.x", got["x"])
    for k in pd:
        check(gg, "down24.d." + k[len("downsample."):], got[k])
    # up-sample (the reference hard-codes 8 x 91 x 180)
    x = torch.randn(1, 8 * 91 * 180, 384, generator=g)
    r = torch.randn(1, 8 * 181 * 360, 192, generator=g)
    assert same(x, gg["up.x.sum"])
    pd = {k: v for k, v in params.items() if k.startswith("upsample.")}
    got = orc.grads(lambda lv: orc.up_sample(lv["x"], lv), {"x": x, **pd}, r)
    check(gg, "up.d.x", got["x"])
    for k in pd:
        check(gg, "up.d." + k[len("upsample."):], got[k])
    # patch embedding / recovery (full resolution)
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    r = torch.randn(1, 8 * 181 * 360, 192, generator=g)
    pd = {k: v for k, v in params.items() if k.startswith("_input_layer.")}
    got = orc.grads(lambda lv: orc.patch_embed(inp, inp_s, stats, maps, const_h, lv), pd, r)
    for k in pd:
        check(gg, "embed.d." + k[len("_input_layer."):], got[k])
    x = torch.randn(1, 8 * 181 * 360, 384, generator=g)
    ro, rs = torch.randn(1, 5, 13, 721, 1440, generator=g), torch.randn(1, 4, 721, 1440, generator=g)
    assert same(x, gg["recover.x.sum"])
    pd = {k: v for k, v in params.items() if k.startswith("_output_layer.")}
    got = orc.grads(lambda lv: orc.patch_recover(lv["x"], 8, 181, 360, lv), {"x": x, **pd}, (ro, rs))
    check(gg, "recover.d.x", got["x"])
    for k in pd:
        check(gg, "recover.d." + k[len("_output_layer."):], got[k])
