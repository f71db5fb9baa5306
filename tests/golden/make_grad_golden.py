"""Gradient goldens of the REFERENCE (fine-tune path): imports /root/reference in the build container
(same stubs as make_golden.py), runs module forwards + torch.autograd backward on seeded inputs / output gradients and
stores digests (sampled values + norm + sum) of every gradient.  tests/test_oracle_golden.py replays them against the
oracle's autograd; the GPU tests compare the CUDA backward with the oracle.

    python tests/golden/make_grad_golden.py        ->  tests/golden/reference_grad_goldens.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (import_reference, digest, load_into; puts oracle/ on sys.path)

orc = mg.orc


def grad_digests(tag, module, inputs, outs_fn, gouts, out, prefix=""):
    module.zero_grad(set_to_none=True)
    outs = outs_fn()
    if isinstance(outs, torch.Tensor):
        outs, gouts = (outs,), (gouts,)
    torch.autograd.backward(outs, gouts)
    for name, t in inputs.items():
        mg.digest(f"{tag}.d.{name}", t.grad, out)
    for name, p in module.named_parameters():
        mg.digest(f"{tag}.d.{prefix}{name}", p.grad, out)


def main():
    L, M = mg.import_reference()
    out = {}
    params = orc.synth_params(seed=0)
    g = torch.Generator().manual_seed(77)
    for tag, dim, heads, Z, H, W, pfx in (
            ("blockA", 192, 6, 8, 181, 24, "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."),
            ("blockB", 384, 12, 8, 91, 24, "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock3.")):
        blk = mg.load_into(L.EarthSpecificBlock(dim, 0.0, heads, "cpu"), params, pfx)
        x = torch.randn(1, Z * H * W, dim, generator=g)
        r = torch.randn(1, Z * H * W, dim, generator=g)
        out[f"{tag}.x.sum"] = np.float64(x.double().sum().item())
        out[f"{tag}.r.sum"] = np.float64(r.double().sum().item())
        for roll in (False, True):
            xin = x.clone().requires_grad_()
            grad_digests(f"{tag}.roll{int(roll)}", blk, {"x": xin}, lambda: blk(xin, Z, H, W, roll), r, out)
        print("grad goldens", tag, "done")
    ds = mg.load_into(L.DownSample(192), params, "downsample.")
    x = torch.randn(1, 8 * 181 * 24, 192, generator=g).requires_grad_()
    r = torch.randn(1, 8 * 91 * 12, 384, generator=g)
    out["down24.x.sum"], out["down24.r.sum"] = np.float64(x.double().sum().item()), np.float64(r.double().sum().item())
    grad_digests("down24", ds, {"x": x}, lambda: ds(x, 8, 181, 24), r, out)
    us = mg.load_into(L.UpSample(384, 192), params, "upsample.")
    x = torch.randn(1, 8 * 91 * 180, 384, generator=g).requires_grad_()
    r = torch.randn(1, 8 * 181 * 360, 192, generator=g)
    out["up.x.sum"], out["up.r.sum"] = np.float64(x.double().sum().item()), np.float64(r.double().sum().item())
    grad_digests("up", us, {"x": x}, lambda: us(x), r, out)
    print("grad goldens down / up done")
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    pe = mg.load_into(L.PatchEmbedding_pretrain((2, 4, 4), 192), params, "_input_layer.")
    r = torch.randn(1, 8 * 181 * 360, 192, generator=g)
    out["embed.r.sum"] = np.float64(r.double().sum().item())
    grad_digests("embed", pe, {}, lambda: pe(inp, inp_s, stats, maps, const_h), r, out)
    pr = mg.load_into(L.PatchRecovery_pretrain(384), params, "_output_layer.")
    x = torch.randn(1, 8 * 181 * 360, 384, generator=g).requires_grad_()
    ro, rs = torch.randn(1, 5, 13, 721, 1440, generator=g), torch.randn(1, 4, 721, 1440, generator=g)
    out["recover.x.sum"] = np.float64(x.double().sum().item())
    out["recover.r.sum"] = np.float64(ro.double().sum().item() + rs.double().sum().item())
    grad_digests("recover", pr, {"x": x}, lambda: pr(x, 8, 181, 360), (ro, rs), out)
    print("grad goldens embed / recover done")
    out["torch_version"] = np.array(torch.__version__)
    path = os.path.join(HERE, "reference_grad_goldens.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB,", len(out), "entries")


if __name__ == "__main__":
    main()
