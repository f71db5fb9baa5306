"""Generate tests/golden/reference_goldens.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not travel to the GPU box):

    python tests/golden/make_golden.py [--skip-full]

The reference (models/layers.py, models/pangu_model.py) is imported as-is; the two imports that
cannot be satisfied offline are stubbed exactly as SURVEY Appendix C describes:
`timm.models.layers.{DropPath, trunc_normal_}` and an empty `era5_data.utils_data`.

What is stored (all small):
  * index goldens: partition / reverse source maps obtained by pushing token ids through the
    reference EarthSpecificBlock with an identity attention (full arrays for narrow W=24,
    sha256 for the full geometries), the shift mask of gen_mask (bit-packed), position_index;
  * float goldens: for each module and for the whole model a fixed pseudo-random SUBSAMPLE of the
    reference output plus its L2 norm, computed on seeded synthetic weights/inputs that the oracle's
    `synth_params` / `synth_inputs` reproduce bit-identically on any machine with the same torch.
"""
import argparse
import hashlib
import os
import sys
import time
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pangu_oracle as orc  # noqa: E402

REF = "/root/reference"


def import_reference():
    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                m.div_(keep)
            return x * m

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return torch.nn.init.trunc_normal_(t, mean, std, a, b)

    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")
    timm_layers.DropPath, timm_layers.trunc_normal_ = DropPath, trunc_normal_
    timm.models, timm_models.layers = timm_models, timm_layers
    sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers})
    era5 = types.ModuleType("era5_data")
    era5.__path__ = []
    era5_utils = types.ModuleType("era5_data.utils_data")
    era5.utils_data = era5_utils
    sys.modules.update({"era5_data": era5, "era5_data.utils_data": era5_utils})
    sys.path.insert(0, REF)
    import models.layers as L           # noqa
    import models.pangu_model as M      # noqa
    return L, M


def sample_positions(n, count, seed):
    rng = np.random.RandomState(seed)
    return np.sort(rng.choice(n, size=min(count, n), replace=False)).astype(np.int64)


def digest(name, t, out, count=4096, seed=1234):
    flat = t.detach().reshape(-1).double()
    pos = sample_positions(flat.numel(), count, seed)
    out[name + ".pos"] = pos
    out[name + ".val"] = flat[pos].float().numpy()
    out[name + ".norm"] = np.float64(flat.norm().item())
    out[name + ".sum"] = np.float64(flat.sum().item())
    out[name + ".numel"] = np.int64(flat.numel())


class _Identity3(nn.Module):
    """Stands in for EarthAttention3D: records the partitioned windows, returns them unchanged."""

    def __init__(self):
        super().__init__()
        self.seen = None

    def forward(self, x, mask):
        self.seen = x.detach().clone()
        self.mask = None if mask is None else mask.detach().clone()
        return x


def index_goldens(L, out):
    for tag, dim, Z, H, W in (("A24", 192, 8, 181, 24), ("B24", 384, 8, 91, 24),
                              ("A", 192, 8, 181, 360), ("B", 384, 8, 91, 180)):
        # build a block without the 60 MB bias table: swap the attention afterwards
        blk = L.EarthSpecificBlock.__new__(L.EarthSpecificBlock)
        nn.Module.__init__(blk)
        blk.device = "cpu"
        blk.window_size = (2, 6, 12)
        blk.drop_path = nn.Identity()
        blk.norm1, blk.norm2 = nn.Identity(), nn.Identity()
        rec_after = {}

        class _Zero(nn.Module):
            def forward(self, x):
                rec_after["x"] = x.detach().clone()      # tensor after reverse/unroll/crop
                return torch.zeros_like(x)
        blk.linear = _Zero()
        blk.attention = _Identity3()
        blk.padding_front, blk.padding_back = 0, 5
        blk.type_of_windows = (8 // 2) * ((H + 5) // 6)
        N = Z * H * W
        ids = (torch.arange(N, dtype=torch.float32) + 1).reshape(1, N, 1)
        for roll in (False, True):
            y = blk.forward(ids, Z, H, W, roll)
            win = blk.attention.seen                      # [nLon, T, 144, 1] holding id+1, 0 = pad
            src = win[..., 0].to(torch.int64).numpy() - 1
            # reverse must be the exact inverse on real tokens: norm1==identity so
            # rec_after = shortcut + reversed  ->  reversed == ids
            rev_ok = bool(torch.equal(rec_after["x"] - ids, ids))
            key = f"index.{tag}.roll{int(roll)}"
            out[key + ".sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(src).tobytes()).digest(), dtype=np.uint8)
            out[key + ".shape"] = np.array(src.shape, dtype=np.int64)
            out[key + ".reverse_is_inverse"] = np.bool_(rev_ok)
            if W == 24:
                out[key + ".src"] = src.astype(np.int32)
            if roll:
                m = blk.attention.mask                    # [nLon, T, 144, 144]
                same = bool((m == m[0:1]).all())
                out[f"mask.{tag}.all_lon_identical"] = np.bool_(same)
                vals = torch.unique(m).numpy()
                out[f"mask.{tag}.values"] = vals.astype(np.float32)
                out[f"mask.{tag}.bits"] = np.packbits((m[0] != 0).numpy().reshape(-1))
                out[f"mask.{tag}.shape"] = np.array(m[0].shape, dtype=np.int64)
        print("index goldens", tag, "done")
    att = L.EarthAttention3D.__new__(L.EarthAttention3D)
    nn.Module.__init__(att)
    att.device, att.window_size = "cpu", (2, 6, 12)
    att._construct_index()
    out["position_index"] = att.position_index.numpy().astype(np.int16)


def load_into(module, params, prefix):
    sd = {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}
    missing, unexpected = module.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return module.eval()


def float_goldens(L, M, out, skip_full):
    torch.manual_seed(0)
    params = orc.synth_params(seed=0)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        # blocks at narrow width (W=24): stage A / B, un-rolled / rolled
        for tag, dim, heads, Z, H, W, pfx in (
                ("blockA", 192, 6, 8, 181, 24, "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."),
                ("blockB", 384, 12, 8, 91, 24, "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock3.")):
            blk = load_into(L.EarthSpecificBlock(dim, 0.0, heads, "cpu"), params, pfx)
            x = torch.randn(1, Z * H * W, dim, generator=g)
            out[f"{tag}.x.sum"] = np.float64(x.double().sum().item())
            for roll in (False, True):
                y = blk(x, Z, H, W, roll)
                digest(f"{tag}.roll{int(roll)}", y, out)
            print("float goldens", tag, "done")
        # down-sample at narrow width
        ds = load_into(L.DownSample(192), params, "downsample.")
        x = torch.randn(1, 8 * 181 * 24, 192, generator=g)
        out["down24.x.sum"] = np.float64(x.double().sum().item())
        digest("down24", ds(x, 8, 181, 24), out)
        # attention + mlp alone (narrow)
        att = load_into(L.EarthAttention3D(192, 6, 0, (2, 6, 12), "cpu"), params,
                        "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1.attention.")
        xw = torch.randn(2, 124, 144, 192, generator=g)
        mask = torch.from_numpy(orc.shift_mask(8, 181, 24)).unsqueeze(0).expand(2, -1, -1, -1)
        digest("attnA.nomask", att(xw, None), out)
        digest("attnA.mask", att(xw, mask), out)
        ml = load_into(L.Mlp(192, 0), params, "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1.linear.")
        digest("mlpA", ml(xw[0, :8]), out)
        if skip_full:
            return
        # full-size-only modules and the whole model
        inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
        out["inputs.sum"] = np.float64(inp.double().sum().item() + inp_s.double().sum().item())
        model = M.PanguModel(device="cpu")
        assert list(model.state_dict().keys()) == list(params.keys()) or \
            set(model.state_dict().keys()) == set(params.keys())
        out["state_dict.keys"] = np.array(list(model.state_dict().keys()))
        model.load_state_dict(params, strict=True)
        model.eval()
        t0 = time.perf_counter()
        x0 = model._input_layer(inp, inp_s, stats, maps, const_h)
        digest("embed", x0, out)
        x1 = model.layers[0](x0, 8, 181, 360)
        digest("layer0", x1, out)
        x2 = model.downsample(x1, 8, 181, 360)
        digest("down", x2, out)
        x3 = model.layers[1](x2, 8, 91, 180)
        digest("layer1", x3, out)
        x4 = model.layers[2](x3, 8, 91, 180)
        digest("layer2", x4, out)
        x5 = model.upsample(x4)
        digest("up", x5, out)
        x6 = model.layers[3](x5, 8, 181, 360)
        digest("layer3", x6, out)
        o, os_ = model._output_layer(torch.cat((x1, x6), dim=-1), 8, 181, 360)
        digest("output", o, out, count=16384)
        digest("output_surface", os_, out, count=16384)
        print("staged full forward %.1f s" % (time.perf_counter() - t0))
        t0 = time.perf_counter()
        o2, os2 = model(inp, inp_s, stats, maps, const_h)
        dt = time.perf_counter() - t0
        assert torch.equal(o, o2) and torch.equal(os_, os2)
        out["full_forward_seconds"] = np.float64(dt)
        out["full_forward_threads"] = np.int64(torch.get_num_threads())
        print("reference full forward %.1f s" % dt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-full", action="store_true")
    args = ap.parse_args()
    L, M = import_reference()
    out = {}
    index_goldens(L, out)
    float_goldens(L, M, out, args.skip_full)
    out["torch_version"] = np.array(torch.__version__)
    path = os.path.join(HERE, "reference_goldens.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB,", len(out), "entries")


if __name__ == "__main__":
    main()
