"""CPU oracle for the Pangu-Weather forward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *restatement* (numpy index closed forms + plain torch
fp32 CPU math) of the algorithm in the reference repo comdaze/pangu-pytorch-demo:

    models/layers.py        (PatchEmbedding_pretrain, EarthSpecificBlock, EarthAttention3D,
                             Mlp, DownSample, UpSample, PatchRecovery_pretrain)
    models/pangu_model.py   (PanguModel.forward)

It exists so that the CUDA path can be checked without the reference being present
(`/root/reference` does not exist on the GPU box).  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg may import it; the product
(`pangu-pytorch-demo_b200/`) never does.

Parity pin: the reference has no tests or golden vectors of its own, so the
oracle is pinned against *outputs of the reference itself*, generated in the build container by
`tests/golden/make_golden.py` (imports /root/reference with a timm / era5_data stub) and
committed as `tests/golden/*.npz`.  `tests/test_oracle_golden.py` replays them.

Everything here is written for batch = 1: the reference's window-reverse does `view(1, ...)`
(models/layers.py:269) and is only correct for B = 1.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

WIN = (2, 6, 12)          # models/layers.py:27,168  window (Z, H, W)
WIN_TOKENS = 144          # 2*6*12
PAD_H = 5                 # models/layers.py:178  padding_back
HEAD_DIM = 32             # dim // heads for both stages (192/6, 384/12)
MASK_VALUE = -100.0       # models/layers.py:213-214


# --------------------------------------------------------------------------------------
# Index closed forms (integer work: must be bit-exact)
# --------------------------------------------------------------------------------------
def window_geometry(Z: int, H: int, W: int) -> Dict[str, int]:
    """Window-grid sizes of one stage.  models/layers.py:228 (pad 5), :253-262 (partition)."""
    Hp = H + PAD_H
    assert Z % WIN[0] == 0 and Hp % WIN[1] == 0 and W % WIN[2] == 0, (Z, H, W)
    nZ, nH, nLon = Z // WIN[0], Hp // WIN[1], W // WIN[2]
    return dict(Z=Z, H=H, W=W, Hp=Hp, nZ=nZ, nH=nH, nLon=nLon, T=nZ * nH)


def window_source_index(Z: int, H: int, W: int, roll: bool) -> np.ndarray:
    """For every window element (l, t, k) the flat source token n = (z*H + h)*W + w it is read
    from, or -1 when the element is a zero pad row.

    Restates F.pad (models/layers.py:228) + torch.roll shifts (-1,-3,-6) (:237-238) + the
    view/permute/reshape partition (:253-262).  Returns int64 [nLon, T, 144].
    """
    g = window_geometry(Z, H, W)
    Hp, nZ, nH, nLon = g["Hp"], g["nZ"], g["nH"], g["nLon"]
    l = np.arange(nLon).reshape(nLon, 1, 1, 1, 1, 1)
    zw = np.arange(nZ).reshape(1, nZ, 1, 1, 1, 1)
    hw = np.arange(nH).reshape(1, 1, nH, 1, 1, 1)
    dz = np.arange(WIN[0]).reshape(1, 1, 1, WIN[0], 1, 1)
    dh = np.arange(WIN[1]).reshape(1, 1, 1, 1, WIN[1], 1)
    dw = np.arange(WIN[2]).reshape(1, 1, 1, 1, 1, WIN[2])
    z = 2 * zw + dz
    h = 6 * hw + dh
    w = 12 * l + dw
    if roll:
        z = (z + WIN[0] // 2) % Z
        h = (h + WIN[1] // 2) % Hp
        w = (w + WIN[2] // 2) % W
    n = (z * H + h) * W + w
    n = np.where(h < H, n, -1)
    n = np.broadcast_to(n, (nLon, nZ, nH, WIN[0], WIN[1], WIN[2]))
    return np.ascontiguousarray(n.reshape(nLon, nZ * nH, WIN_TOKENS)).astype(np.int64)


def shift_region_ids(Z: int, H: int, W: int) -> np.ndarray:
    """Region id of every window element on the *rolled padded* grid, int32 [T, 144]
    (identical for every longitude window).

    Restates EarthSpecificBlock.gen_mask's slice painting (models/layers.py:187-202):
    z slices [0:-2], [-2:-1], [-1:]; h slices [0:-6], [6:-3], [-3:] -- note the second h slice
    starts at +6, not -6 (:197) -- painted in that order with a running counter; later
    assignments overwrite earlier ones; W is never split.
    """
    g = window_geometry(Z, H, W)
    Hp, nZ, nH = g["Hp"], g["nZ"], g["nH"]
    img = np.zeros((Z, Hp), dtype=np.int32)
    z_slices = (slice(0, -WIN[0]), slice(-WIN[0], -WIN[0] // 2), slice(-WIN[0] // 2, None))
    h_slices = (slice(0, -WIN[1]), slice(WIN[1], -WIN[1] // 2), slice(-WIN[1] // 2, None))
    cnt = 0
    for zs in z_slices:
        for hs in h_slices:
            img[zs, hs] = cnt
            cnt += 1
    ids = img.reshape(nZ, WIN[0], nH, WIN[1]).transpose(0, 2, 1, 3)      # zw, hw, dz, dh
    ids = np.repeat(ids[..., None], WIN[2], axis=-1)                    # ... dw
    return np.ascontiguousarray(ids.reshape(nZ * nH, WIN_TOKENS))


def shift_group_ids_closed_form(Z: int, H: int, W: int) -> np.ndarray:
    """Attention-equivalent compact group id used by the CUDA kernels:
    g(k) = 2*[zw==nZ-1]*dz + [hw==nH-1]*[dh>=3]   (SURVEY Appendix A).  int32 [T, 144].
    Two elements are masked apart iff their ids differ -- same partition as shift_region_ids."""
    g = window_geometry(Z, H, W)
    nZ, nH = g["nZ"], g["nH"]
    out = np.zeros((nZ, nH, WIN[0], WIN[1], WIN[2]), dtype=np.int32)
    dz = np.arange(WIN[0]).reshape(WIN[0], 1, 1)
    dh = np.arange(WIN[1]).reshape(1, WIN[1], 1)
    for zw in range(nZ):
        for hw in range(nH):
            out[zw, hw] = 2 * int(zw == nZ - 1) * dz + int(hw == nH - 1) * (dh >= 3)
    return out.reshape(nZ * nH, WIN_TOKENS)


def shift_mask(Z: int, H: int, W: int) -> np.ndarray:
    """Additive attention mask float32 [T, 144, 144]: -100 where region ids differ else 0.
    models/layers.py:212-214.  (The reference materialises it per longitude window; all
    longitude windows are identical.)"""
    ids = shift_region_ids(Z, H, W)
    diff = ids[:, None, :] != ids[:, :, None]
    return np.where(diff, np.float32(MASK_VALUE), np.float32(0.0)).astype(np.float32)


def position_index() -> np.ndarray:
    """EarthAttention3D._construct_index (models/layers.py:371-411), int64 [20736]:
    idx[i*144+j] = (zi + 2*zj)*828 + (hi + 6*hj)*23 + (wi - wj + 11)."""
    k = np.arange(WIN_TOKENS)
    z, h, w = k // 72, (k // 12) % 6, k % 12
    zi, zj = z[:, None], z[None, :]
    hi, hj = h[:, None], h[None, :]
    wi, wj = w[:, None], w[None, :]
    idx = (zi + 2 * zj) * (23 * 36) + (hi + 6 * hj) * 23 + (wi - wj + 11)
    return idx.reshape(-1).astype(np.int64)


def expand_bias_table(table: np.ndarray) -> np.ndarray:
    """The compact-bias gather the reference keeps as commented code (models/layers.py:442-449): table [3312, T, heads] ->
    table[position_index].view(144, 144, T, heads).permute(2, 3, 0, 1).unsqueeze(0) = [1, T, heads, 144, 144]."""
    T, heads = table.shape[1:]
    return np.ascontiguousarray(table[position_index()].reshape(WIN_TOKENS, WIN_TOKENS, T, heads).transpose(2, 3, 0, 1))[None]


# --------------------------------------------------------------------------------------
# Floating-point modules (fp32, CPU)
# --------------------------------------------------------------------------------------
def _p(params: Dict[str, torch.Tensor], prefix: str, name: str) -> torch.Tensor:
    return params[prefix + name]


def patch_embed(input, input_surface, statistics, maps, const_h, params, prefix="_input_layer."):
    """PatchEmbedding_pretrain.forward, models/layers.py:53-120 -> tokens [1, 8*181*360, dim]."""
    surface_mean, surface_std, upper_mean, upper_std = statistics
    B = input.shape[0]
    assert B == 1
    Hs, Ws = input_surface.shape[-2:]
    Hpad = Hs + 3                                             # :37 pads H by 3 (721 -> 724)
    # surface: (x - mean) / std per variable (:65)
    s = (input_surface[0] - surface_mean.reshape(-1, 1, 1)) / surface_std.reshape(-1, 1, 1)
    s = F.pad(s, (0, 0, 0, 3))
    s = torch.cat((s, maps.reshape(-1, Hpad, Ws)), dim=0)     # 7 channels (:75)
    h4, w4 = Hpad // 4, Ws // 4
    # feature = c*16 + ph*4 + pw  (:79-87); token (h', w')
    s = s.reshape(7, h4, 4, w4, 4).permute(1, 3, 0, 2, 4).reshape(h4 * w4, 7 * 16)
    tok_s = s @ _p(params, prefix, "conv_surface.weight")[:, :, 0].t() + _p(params, prefix, "conv_surface.bias")

    # upper air: stats are indexed by the flipped level (:95-99)
    um = torch.flip(upper_mean.reshape(13, 5), [0]).t().reshape(5, 13, 1, 1)
    us = torch.flip(upper_std.reshape(13, 5), [0]).t().reshape(5, 13, 1, 1)
    u = (input[0] - um) / us
    u = torch.cat((u, const_h.reshape(1, 13, Hs, Ws)), dim=0)           # 6 channels (:101)
    u = F.pad(u, (0, 0, 0, 3, 0, 1))                                    # Z 13->14, H 721->724 (:49)
    # feature = c*32 + pz*16 + ph*4 + pw (:107-112); token (z', h', w')
    u = u.reshape(6, 7, 2, h4, 4, w4, 4).permute(1, 3, 5, 0, 2, 4, 6).reshape(7 * h4 * w4, 6 * 32)
    tok_u = u @ _p(params, prefix, "conv.weight")[:, :, 0].t() + _p(params, prefix, "conv.bias")
    return torch.cat((tok_s, tok_u), dim=0).unsqueeze(0)                # z=0 surface plane first (:116)


def window_attention(xw, mask, params, prefix, heads):
    """EarthAttention3D.forward, models/layers.py:413-484.
    xw [nLon, T, 144, C]; mask [T,144,144] or None -> [nLon, T, 144, C]."""
    nLon, T, L, C = xw.shape
    w1, b1 = params[prefix + "linear1.weight"], params[prefix + "linear1.bias"]
    w2, b2 = params[prefix + "linear2.weight"], params[prefix + "linear2.bias"]
    bias = params[prefix + "earth_specific_bias"][0]                     # [T, heads, 144, 144]
    scale = (C // heads) ** -0.5                                         # :338
    out = torch.empty_like(xw)
    step = 6 if C <= 192 else 5                                          # bound the score tensor
    for s in range(0, nLon, step):
        x = xw[s:s + step]
        qkv = (x @ w1.t() + b1).reshape(x.shape[0], T, L, 3, heads, C // heads)
        qkv = qkv.permute(3, 0, 1, 4, 2, 5)                              # :426
        q, k, v = qkv[0] * scale, qkv[1], qkv[2]
        att = q @ k.transpose(-2, -1)
        att = att + bias.unsqueeze(0)                                    # :453
        if mask is not None:
            att = att + mask.reshape(1, T, 1, L, L)                      # :461-462
        att = torch.softmax(att, dim=-1)
        y = (att @ v).permute(0, 1, 3, 2, 4).reshape(x.shape[0], T, L, C)
        out[s:s + step] = y @ w2.t() + b2
    return out


def mlp(x, params, prefix):
    """Mlp.forward, models/layers.py:311-317 (exact erf GELU)."""
    h = F.gelu(x @ params[prefix + "linear1.weight"].t() + params[prefix + "linear1.bias"])
    return h @ params[prefix + "linear2.weight"].t() + params[prefix + "linear2.bias"]


def layer_norm(x, params, prefix):
    return F.layer_norm(x, (x.shape[-1],), params[prefix + "weight"], params[prefix + "bias"], 1e-5)


def earth_block(x, Z, H, W, roll, params, prefix, heads, s1=1.0, s2=1.0):
    """EarthSpecificBlock.forward, models/layers.py:218-299.  s1 / s2: what DropPath multiplies the two residual
    branches of this (single) sample by -- 1 in eval mode; in training 0 or 1/keep (timm DropPath, per-sample
    Bernoulli mask divided by keep, models/layers.py:171,296-297).  Differentiable torch code: the fine-tune
    gradients of the reference are torch.autograd over exactly these ops."""
    assert x.shape[0] == 1
    C = x.shape[-1]
    src = torch.from_numpy(window_source_index(Z, H, W, roll)).to(x.device)   # [nLon, T, 144]
    padded = torch.cat((x[0], x.new_zeros(1, C)), dim=0)                 # index -1 -> zero row
    xw = padded[src]                                                     # gather == pad+roll+partition
    mask = torch.from_numpy(shift_mask(Z, H, W)).to(x.device) if roll else None
    aw = window_attention(xw, mask, params, prefix + "attention.", heads)
    keep = src >= 0                                                      # reverse+unroll+crop (:269-293)
    y = torch.empty_like(x[0])
    y[src[keep]] = aw[keep]
    y = y.unsqueeze(0)
    x = x + s1 * layer_norm(y, params, prefix + "norm1.")                # :296 (post-norm)
    x = x + s2 * layer_norm(mlp(x, params, prefix + "linear."), params, prefix + "norm2.")   # :297
    return x


def earth_layer(x, Z, H, W, depth, params, prefix, heads):
    """EarthSpecificLayer.forward, models/layers.py:138-155: roll on odd blocks."""
    for i in range(depth):
        x = earth_block(x, Z, H, W, i % 2 == 1, params, f"{prefix}blocks.EarthSpecificBlock{i}.", heads)
    return x


def down_sample(x, Z, H, W, params, prefix="downsample."):
    """DownSample.forward, models/layers.py:497-524."""
    C = x.shape[-1]
    v = F.pad(x.reshape(Z, H, W, C), (0, 0, 0, 0, 0, H % 2))             # pad H to even (:506)
    H2, W2 = v.shape[1] // 2, W // 2
    v = v.reshape(Z, H2, 2, W2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(1, Z * H2 * W2, 4 * C)
    v = layer_norm(v, params, prefix + "norm.")
    return v @ params[prefix + "linear.weight"].t()


def up_sample(x, params, prefix="upsample.", Z=8, H2=91, W2=180, H=181):
    """UpSample.forward, models/layers.py:540-567 (sizes hard-coded in the reference)."""
    y = x @ params[prefix + "linear1.weight"].t()
    Co = y.shape[-1] // 4
    y = y.reshape(Z, H2, W2, 2, 2, Co).permute(0, 1, 3, 2, 4, 5).reshape(Z, 2 * H2, 2 * W2, Co)
    y = y[:, :H].reshape(1, Z * H * 2 * W2, Co)                          # crop (:555-556)
    y = layer_norm(y, params, prefix + "norm.")
    return y @ params[prefix + "linear2.weight"].t()


def patch_recover(x, Z, H, W, params, prefix="_output_layer."):
    """PatchRecovery_pretrain.forward, models/layers.py:582-621 (no de-normalisation)."""
    C = x.shape[-1]
    t = x[0].reshape(Z, H, W, C)
    up = t[1:].reshape(-1, C) @ params[prefix + "conv.weight"][:, :, 0].t() + params[prefix + "conv.bias"]
    # channel = v*32 + pz*16 + ph*4 + pw  (:593-596)
    up = up.reshape(Z - 1, H, W, 5, 2, 4, 4).permute(3, 0, 4, 1, 5, 2, 6).reshape(5, 2 * (Z - 1), 4 * H, 4 * W)
    output = up[:, :-1, :-3, :].unsqueeze(0).contiguous()
    sf = t[0].reshape(-1, C) @ params[prefix + "conv_surface.weight"][:, :, 0].t() + params[prefix + "conv_surface.bias"]
    sf = sf.reshape(H, W, 4, 4, 4).permute(2, 0, 3, 1, 4).reshape(4, 4 * H, 4 * W)
    output_surface = sf[:, :-3, :].unsqueeze(0).contiguous()
    return output, output_surface


DEPTHS = (2, 6, 6, 2)
HEADS = (6, 12, 12, 6)
DIMS = (192, 384, 384, 192)


def pangu_forward(params, input, input_surface, statistics, maps, const_h,
                  depths: Sequence[int] = DEPTHS, heads: Sequence[int] = HEADS):
    """PanguModel.forward, models/pangu_model.py:61-104, eval mode, fp32, batch 1."""
    with torch.no_grad():
        x = patch_embed(input, input_surface, statistics, maps, const_h, params)
        x = earth_layer(x, 8, 181, 360, depths[0], params, "layers.EarthSpecificLayer0.", heads[0])
        skip = x
        x = down_sample(x, 8, 181, 360, params)
        x = earth_layer(x, 8, 91, 180, depths[1], params, "layers.EarthSpecificLayer1.", heads[1])
        x = earth_layer(x, 8, 91, 180, depths[2], params, "layers.EarthSpecificLayer2.", heads[2])
        x = up_sample(x, params)
        x = earth_layer(x, 8, 181, 360, depths[3], params, "layers.EarthSpecificLayer3.", heads[3])
        x = torch.cat((skip, x), dim=-1)
        return patch_recover(x, 8, 181, 360, params)


# --------------------------------------------------------------------------------------
# Seeded synthetic weights / inputs -- shared by tests, bench and the golden maker
# --------------------------------------------------------------------------------------
def trunc_normal_(t: torch.Tensor, std: float, gen: torch.Generator) -> torch.Tensor:
    """Truncated normal on [-2, 2] (absolute), like timm/torch trunc_normal_(std=.02)."""
    lo = (1.0 + math.erf(-2.0 / std / math.sqrt(2.0))) / 2.0
    hi = (1.0 + math.erf(2.0 / std / math.sqrt(2.0))) / 2.0
    t.uniform_(2 * lo - 1, 2 * hi - 1, generator=gen)
    t.erfinv_().mul_(std * math.sqrt(2.0)).clamp_(-2.0, 2.0)
    return t


def param_shapes(depths=DEPTHS, heads=HEADS, dims=DIMS) -> "Dict[str, Tuple[int, ...]]":
    """The 223-key state_dict contract of the reference (keys_all.csv col torch_name), in the
    reference's registration order.  Shapes from models/layers.py ctors."""
    from collections import OrderedDict
    s = OrderedDict()
    d0 = dims[0]
    s["_input_layer.conv.weight"] = (d0, 192, 1)
    s["_input_layer.conv.bias"] = (d0,)
    s["_input_layer.conv_surface.weight"] = (d0, 112, 1)
    s["_input_layer.conv_surface.bias"] = (d0,)
    s["downsample.linear.weight"] = (2 * d0, 4 * d0)
    s["downsample.norm.weight"] = (4 * d0,)
    s["downsample.norm.bias"] = (4 * d0,)
    for li, (dep, hd, dim) in enumerate(zip(depths, heads, dims)):
        T = 124 if dim == 192 else 64
        for bi in range(dep):
            p = f"layers.EarthSpecificLayer{li}.blocks.EarthSpecificBlock{bi}."
            s[p + "norm1.weight"] = (dim,)
            s[p + "norm1.bias"] = (dim,)
            s[p + "norm2.weight"] = (dim,)
            s[p + "norm2.bias"] = (dim,)
            s[p + "linear.linear1.weight"] = (4 * dim, dim)
            s[p + "linear.linear1.bias"] = (4 * dim,)
            s[p + "linear.linear2.weight"] = (dim, 4 * dim)
            s[p + "linear.linear2.bias"] = (dim,)
            s[p + "attention.earth_specific_bias"] = (1, T, hd, 144, 144)
            s[p + "attention.linear1.weight"] = (3 * dim, dim)
            s[p + "attention.linear1.bias"] = (3 * dim,)
            s[p + "attention.linear2.weight"] = (dim, dim)
            s[p + "attention.linear2.bias"] = (dim,)
    s["upsample.linear1.weight"] = (4 * dims[-1], dims[-2])
    s["upsample.linear2.weight"] = (dims[-1], dims[-1])
    s["upsample.norm.weight"] = (dims[-1],)
    s["upsample.norm.bias"] = (dims[-1],)
    s["_output_layer.conv.weight"] = (160, dims[-2], 1)
    s["_output_layer.conv.bias"] = (160,)
    s["_output_layer.conv_surface.weight"] = (64, dims[-2], 1)
    s["_output_layer.conv_surface.bias"] = (64,)
    return s


def synth_params(seed: int = 0, only_prefix: str = "", depths=DEPTHS, heads=HEADS, dims=DIMS,
                 perturb: bool = True) -> "Dict[str, torch.Tensor]":
    """Deterministic synthetic weights with the reference's *distributions* (trunc-normal 0.02 for
    Linear weights and bias tables, uniform(+-1/sqrt(fan_in)) for the k=1 convs).  Every key has
    its own CPU generator seeded from (seed, crc32(key)), so any consumer
    (reference module via load_state_dict, this oracle, the CUDA model) gets bit-identical
    tensors and a single block can be drawn without drawing the whole model.
    With perturb=True the reference's degenerate inits (LayerNorm 1/0, Linear bias 0,
    models/pangu_model.py:52-59) are replaced by small random values so that parity tests
    exercise every parameter."""
    out = {}
    for idx, (k, shp) in enumerate(param_shapes(depths, heads, dims).items()):
        if not k.startswith(only_prefix):
            continue
        gen = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) & 0x7FFFFFFF)
        t = torch.empty(shp, dtype=torch.float32)
        if k.endswith("earth_specific_bias") or (k.endswith(".weight") and len(shp) == 2):
            trunc_normal_(t, 0.02, gen)
        elif len(shp) == 3:                       # Conv1d k=1: kaiming-uniform bound 1/sqrt(fan_in)
            b = 1.0 / math.sqrt(shp[1])
            t.uniform_(-b, b, generator=gen)
        elif "norm" in k and k.endswith(".weight"):
            t.fill_(1.0)
            if perturb:
                t.add_(torch.empty(shp).uniform_(-0.1, 0.1, generator=gen))
        elif k.endswith(".bias"):
            if "conv" in k:
                fan_in = 192 if "_input_layer.conv." in k else (112 if "_input_layer" in k else dims[-2])
                b = 1.0 / math.sqrt(fan_in)
                t.uniform_(-b, b, generator=gen)
            else:
                t.zero_()
                if perturb:
                    t.uniform_(-0.05, 0.05, generator=gen)
        else:
            raise KeyError(k)
        out[k] = t
    return out


def synth_inputs(seed: int = 1):
    """Synthetic ERA5-shaped inputs and constants, SURVEY §8(d)."""
    g = torch.Generator().manual_seed(seed)
    inp = torch.randn(1, 5, 13, 721, 1440, generator=g)
    inp_s = torch.randn(1, 4, 721, 1440, generator=g)
    surface_mean = torch.randn(4, generator=g)
    surface_std = torch.rand(4, generator=g) + 0.5
    upper_mean = torch.randn(13, 1, 1, 5, generator=g)
    upper_std = torch.rand(13, 1, 1, 5, generator=g) + 0.5
    maps = torch.randn(1, 3, 724, 1440, generator=g)
    const_h = torch.randn(1, 1, 1, 13, 721, 1440, generator=g)
    return inp, inp_s, (surface_mean, surface_std, upper_mean, upper_std), maps, const_h


def grads(fn, tensors, gouts):
    """torch.autograd of `fn(*leaves)` (one of the functions above, closed over the rest) -- the reference's
    loss.backward() (models/pangu_sample.py:226) for a loss whose output gradients are `gouts`.
    tensors: dict name -> tensor; -> dict name -> gradient."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in tensors.items()}
    outs = fn(leaves)
    if isinstance(outs, torch.Tensor):
        outs, gouts = (outs,), (gouts,)
    gs = torch.autograd.grad(outs, list(leaves.values()), grad_outputs=list(gouts), allow_unused=True)
    return {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(leaves, gs)}


# ---------------------------------------------------------------------------------------------
# Verification scores (era5_data/score.py), the step after the forward in models/pangu_sample.py:test()
# ---------------------------------------------------------------------------------------------
def latitude_weights(num_lat: int) -> torch.Tensor:
    """era5_data/score.py:99-106 (lat, latitude_weighting_factor_torch): fp32, the reference's constant 3.1416."""
    j = torch.arange(start=0, end=num_lat)
    lat = 90. - j * 180. / float(num_lat - 1)
    c = torch.cos(3.1416 / 180. * lat)
    return num_lat * c / torch.sum(c)


def weighted_rmse_channels(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor = None) -> torch.Tensor:
    """era5_data/score.py:126-161 weighted_rmse_torch_channels: [n,c,h,w] or [c,h,w] -> per-channel latitude-weighted RMSE;
    with a mask the weighted squared error is summed over the valid points and divided by their summed weight."""
    w = latitude_weights(pred.shape[-2]).reshape(-1, 1)
    se = (pred - target) ** 2.
    if mask is None:
        return torch.sqrt(torch.mean(w * se, dim=(-1, -2)))
    m = mask.reshape((1,) * (pred.dim() - 2) + tuple(mask.shape))
    return torch.sqrt((w * m * se).sum(dim=(-1, -2)) / (w * m).sum(dim=(-1, -2)))


def weighted_acc_channels(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """era5_data/score.py:181-201 weighted_acc_torch_channels (inputs are anomalies, models/pangu_sample.py:549-556)."""
    w = latitude_weights(pred.shape[-2]).reshape(-1, 1)
    return torch.sum(w * pred * target, dim=(-1, -2)) / torch.sqrt(
        torch.sum(w * pred * pred, dim=(-1, -2)) * torch.sum(w * target * target, dim=(-1, -2)))


# ---------------------------------------------------------------------------------------------
# Training losses (models/pangu_sample.py:163-218), the step after the forward in train(); pinned by
# tests/golden/reference_loss_goldens.npz = the reference's own train() loop run on a stand-in model
# (tests/golden/make_loss_golden.py)
# ---------------------------------------------------------------------------------------------
def wind_speed(output_surface, target_surface, output, target):
    """models/pangu_sample.py:74-93 get_wind_speed: sqrt(u^2 + v^2) of surface channels (1, 2) and upper channels (3, 4)."""
    ws = lambda t, a, b: torch.sqrt(t[:, a] ** 2 + t[:, b] ** 2)
    return ws(output_surface, 1, 2), ws(target_surface, 1, 2), ws(output, 3, 4), ws(target, 3, 4)


def training_loss(output, output_surface, target, target_surface, statistics_last, upper_weights, surface_weights,
                  upper_loss_weight=1.0, surface_loss_weight=0.25, only_use_wind_speed_loss=False, custom_mask=None):
    """The loss of one train() iteration, models/pangu_sample.py:168-204 (before the division by accumulation_steps):
    normData on the targets (era5_data/utils_data.py:531-537), then one of the four branches."""
    sm, ss, um, us = statistics_last
    target, target_surface = (target - um) / us, (target_surface - sm) / ss
    if custom_mask is not None:
        keep = ~(custom_mask == 0)                                           # :123, (~mask_bool)
        valid = custom_mask.sum()                                            # :127
    if only_use_wind_speed_loss:                                             # :184-193
        wos, wts, wo, wt = wind_speed(output_surface, target_surface, output, target)
        ls, lu = (wos - wts).abs(), (wo - wt).abs()
        if custom_mask is not None:
            return (ls * keep).sum() / valid + (lu * keep).sum() / valid
        return ls.mean() + lu.mean()
    ls, lu = (output_surface - target_surface).abs(), (output - target).abs()   # :194-195, L1Loss(reduction='none')
    if custom_mask is not None:                                              # :196-199
        wsl = (ls * surface_weights * keep[None, None]).sum() / valid
        wul = (lu * upper_weights * keep[None]).sum() / valid
    else:                                                                    # :200-202
        wsl, wul = torch.mean(ls * surface_weights), torch.mean(lu * upper_weights)
    return wul * upper_loss_weight + wsl * surface_loss_weight               # :204


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
