"""Op sequences of the reference modules (models/layers.py) on the B200 kernels.

Two numeric modes:
  * "fp32": CUDA-core FFMA kernels end to end -- the parity path (rel-L2 <= 1e-5 vs the reference);
  * "bf16": bf16 GEMM operands on tcgen05 tensor cores with fp32 accumulation; the residual stream,
    LayerNorm statistics and softmax stay fp32 (a bf16 shadow copy of the stream feeds the next GEMM).

All functions take plain [N, C] token tensors of ONE sample (the reference is only correct for
batch 1; batches are looped by the modules).
"""
import os

import torch

from . import ops
from .abi import ACT_GELU, ACT_NONE, ROLL_NONE, ROLL_SHIFT, WINDOWED, PanguError  # noqa: F401

MODES = ("fp32", "bf16")
# bf16 mode: Mlp + norm2 + residual as ONE tcgen05 kernel (0 = two GEMM kernels; kept for A/B measurements)
FUSED_MLP = os.environ.get("PANGU_B200_FUSED_MLP", "1") != "0"
# bf16 mode, C = 384: attention.linear2 + norm1 + shortcut fused IN FRONT of the Mlp kernel (x1 never reaches HBM); 0 = two kernels
FUSED_PROJ = os.environ.get("PANGU_B200_FUSED_PROJ", "1") != "0"


def default_mode():
    m = os.environ.get("PANGU_B200_COMPUTE", "bf16").lower()
    if m not in MODES:
        raise PanguError(f"PANGU_B200_COMPUTE={m!r}: expected one of {MODES}")
    return m


class WeightCache:
    """bf16 copies of parameters, refreshed when the parameter is updated in place or replaced."""

    def __init__(self):
        self._c = {}

    def f16(self, key, p):
        return self.bf16(key, p, torch.float16)

    @staticmethod
    def _transient(p):
        """Folded (LoRA) weights are fresh tensors every call: their data_ptr may be recycled, so they are never cached."""
        return not isinstance(p, torch.nn.Parameter)

    def bf16(self, key, p, dtype=torch.bfloat16):
        if self._transient(p):
            t = p.detach()
            return (t[:, :, 0] if t.dim() == 3 else t).contiguous().to(dtype)
        ent = self._c.get(key)
        tag = (p.data_ptr(), p._version, p.device)
        if ent is None or ent[0] != tag:
            t = p.detach()
            if t.dim() == 3:                       # Conv1d(k=1) weight [out, in, 1]
                t = t[:, :, 0]
            ent = (tag, t.contiguous().to(dtype))
            self._c[key] = ent
        return ent[1]

    def derived(self, key, params, fn):
        """Cache fn(*params) until one of the parameters is updated in place or replaced."""
        if any(self._transient(p) for p in params):
            with torch.no_grad():
                return fn(*[p.detach() for p in params])
        ent = self._c.get(key)
        tag = tuple((p.data_ptr(), p._version, p.device) for p in params)
        if ent is None or ent[0] != tag:
            with torch.no_grad():
                ent = (tag, fn(*[p.detach() for p in params]))
            self._c[key] = ent
        return ent[1]

    def clear(self):
        self._c.clear()


# ------------------------------------------------------------------------------------------
# Holders of dense weights.  The kernels read weights straight off the sub-modules (they never call the sub-modules'
# forward), so anything that changes what `linear(x)` computes must either be folded into the operands or refused:
# a silently ignored adapter / hook would be a wrong answer (reference: finetune/lora_tune.py:169-186 wraps every
# nn.Linear with peft LoRA and puts the two output convolutions into `modules_to_save`).
def _has_hooks(mod):
    return bool(getattr(mod, "_forward_hooks", None)) or bool(getattr(mod, "_forward_pre_hooks", None))


def _dropout_p(d):
    return float(getattr(d, "p", 0.0)) if isinstance(d, torch.nn.Dropout) else 0.0


def lin_wb(mod, who="linear"):
    """(weight, bias) of a Linear / Conv1d(k=1) holder as the kernels must see them.

    * exactly nn.Linear / nn.Conv1d without forward hooks -> its parameters;
    * a peft ModulesToSaveWrapper (`modules_to_save`, `original_module`) -> the active copy;
    * a peft LoRA layer (`base_layer`, `lora_A`, `lora_B`, `scaling`) -> W + sum_a scaling_a * B_a @ A_a as a differentiable
      expression of (W, A, B), so the backward kernels' weight gradient reaches lora_A / lora_B through autograd;
      with `lora_dropout` > 0 in train() mode the adapter input differs from the base input and cannot be folded: refused;
    * anything else (sub-classes, other adapters, hooked modules) -> PanguError.  No silent fallback."""
    if _has_hooks(mod):
        raise PanguError(f"{who}: forward hooks on {type(mod).__name__} would be bypassed by the fused kernels; remove them")
    t = type(mod)
    if t is torch.nn.Linear or t is torch.nn.Conv1d:
        return mod.weight, mod.bias
    if hasattr(mod, "modules_to_save") and hasattr(mod, "original_module"):          # peft ModulesToSaveWrapper
        act = [a for a in getattr(mod, "active_adapters", []) if a in mod.modules_to_save]
        if getattr(mod, "disable_adapters", False) or not act:
            return lin_wb(mod.original_module, who)
        return lin_wb(mod.modules_to_save[act[0]], who)
    if hasattr(mod, "base_layer") and hasattr(mod, "lora_A") and hasattr(mod, "lora_B"):   # peft lora.Linear
        w, b = lin_wb(mod.base_layer, who)
        if getattr(mod, "merged", False) or getattr(mod, "disable_adapters", False):
            return w, b
        dora = getattr(mod, "use_dora", None)
        for name in getattr(mod, "active_adapters", list(mod.lora_A.keys())):
            if name not in mod.lora_A:
                continue
            drop = mod.lora_dropout[name] if hasattr(mod, "lora_dropout") and name in mod.lora_dropout else None
            if mod.training and _dropout_p(drop) > 0.0:
                raise PanguError(f"{who}: LoRA adapter {name!r} has lora_dropout={_dropout_p(drop)} in train() mode; the adapter "
                                 "then sees a different input than the base layer and cannot be folded into the fused "
                                 "kernels' weights. Use lora_dropout=0 (or eval()) with the B200 path")
            if isinstance(dora, dict) and dora.get(name, False):
                raise PanguError(f"{who}: DoRA adapters are not supported by the B200 path")
            A, Bm = mod.lora_A[name].weight, mod.lora_B[name].weight
            w = w + float(mod.scaling[name]) * (Bm @ A).reshape(w.shape)
        return w, b
    raise PanguError(f"{who}: expected nn.Linear / nn.Conv1d (or a peft LoRA / modules_to_save wrapper of one), got "
                     f"{t.__module__}.{t.__name__}; the fused kernels would ignore what it adds to the forward")


def norm_wb(mod, who="norm"):
    if type(mod) is not torch.nn.LayerNorm or _has_hooks(mod):
        raise PanguError(f"{who}: expected a plain nn.LayerNorm without hooks, got {type(mod).__name__}")
    return mod.weight, mod.bias


def lin_w(mod):
    return lin_wb(mod)[0]


def lin_b(mod):
    return lin_wb(mod)[1]


LOG2E = 1.4426950408889634


def attention_operands(att, wc):
    """bf16 tensor-core operands of one EarthAttention3D with the softmax scaling folded in once per weight
    update: the q rows of linear1 (weight and bias) carry scale * log2(e) and the Earth-specific bias carries
    log2(e), so that the kernel's q k^T + bias is directly the exp2 exponent (models/layers.py:431-453).
    -> (w_qkv bf16 [3C, C], b_qkv fp32 [3C], bias table bf16 [T, heads, 144, 144])"""
    C = att.dim
    qs = att.scale * LOG2E

    def scaled_w(w):
        w = w.float().clone()
        w[:C] *= qs
        return w.to(torch.bfloat16).contiguous()

    def scaled_b(b):
        b = b.float().clone()
        b[:C] *= qs
        return b.contiguous()

    return (wc.derived("a1s", (lin_w(att.linear1),), scaled_w), wc.derived("b1s", (lin_b(att.linear1),), scaled_b),
            wc.derived("ebs", (att.earth_specific_bias,), lambda e: (e[0].float() * LOG2E).to(torch.bfloat16).contiguous()))


# The tcgen05 attention kernel shifts each score row by the BOUND max(S) + max(bias row) (no bias traffic in its max pass);
# the bound is loose by at most the spread of the bias row, and bf16 P / fp32 sums absorb ~100 log2 units of looseness.
# A bias table with wider rows (not seen at init or in the stress tests, but nothing forbids it) gets the exact maximum.
BIAS_SPREAD_LIMIT = 80.0          # log2 units


def attention_exact_max(att, wc):
    """True when some row of the (log2-scaled) Earth-specific bias spreads over more than BIAS_SPREAD_LIMIT: the caller then
    asks the kernel for the exact row maximum (PANGU_ATTN_EXACT_MAX).  One device reduction + host read per weight update,
    made while the operand caches are (re)built, i.e. outside CUDA-graph capture."""
    def spread(e):
        t = e[0].float()
        return float(((t.amax(-1) - t.amin(-1)).max() * LOG2E).item()) > BIAS_SPREAD_LIMIT
    return wc.derived("ebx", (att.earth_specific_bias,), spread)


def _w2d(p):
    """Linear [out,in] or Conv1d(k=1) [out,in,1] weight as a contiguous fp32 matrix."""
    t = p.detach()
    if t.dim() == 3:
        t = t[:, :, 0]
    return t.contiguous()


def _f(p):
    return None if p is None else p.detach().contiguous()


# ------------------------------------------------------------------------------------------
def _affine(norm, s):
    """LayerNorm affine parameters with a DropPath factor folded in: s * (LN(y) g + b) = LN(y) (s g) + (s b)."""
    g, b = (_f(t) for t in norm_wb(norm))
    return (g, b) if s == 1.0 else (g * s, b * s)


def block_forward(blk, x, Z, H, W, roll, mode, xb=None, s1=1.0, s2=1.0):
    """EarthSpecificBlock.forward (models/layers.py:218-299).  s1 / s2 are the DropPath factors of the attention and
    Mlp branches for this sample (1 in eval; in training 0 = branch dropped, else 1/keep -- timm DropPath with batch 1).
    x fp32 [N, C]; returns (x_out fp32, x_out_bf16 or None)."""
    att, mlp = blk.attention, blk.linear
    heads = att.head_number
    rmode = ROLL_SHIFT if roll else ROLL_NONE
    if (s1 != 1.0 or s2 != 1.0) and mode != "bf16":
        raise PanguError("stochastic depth (training) runs in compute_dtype='bf16' only")
    if mode == "fp32":
        qkv = ops.linear(x, _w2d(lin_w(att.linear1)), _f(lin_b(att.linear1)))
        o = ops.window_attention(qkv, _f(lin_b(att.linear1)), _f(att.earth_specific_bias), Z, H, W, heads, rmode)
        del qkv
        y = ops.linear(o, _w2d(lin_w(att.linear2)), _f(lin_b(att.linear2)))
        x1, _ = ops.ln_residual(y, _f(norm_wb(blk.norm1)[0]), _f(norm_wb(blk.norm1)[1]), residual=x, eps=blk.norm1.eps)
        h = ops.linear(x1, _w2d(lin_w(mlp.linear1)), _f(lin_b(mlp.linear1)), act=ACT_GELU)
        y = ops.linear(h, _w2d(lin_w(mlp.linear2)), _f(lin_b(mlp.linear2)))
        del h
        x2, _ = ops.ln_residual(y, _f(norm_wb(blk.norm2)[0]), _f(norm_wb(blk.norm2)[1]), residual=x1, eps=blk.norm2.eps)
        return x2, None
    wc = blk._wcache
    if xb is None:
        xb = ops.cast_bf16(x)
    if s1 != 0.0:
        w_qkv, b_qkv, eb = attention_operands(att, wc)
        qkv = ops.linear(xb, w_qkv, b_qkv)
        o, _ = ops.window_attention_band(qkv, None, b_qkv, eb, Z, H, W, heads, ops.full_band(H), 1 if roll else 0,
                                         prescaled=True, exact_max=attention_exact_max(att, wc))
        del qkv
        g1, b1 = _affine(blk.norm1, s1)
        if FUSED_MLP and FUSED_PROJ and s2 != 0.0 and x.shape[-1] == 384:
            g2, b2 = _affine(blk.norm2, s2)
            return ops.attn_proj_mlp_ln_bf16(o, wc.bf16("a2", lin_w(att.linear2)), _f(lin_b(att.linear2)), g1, b1, x,
                                             wc.bf16("m1", lin_w(mlp.linear1)), _f(lin_b(mlp.linear1)),
                                             wc.f16("m2h", lin_w(mlp.linear2)), _f(lin_b(mlp.linear2)), g2, b2,
                                             eps1=blk.norm1.eps, eps2=blk.norm2.eps)
        x1, x1b = ops.linear_ln_residual_bf16(o, wc.bf16("a2", lin_w(att.linear2)), _f(lin_b(att.linear2)), g1, b1, x,
                                              eps=blk.norm1.eps)
        del o
    else:
        x1, x1b = x, xb
    if s2 == 0.0:
        return x1, x1b
    g2, b2 = _affine(blk.norm2, s2)
    if FUSED_MLP:
        x2, x2b = ops.mlp_ln_residual_bf16(x1b, wc.bf16("m1", lin_w(mlp.linear1)), _f(lin_b(mlp.linear1)),
                                           wc.f16("m2h", lin_w(mlp.linear2)), _f(lin_b(mlp.linear2)), g2, b2, x1,
                                           eps=blk.norm2.eps)
        return x2, x2b
    h = ops.linear(x1b, wc.bf16("m1", lin_w(mlp.linear1)), _f(lin_b(mlp.linear1)), act=ACT_GELU)
    x2, x2b = ops.linear_ln_residual_bf16(h, wc.bf16("m2", lin_w(mlp.linear2)), _f(lin_b(mlp.linear2)), g2, b2, x1,
                                          eps=blk.norm2.eps)
    return x2, x2b


def attention_windows_forward(att, xw, mask, mode):
    """EarthAttention3D.forward (models/layers.py:413-484) on pre-partitioned windows
    xw [nLon, T, 144, C]; mask [nLon, T, 144, 144] (identical across nLon, as gen_mask builds it) or None."""
    nLon, T, L, C = xw.shape
    heads = att.head_number
    bias = att.earth_specific_bias.detach()[0]                       # [T, heads, 144, 144]
    if mask is not None:
        m0 = mask[0] if mask.dim() == 4 else mask
        if mask.dim() == 4 and mask.shape[0] > 1 and not bool((mask == mask[0:1]).all()):
            raise PanguError("EarthAttention3D: masks that differ between longitude windows are not supported")
        bias = bias + m0.to(bias.dtype).unsqueeze(1)
    bias = bias.contiguous()
    # geometry that reproduces (nLon, T) for the windowed identity map: Z=2*nZ.. any (Z,H,W) with the
    # same window counts works because mode WINDOWED ignores coordinates.
    Zf, Hf, Wf = 2, 6 * T - 5, 12 * nLon
    flat = xw.reshape(nLon * T * L, C)
    if mode == "fp32":
        qkv = ops.linear(flat.contiguous(), _w2d(lin_w(att.linear1)), _f(lin_b(att.linear1)))
        o = ops.window_attention(qkv, _f(lin_b(att.linear1)), bias, Zf, Hf, Wf, heads, WINDOWED)
        y = ops.linear(o, _w2d(lin_w(att.linear2)), _f(lin_b(att.linear2)))
    else:
        wc = att._wcache
        qkv = ops.linear(ops.cast_bf16(flat.contiguous()), wc.bf16("a1", lin_w(att.linear1)), _f(lin_b(att.linear1)))
        o = ops.window_attention(qkv, _f(lin_b(att.linear1)), bias.to(torch.bfloat16), Zf, Hf, Wf, heads, WINDOWED)
        y = ops.linear(o, wc.bf16("a2", lin_w(att.linear2)), _f(lin_b(att.linear2)), out_dtype=torch.float32)
    return y.reshape(nLon, T, L, C)


def mlp_forward(mlp, x2d, mode):
    """Mlp.forward (models/layers.py:311-317) on [M, C]."""
    if mode == "fp32":
        h = ops.linear(x2d, _w2d(lin_w(mlp.linear1)), _f(lin_b(mlp.linear1)), act=ACT_GELU)
        return ops.linear(h, _w2d(lin_w(mlp.linear2)), _f(lin_b(mlp.linear2)))
    wc = mlp._wcache
    h = ops.linear(ops.cast_bf16(x2d), wc.bf16("m1", lin_w(mlp.linear1)), _f(lin_b(mlp.linear1)), act=ACT_GELU)
    return ops.linear(h, wc.bf16("m2", lin_w(mlp.linear2)), _f(lin_b(mlp.linear2)), out_dtype=torch.float32)


def patch_embed_forward(pe, inp, inp_s, statistics, maps, const_h, mode):
    """PatchEmbedding_pretrain.forward (models/layers.py:53-120) for one sample -> ([N, dim] fp32, bf16|None)."""
    dim = pe.conv.out_channels
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    ps, pu = ops.patch_embed_gather(inp, inp_s, statistics, maps, const_h, dt)
    ns = ps.shape[0]                                     # token rows * 360 (181 * 360 for the full grid)
    x = torch.empty((8 * ns, dim), dtype=torch.float32, device=inp.device)
    if mode == "fp32":
        ops.linear(ps, _w2d(lin_w(pe.conv_surface)), _f(lin_b(pe.conv_surface)), out=x[:ns])
        ops.linear(pu, _w2d(lin_w(pe.conv)), _f(lin_b(pe.conv)), out=x[ns:])
        return x, None
    wc = pe._wcache
    xb = torch.empty((8 * ns, dim), dtype=torch.bfloat16, device=inp.device)      # bf16 shadow written by the GEMMs
    ops.linear_ex(ps, wc.bf16("cs", lin_w(pe.conv_surface)), _f(lin_b(pe.conv_surface)), out=x[:ns], shadow=xb[:ns])
    ops.linear_ex(pu, wc.bf16("c", lin_w(pe.conv)), _f(lin_b(pe.conv)), out=x[ns:], shadow=xb[ns:])
    return x, xb


def downsample_forward(ds, x, Z, H, W, mode):
    """DownSample.forward (models/layers.py:497-524) -> ([N/4.., 2C] fp32, bf16|None)."""
    if mode == "fp32":
        m = ops.downsample_merge_ln(x, _f(norm_wb(ds.norm)[0]), _f(norm_wb(ds.norm)[1]), Z, H, W, torch.float32, ds.norm.eps)
        return ops.linear(m, _w2d(lin_w(ds.linear)), None), None
    m = ops.downsample_merge_ln(x, _f(norm_wb(ds.norm)[0]), _f(norm_wb(ds.norm)[1]), Z, H, W, torch.bfloat16, ds.norm.eps)
    return ops.linear_ex(m, ds._wcache.bf16("l", lin_w(ds.linear)), None, want_shadow=True)


def upsample_forward(us, x, mode, xb=None, Z=8, H2=91, W2=180, H=181):
    """UpSample.forward (models/layers.py:540-567; sizes hard-coded there)."""
    if mode == "fp32":
        y = ops.linear(x, _w2d(lin_w(us.linear1)), None)
        n = ops.upsample_shuffle_ln(y, _f(norm_wb(us.norm)[0]), _f(norm_wb(us.norm)[1]), Z, H2, W2, H, torch.float32, us.norm.eps)
        return ops.linear(n, _w2d(lin_w(us.linear2)), None), None
    wc = us._wcache
    if xb is None:
        xb = ops.cast_bf16(x)
    y = ops.linear(xb, wc.bf16("l1", lin_w(us.linear1)), None)
    n = ops.upsample_shuffle_ln(y, _f(norm_wb(us.norm)[0]), _f(norm_wb(us.norm)[1]), Z, H2, W2, H, torch.bfloat16, us.norm.eps)
    return ops.linear_ex(n, wc.bf16("l2", lin_w(us.linear2)), None, want_shadow=True)


def patch_recover_forward(pr, x, Z, H, W, mode, skip=None, lat=721, denorm=None, xb=None, skip_b=None):
    """PatchRecovery_pretrain.forward (models/layers.py:582-621).  x [N, dim] fp32, or when `skip` is given
    the pair (skip, x) whose channel concat (models/pangu_model.py:98) is the input."""
    if (Z, W) != (8, 360) or H != (lat + 3) // 4:
        raise PanguError("PatchRecovery_pretrain is hard-wired to the (8,181,360) grid, like the reference "
                         "(or a latitude band of it)")
    ns = H * W
    if mode == "fp32":
        if skip is not None:
            x = torch.cat((skip, x), dim=-1)
        yu = ops.linear(x[ns:], _w2d(lin_w(pr.conv)), _f(lin_b(pr.conv)))
        ys = ops.linear(x[:ns], _w2d(lin_w(pr.conv_surface)), _f(lin_b(pr.conv_surface)))
        return ops.patch_recover_scatter(yu, ys, lat, denorm)
    wc = pr._wcache
    if skip is not None and xb is not None and skip_b is not None:
        # the skip concat (models/pangu_model.py:98) is read by the GEMMs from the two bf16 shadows directly
        yu = ops.linear_ex(skip_b[ns:], wc.bf16("c", lin_w(pr.conv)), _f(lin_b(pr.conv)), a2=xb[ns:])
        ys = ops.linear_ex(skip_b[:ns], wc.bf16("cs", lin_w(pr.conv_surface)), _f(lin_b(pr.conv_surface)), a2=xb[:ns])
        return ops.patch_recover_scatter(yu, ys, lat, denorm)
    xb = ops.concat_cast_bf16(skip, x) if skip is not None else ops.cast_bf16(x)
    yu = ops.linear(xb[ns:], wc.bf16("c", lin_w(pr.conv)), _f(lin_b(pr.conv)), out_dtype=torch.float32)
    ys = ops.linear(xb[:ns], wc.bf16("cs", lin_w(pr.conv_surface)), _f(lin_b(pr.conv_surface)), out_dtype=torch.float32)
    return ops.patch_recover_scatter(yu, ys, lat, denorm)
