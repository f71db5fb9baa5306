"""Fine-tune path: torch.autograd.Function wrappers whose backward runs on the B200 kernels.

The reference trains `PanguModel` with plain autograd (`loss.backward()`, models/pangu_sample.py:226) under DDP
(finetune/finetune_fully.py:220) and re-computes every block in the backward pass (checkpoint.checkpoint,
models/layers.py:143-149).  This module keeps that shape: one Function per EarthSpecificBlock / PatchEmbedding /
DownSample / UpSample / PatchRecovery.  With 180 GB of HBM3e per GPU the block Function SAVES its intermediates
(35 GB per sample) instead of re-computing them; $PANGU_B200_TRAIN_RECOMPUTE=1 gives the reference's behaviour
(only the block inputs are kept; the backward re-runs the un-fused forward first).  The backward runs

  * dgrad GEMMs  -- the forward tcgen05 GEMM kernels on transposed bf16 weight copies, the residual-gradient add
                    fused into their epilogue (`ops.linear_add`);
  * wgrad GEMMs  -- `ops.linear_wgrad`: MN-major tcgen05 split over token ranges, no transposed copies;
  * the window-attention backward kernel (`ops.window_attention_backward`), LayerNorm / GELU backward kernels.

Parameter gradients are fp32 tensors shaped like the parameters, so DDP's bucketed NCCL all-reduce (or
`pangu_b200.dist.allreduce_gradients`) sees exactly what it sees with the reference.  bf16 compute mode only.
"""
import os

import torch

from . import functional as PF
from . import ops
from .abi import PanguError

F32 = torch.float32
# 1 = keep only block inputs and re-compute the block in its backward (the reference's checkpointing)
RECOMPUTE = os.environ.get("PANGU_B200_TRAIN_RECOMPUTE", "0") != "0"


def _zeros(n, dev):
    return torch.zeros((n,), dtype=F32, device=dev)


def _wT(wc, key, p):
    """bf16 copy of a weight TRANSPOSED to [in, out] -- the 'W' operand of a dgrad GEMM (dX = dY @ W)."""
    def f(w):
        if w.dim() == 3:
            w = w[:, :, 0]
        return w.t().contiguous().to(torch.bfloat16)
    return wc.derived(key + ".T", (p,), f)


def _like(p, g2d):
    return g2d.reshape(p.shape)


# ------------------------------------------------------------------------------------------ EarthSpecificBlock
BLOCK_PARAMS = ("attention.linear1.weight", "attention.linear1.bias", "attention.earth_specific_bias",
                "attention.linear2.weight", "attention.linear2.bias", "norm1.weight", "norm1.bias",
                "linear.linear1.weight", "linear.linear1.bias", "linear.linear2.weight", "linear.linear2.bias",
                "norm2.weight", "norm2.bias")


def block_params(blk):
    a, m = blk.attention, blk.linear
    return (PF.lin_w(a.linear1), PF.lin_b(a.linear1), a.earth_specific_bias, PF.lin_w(a.linear2), PF.lin_b(a.linear2),
            blk.norm1.weight, blk.norm1.bias, PF.lin_w(m.linear1), PF.lin_b(m.linear1), PF.lin_w(m.linear2), PF.lin_b(m.linear2),
            blk.norm2.weight, blk.norm2.bias)


SAVED_KEYS = ("x0b", "qkv", "o", "lse", "y1", "x1b", "h_pre", "h", "y2")


def block_forward_train(blk, x0, x0b, Z, H, W, roll, s1, s2):
    """EarthSpecificBlock.forward (models/layers.py:218-299) with the UN-fused kernels, so that every intermediate the
    backward needs is left in HBM: qkv, the attention output and its log2-sum-exp rows, the two pre-LayerNorm
    tensors, the Mlp hidden activation before and after the GELU.  -> (x2 fp32, x2 bf16, saved dict)."""
    att, mlp = blk.attention, blk.linear
    wc = blk._wcache
    f = PF._f
    sv = {"x0b": x0b}
    if s1 != 0.0:
        w_qkv, b_qkv, eb = PF.attention_operands(att, wc)               # pre-scaled (scale*log2e folded into q)
        qkv = ops.linear(x0b, w_qkv, b_qkv)
        o, lse = ops.window_attention_train(qkv, b_qkv, eb, Z, H, W, att.head_number, 1 if roll else 0,
                                            exact_max=True)   # parameters change every step: no per-step host read of the bias spread
        y1 = ops.linear(o, wc.bf16("a2", PF.lin_w(att.linear2)), f(PF.lin_b(att.linear2)), out_dtype=F32)
        gam1, bet1 = PF._affine(blk.norm1, s1)
        x1, x1b = ops.ln_residual(y1, gam1, bet1, residual=x0, want_bf16=True, eps=blk.norm1.eps)
        sv.update(qkv=qkv, o=o, lse=lse, y1=y1)
    else:
        x1, x1b = x0, x0b
    sv["x1b"] = x1b
    if s2 != 0.0:
        w1, b1 = wc.bf16("m1", PF.lin_w(mlp.linear1)), f(PF.lin_b(mlp.linear1))
        h_pre, h = ops.linear_gelu_pre(x1b, w1, b1)                     # one GEMM pass leaves both (aux epilogue)
        y2 = ops.linear(h, wc.bf16("m2", PF.lin_w(mlp.linear2)), f(PF.lin_b(mlp.linear2)), out_dtype=F32)
        gam2, bet2 = PF._affine(blk.norm2, s2)
        x2, x2b = ops.ln_residual(y2, gam2, bet2, residual=x1, want_bf16=True, eps=blk.norm2.eps)
        sv.update(h_pre=h_pre, h=h, y2=y2)
    else:
        x2, x2b = x1, x1b
    return x2, x2b, sv


def block_backward(blk, sv, Z, H, W, roll, s1, s2, g2):
    """Backward of EarthSpecificBlock.forward for one sample.  sv: the intermediates of block_forward_train;
    s1, s2: DropPath factors of the two branches (0 = branch dropped); g2 fp32 [N, C]: gradient of the block output.
    -> (dx0 fp32, 13 parameter gradients in BLOCK_PARAMS order)."""
    att, mlp = blk.attention, blk.linear
    wc = blk._wcache
    x0b, x1b = sv["x0b"], sv["x1b"]
    dev = x0b.device
    N, C = x0b.shape
    heads = att.head_number
    f = PF._f
    g2 = g2.contiguous()
    ps = block_params(blk)
    grads = [None] * 13

    # ---- x2 = x1 + s2 * LN2(Mlp(x1))
    if s2 != 0.0:
        dg2, db2n, db2 = _zeros(C, dev), _zeros(C, dev), _zeros(C, dev)
        dy2 = ops.ln_backward(g2, sv["y2"], f(blk.norm2.weight), scale=s2, dgamma=dg2, dbeta=db2n, dcolsum=db2, eps=blk.norm2.eps)
        dw2 = ops.linear_wgrad(dy2, sv["h"])
        db1 = _zeros(4 * C, dev)
        dh = ops.linear_gelu_backward(dy2, _wT(wc, "m2", PF.lin_w(mlp.linear2)), sv["h_pre"], db1)   # [N, 4C] bf16, * gelu'(h_pre)
        del dy2
        dw1 = ops.linear_wgrad(dh, x1b)
        g1 = ops.linear_add(dh, _wT(wc, "m1", PF.lin_w(mlp.linear1)), None, addend=g2)
        del dh
        grads[7], grads[8], grads[9], grads[10], grads[11], grads[12] = dw1, db1, dw2, db2, dg2, db2n
    else:
        g1 = g2
        for i in range(7, 13):
            grads[i] = torch.zeros_like(ps[i])

    # ---- x1 = x0 + s1 * LN1(proj(attention(qkv(x0))))
    if s1 != 0.0:
        _w_qkv, b_qkv, eb = PF.attention_operands(att, wc)
        dg1, db1n, dba2 = _zeros(C, dev), _zeros(C, dev), _zeros(C, dev)
        dy1 = ops.ln_backward(g1, sv["y1"], f(blk.norm1.weight), scale=s1, dgamma=dg1, dbeta=db1n, dcolsum=dba2, eps=blk.norm1.eps)
        dwa2 = ops.linear_wgrad(dy1, sv["o"])
        do = ops.linear(dy1, _wT(wc, "a2", PF.lin_w(att.linear2)), None)   # [N, C] bf16
        del dy1
        d_eb = torch.zeros(att.earth_specific_bias.shape[1:], dtype=F32, device=dev)
        dba1 = _zeros(3 * C, dev)
        dqkv = ops.window_attention_backward(sv["qkv"], b_qkv, eb, sv["o"], do, sv["lse"], Z, H, W, heads, 1 if roll else 0,
                                             d_eb, dba1)
        del do
        dwa1 = ops.linear_wgrad(dqkv, x0b)
        g0 = ops.linear_add(dqkv, _wT(wc, "a1", PF.lin_w(att.linear1)), None, addend=g1)
        del dqkv
        grads[0], grads[1], grads[2], grads[3], grads[4], grads[5], grads[6] = \
            dwa1, dba1, d_eb.unsqueeze(0), dwa2, dba2, dg1, db1n
    else:
        g0 = g1
        for i in range(0, 7):
            grads[i] = torch.zeros_like(ps[i])
    return g0, [_like(p, g) for p, g in zip(ps, grads)]


class BlockFn(torch.autograd.Function):
    """(x fp32 [N,C], xb bf16 or None, 13 parameters) -> (x_out fp32, x_out bf16).

    Default: the forward runs the un-fused kernels and SAVES the intermediates (about 3.4 GB per stage-A block,
    1.7 GB per stage-B block, 35 GB per sample in total -- a B200 has 180 GB), so the backward re-computes nothing.
    $PANGU_B200_TRAIN_RECOMPUTE=1 restores the reference's checkpoint behaviour (models/layers.py:143-149): fused
    forward kernels, only the block input is kept, the backward re-runs the un-fused forward first."""

    @staticmethod
    def forward(ctx, x, xb, blk, Z, H, W, roll, s1, s2, *params):
        x = x.contiguous()
        if xb is None:
            xb = ops.cast_bf16(x)
        ctx.blk, ctx.geo, ctx.scales = blk, (Z, H, W, roll), (s1, s2)
        ctx.recompute = RECOMPUTE
        if RECOMPUTE:
            ctx.save_for_backward(x, xb)
            y, yb = PF.block_forward(blk, x, Z, H, W, roll, "bf16", xb, s1, s2)
        else:
            y, yb, sv = block_forward_train(blk, x, xb, Z, H, W, roll, s1, s2)
            ctx.keys = [k for k in SAVED_KEYS if k in sv]
            ctx.save_for_backward(*[sv[k] for k in ctx.keys])
        if y is x:                                              # both branches dropped: outputs must not alias inputs
            y, yb = x.clone(), xb.clone()
        ctx.mark_non_differentiable(yb)
        return y, yb

    @staticmethod
    def backward(ctx, g, _gb):
        Z, H, W, roll = ctx.geo
        s1, s2 = ctx.scales
        with torch.no_grad():
            if ctx.recompute:
                x, xb = ctx.saved_tensors
                _y, _yb, sv = block_forward_train(ctx.blk, x, xb, Z, H, W, roll, s1, s2)
            else:
                sv = dict(zip(ctx.keys, ctx.saved_tensors))
            dx, pg = block_backward(ctx.blk, sv, Z, H, W, roll, s1, s2, g)
        need = ctx.needs_input_grad
        return (dx if need[0] else None, None, None, None, None, None, None, None, None,
                *[gp if need[9 + i] else None for i, gp in enumerate(pg)])


# ------------------------------------------------------------------------------------------ PatchEmbedding
class PatchEmbedFn(torch.autograd.Function):
    """models/layers.py:53-120 for one sample; the input fields carry no gradient (they are data)."""

    @staticmethod
    def forward(ctx, inp, inp_s, pe, stats, maps, const_h, w, b, ws, bs):
        ctx.pe, ctx.aux = pe, (stats, maps, const_h)
        ctx.save_for_backward(inp, inp_s)
        x, xb = PF.patch_embed_forward(pe, inp, inp_s, stats, maps, const_h, "bf16")
        ctx.mark_non_differentiable(xb)
        return x, xb

    @staticmethod
    def backward(ctx, g, _gb):
        inp, inp_s = ctx.saved_tensors
        stats, maps, const_h = ctx.aux
        pe = ctx.pe
        with torch.no_grad():
            g = g.contiguous()
            ps, pu = ops.patch_embed_gather(inp, inp_s, stats, maps, const_h, torch.bfloat16)
            ns = ps.shape[0]
            gb = ops.cast_bf16(g)
            dws = ops.linear_wgrad(gb[:ns], ps)
            dw = ops.linear_wgrad(gb[ns:], pu)
            dbs, db = ops.colsum(g[:ns]), ops.colsum(g[ns:])
        return (None, None, None, None, None, None, _like(PF.lin_w(pe.conv), dw), db, _like(PF.lin_w(pe.conv_surface), dws), dbs)


# ------------------------------------------------------------------------------------------ DownSample
class DownSampleFn(torch.autograd.Function):
    """models/layers.py:497-524."""

    @staticmethod
    def forward(ctx, x, ds, Z, H, W, w, gamma, beta):
        x = x.contiguous()
        ctx.ds, ctx.geo = ds, (Z, H, W)
        ctx.save_for_backward(x)
        y, yb = PF.downsample_forward(ds, x, Z, H, W, "bf16")
        ctx.mark_non_differentiable(yb)
        return y, yb

    @staticmethod
    def backward(ctx, g, _gb):
        (x,) = ctx.saved_tensors
        ds = ctx.ds
        Z, H, W = ctx.geo
        with torch.no_grad():
            f = PF._f
            C4 = ds.norm.weight.shape[0]
            m = ops.downsample_merge_ln(x, f(ds.norm.weight), f(ds.norm.bias), Z, H, W, torch.bfloat16, ds.norm.eps)
            gb = ops.cast_bf16(g.contiguous())
            dw = ops.linear_wgrad(gb, m)
            del m
            dm = ops.linear_add(gb, _wT(ds._wcache, "l", PF.lin_w(ds.linear)), None)
            dgam, dbet = _zeros(C4, x.device), _zeros(C4, x.device)
            dx = ops.downsample_merge_ln_backward(dm, x, f(ds.norm.weight), dgam, dbet, Z, H, W, ds.norm.eps)
        return dx, None, None, None, None, dw, dgam, dbet


# ------------------------------------------------------------------------------------------ UpSample
class UpSampleFn(torch.autograd.Function):
    """models/layers.py:540-567."""

    @staticmethod
    def forward(ctx, x, xb, us, geo, w1, w2, gamma, beta):
        x = x.contiguous()
        if xb is None:
            xb = ops.cast_bf16(x)
        ctx.us, ctx.geo = us, geo
        ctx.save_for_backward(xb)
        Z, H2, W2, H = geo
        y, yb = PF.upsample_forward(us, x, "bf16", xb, Z, H2, W2, H)
        ctx.mark_non_differentiable(yb)
        return y, yb

    @staticmethod
    def backward(ctx, g, _gb):
        (xb,) = ctx.saved_tensors
        us = ctx.us
        Z, H2, W2, H = ctx.geo
        with torch.no_grad():
            f, wc = PF._f, us._wcache
            Co = us.norm.weight.shape[0]
            y = ops.linear(xb, wc.bf16("l1", PF.lin_w(us.linear1)), None)
            n = ops.upsample_shuffle_ln(y, f(us.norm.weight), f(us.norm.bias), Z, H2, W2, H, torch.bfloat16, us.norm.eps)
            gb = ops.cast_bf16(g.contiguous())
            dw2 = ops.linear_wgrad(gb, n)
            del n
            dn = ops.linear_add(gb, _wT(wc, "l2", PF.lin_w(us.linear2)), None)
            dgam, dbet = _zeros(Co, xb.device), _zeros(Co, xb.device)
            dy = ops.upsample_shuffle_ln_backward(dn, y, f(us.norm.weight), dgam, dbet, Z, H2, W2, H, us.norm.eps)
            del dn, y
            dw1 = ops.linear_wgrad(dy, xb)
            dx = ops.linear_add(dy, _wT(wc, "l1", PF.lin_w(us.linear1)), None)
        return dx, None, None, None, dw1, dw2, dgam, dbet


# ------------------------------------------------------------------------------------------ PatchRecovery
class PatchRecoverFn(torch.autograd.Function):
    """models/layers.py:582-621 on the channel concat cat(skip, x) (models/pangu_model.py:98) for one sample.
    `skip` may be None: then x already holds the concatenated channels."""

    @staticmethod
    def forward(ctx, skip, skip_b, x, xb, pr, geo, w, b, ws, bs):
        Z, H, W = geo
        x = x.contiguous()
        if xb is None:
            xb = ops.cast_bf16(x)
        if skip is not None and skip_b is None:
            skip_b = ops.cast_bf16(skip.contiguous())
        ctx.pr, ctx.geo, ctx.has_skip = pr, geo, skip is not None
        ctx.save_for_backward(skip_b, xb)
        if skip is not None:
            return PF.patch_recover_forward(pr, x, Z, H, W, "bf16", skip=skip, xb=xb, skip_b=skip_b)
        return PF.patch_recover_forward(pr, x, Z, H, W, "bf16")

    @staticmethod
    def backward(ctx, g_out, g_out_s):
        skip_b, xb = ctx.saved_tensors
        pr = ctx.pr
        Z, H, W = ctx.geo
        with torch.no_grad():
            dev = xb.device
            ns = H * W
            if g_out is None:
                g_out = torch.zeros((1, 5, 13, 721, 1440), dtype=F32, device=dev)
            if g_out_s is None:
                g_out_s = torch.zeros((1, 4, 721, 1440), dtype=F32, device=dev)
            dyu, dys = ops.patch_recover_gather_backward(g_out.contiguous().float(), g_out_s.contiguous().float(), 721)
            wc = pr._wcache
            wT, wsT = _wT(wc, "c", PF.lin_w(pr.conv)), _wT(wc, "cs", PF.lin_w(pr.conv_surface))       # [Cin, 160] / [Cin, 64]
            Cin = wT.shape[0]
            dw = torch.zeros((160, Cin), dtype=F32, device=dev)
            dws = torch.zeros((64, Cin), dtype=F32, device=dev)
            db, dbs = ops.colsum(dyu), ops.colsum(dys)
            if ctx.has_skip:
                Cs = skip_b.shape[1]
                parts = ((skip_b, 0, Cs), (xb, Cs, Cin))
            else:
                parts = ((xb, 0, Cin),)
            dins = []
            for src, c0, c1 in parts:
                ops.linear_wgrad(dyu, src[ns:], dw[:, c0:c1])
                ops.linear_wgrad(dys, src[:ns], dws[:, c0:c1])
                d = torch.empty((src.shape[0], c1 - c0), dtype=F32, device=dev)
                ops.linear_add(dyu, wT[c0:c1].contiguous(), None, out=d[ns:])
                ops.linear_add(dys, wsT[c0:c1].contiguous(), None, out=d[:ns])
                dins.append(d)
        dskip, dx = (dins[0], dins[1]) if ctx.has_skip else (None, dins[0])
        return (dskip, None, dx, None, None, None, _like(PF.lin_w(pr.conv), dw), db, _like(PF.lin_w(pr.conv_surface), dws), dbs)


# ------------------------------------------------------------------------------------------ module-level helpers
def wants_graph(mod, *tensors):
    """True when a forward call must build an autograd graph: grad mode is on and an input or any parameter of the
    module requires grad -- nn.Module / autograd semantics, in train() AND eval() mode (fine-tuning with DropPath
    switched off by eval(), test-time adaptation).  The graph path saves its intermediates (about 35 GB per sample at
    full resolution), so forward-only callers should do what they do with any PyTorch model: wrap the call
    in torch.no_grad() / inference_mode(), freeze the parameters, or set `module.forward_only = True` (checked on the
    module that is called; `PanguModel.set_forward_only()` sets it on the whole tree), which is how the reference's
    grad-enabled eval loop (models/pangu_sample.py:443) keeps the fused inference kernels."""
    if not torch.is_grad_enabled() or getattr(mod, "forward_only", False):
        return False
    if any(t is not None and t.requires_grad for t in tensors):
        return True
    return any(p.requires_grad for p in mod.parameters())


def require_bf16(mod):
    if mod._mode() != "bf16":
        raise PanguError(f"{type(mod).__name__}: the backward kernels exist for compute_dtype='bf16' only "
                         "(fp32 is the inference parity path)")


def block_apply(blk, x, xb, Z, H, W, roll):
    s1, s2 = blk.branch_scales()
    return BlockFn.apply(x, xb, blk, Z, H, W, roll, s1, s2, *block_params(blk))


def embed_apply(pe, inp, inp_s, stats, maps, const_h):
    return PatchEmbedFn.apply(inp, inp_s, pe, stats, maps, const_h, PF.lin_w(pe.conv), PF.lin_b(pe.conv),
                              PF.lin_w(pe.conv_surface), PF.lin_b(pe.conv_surface))


def downsample_apply(ds, x, Z, H, W):
    return DownSampleFn.apply(x, ds, Z, H, W, PF.lin_w(ds.linear), ds.norm.weight, ds.norm.bias)


def upsample_apply(us, x, xb, Z=8, H2=91, W2=180, H=181):
    return UpSampleFn.apply(x, xb, us, (Z, H2, W2, H), PF.lin_w(us.linear1), PF.lin_w(us.linear2), us.norm.weight, us.norm.bias)


def recover_apply(pr, x, xb, Z, H, W, skip=None, skip_b=None):
    return PatchRecoverFn.apply(skip, skip_b, x, xb, pr, (Z, H, W), PF.lin_w(pr.conv), PF.lin_b(pr.conv),
                                PF.lin_w(pr.conv_surface), PF.lin_b(pr.conv_surface))
