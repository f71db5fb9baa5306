"""B200-native modules with the API of the reference's models/layers.py.

Every class keeps the reference's constructor signature, forward signature, sub-module names and
parameter shapes, so `state_dict`s, DDP, deepcopy and pickling behave as with the reference (the
Linear sub-modules are real nn.Linear, so LoRA wrapping finds them; see "LoRA" below for what then happens).  The math runs in hand-written CUDA
kernels for sm_100a through the C ABI in include/pangu_b200.h (see pangu_b200/functional.py); there is
no eager-PyTorch or CPU fallback: tensors must live on a CUDA device.

Training: when grad mode is on and a parameter (or the input) requires grad -- in train() or eval() mode, as
with any nn.Module -- the forward builds an autograd graph out of the Functions in pangu_b200/autograd.py, whose
backward runs on the B200 backward kernels (bf16 mode; intermediates are saved, about 35 GB per full-resolution sample;
$PANGU_B200_TRAIN_RECOMPUTE=1 re-computes per block like models/layers.py:143-149).  Forward-only use: torch.no_grad()
or `module.set_forward_only()`.

LoRA: peft-wrapped Linear sub-modules are folded (W + scaling * B @ A) into the kernels' operands, with gradients for
lora_A / lora_B, when the adapter has no active dropout; otherwise, and for any other wrapper or hooked sub-module,
the forward raises PanguError (pangu_b200.functional.lin_wb) -- adapters are never silently ignored.

Numeric mode: `module.compute_dtype` in {"bf16", "fp32"} (default from $PANGU_B200_COMPUTE, "bf16");
`set_compute_dtype(module, mode)` switches a whole tree.
"""
import os
import sys
from collections import OrderedDict

import torch
from torch import nn

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from pangu_b200 import autograd as AG  # noqa: E402
from pangu_b200 import functional as PF  # noqa: E402
from pangu_b200.abi import PanguError  # noqa: E402

__all__ = ["PatchEmbedding_pretrain", "PatchEmbedding", "EarthSpecificLayer", "EarthSpecificBlock", "Mlp",
           "EarthAttention3D", "DownSample", "UpSample", "PatchRecovery_pretrain", "PatchRecovery", "DropPath",
           "trunc_normal_", "set_compute_dtype"]


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    """timm.models.layers.trunc_normal_ (used at models/layers.py:366, models/pangu_model.py:54)."""
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class DropPath(nn.Module):
    """Stochastic depth per sample (timm.models.layers.DropPath, used at models/layers.py:171)."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def branch_scale(self):
        """Scale applied to the residual branch of ONE sample: 1 in eval, else 0 or 1/keep."""
        if self.drop_prob == 0.0 or not self.training:
            return 1.0
        keep = 1.0 - self.drop_prob
        if float(torch.rand(())) >= keep:
            return 0.0
        return 1.0 / keep if (keep > 0.0 and self.scale_by_keep) else 1.0

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            m.div_(keep)
        return x * m


class _B200Module(nn.Module):
    """Common plumbing: numeric mode and the bf16 weight cache (never pickled / deep-copied)."""

    def __init__(self):
        super().__init__()
        self.compute_dtype = PF.default_mode()
        self._wcache = PF.WeightCache()

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_wcache"] = PF.WeightCache()
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = PF.WeightCache() if k == "_wcache" else copy.deepcopy(v, memo)
        return new

    def _mode(self):
        return self.compute_dtype

    def set_forward_only(self, flag=True):
        """Opt this module tree out of autograd-graph building even when grad mode is on and parameters are trainable
        (pangu_b200.autograd.wants_graph): the call then runs the fused inference kernels and returns tensors without
        grad_fn -- for grad-enabled evaluation loops like the reference's test() (models/pangu_sample.py:443)."""
        for m in self.modules():
            if isinstance(m, _B200Module):
                m.forward_only = bool(flag)
        return self


def set_compute_dtype(module, mode):
    if mode not in PF.MODES:
        raise PanguError(f"compute dtype {mode!r}: expected one of {PF.MODES}")
    for m in module.modules():
        if isinstance(m, _B200Module):
            m.compute_dtype = mode
    return module


def _need_cuda(x, who):
    if not x.is_cuda:
        raise PanguError(f"{who}: input is on {x.device}; the B200 path runs on CUDA only (no CPU fallback)")
    return x.contiguous().float()


def _no_training_graph(mod, *tensors):
    """Mlp / EarthAttention3D called on their own: forward only (inside a block they are differentiated by BlockFn)."""
    if AG.wants_graph(mod, *tensors):
        raise PanguError(f"{type(mod).__name__}: called stand-alone this module is forward-only; differentiate it through "
                         "EarthSpecificBlock / PanguModel, or call it under torch.no_grad()")


class PatchEmbedding_pretrain(_B200Module):
    """models/layers.py:18-120.  Conv1d(k=1) over pre-patchified channels (192 upper-air, 112 surface)."""

    def __init__(self, patch_size, dim):
        super().__init__()
        self.conv = nn.Conv1d(in_channels=192, out_channels=dim, kernel_size=1, stride=1)
        self.conv_surface = nn.Conv1d(in_channels=112, out_channels=dim, kernel_size=1, stride=1)
        self.window_size = (2, 6, 12)

    def forward(self, input, input_surface, statistics, maps, const_h):
        inp, inp_s = _need_cuda(input, "PatchEmbedding"), _need_cuda(input_surface, "PatchEmbedding")
        if tuple(inp.shape[1:]) != (5, 13, 721, 1440) or tuple(inp_s.shape[1:]) != (4, 721, 1440):
            raise PanguError("PatchEmbedding_pretrain is hard-wired to 13x721x1440 inputs, like the reference "
                             "(models/layers.py:90,114)")
        stats = tuple(s.to(inp.device) for s in statistics)
        maps_c = maps.to(inp.device).float().contiguous()
        ch = const_h.to(inp.device).float().contiguous()
        if AG.wants_graph(self):
            AG.require_bf16(self)
            outs = [AG.embed_apply(self, inp[b], inp_s[b], stats, maps_c, ch)[0] for b in range(inp.shape[0])]
        else:
            outs = [PF.patch_embed_forward(self, inp[b], inp_s[b], stats, maps_c, ch, self._mode())[0]
                    for b in range(inp.shape[0])]
        return torch.stack(outs, 0)


PatchEmbedding = PatchEmbedding_pretrain      # name used by north_star / models/pangu_model.py:25


class Mlp(_B200Module):
    """models/layers.py:302-317."""

    def __init__(self, dim, dropout_rate):
        super().__init__()
        self.linear1 = nn.Linear(dim, dim * 4)
        self.linear2 = nn.Linear(dim * 4, dim)
        self.activation = nn.GELU()
        self.drop = nn.Dropout(dropout_rate)

    def forward(self, x):
        _no_training_graph(self, x)
        shp = x.shape
        y = PF.mlp_forward(self, _need_cuda(x, "Mlp").reshape(-1, shp[-1]), self._mode())
        return y.reshape(shp)


class EarthAttention3D(_B200Module):
    """models/layers.py:320-484.  The Earth-specific bias is stored already expanded,
    [1, type_of_windows, heads, 144, 144], exactly as the reference stores it (:357-363)."""

    def __init__(self, dim, heads, dropout_rate, window_size, device):
        super().__init__()
        self.device = device
        self.linear1 = nn.Linear(dim, dim * 3, bias=True)
        self.linear2 = nn.Linear(dim, dim)
        self.softmax = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout_rate)
        self.head_number = heads
        self.dim = dim
        self.scale = (dim // heads) ** -0.5
        self.window_size = window_size
        if self.dim == 192:
            input_shape = [8, 186]
        elif self.dim == 384:
            input_shape = [8, 96]
        else:
            raise PanguError("EarthAttention3D: dim must be 192 or 384 (models/layers.py:347-350)")
        if dim // heads != 32 or tuple(window_size) != (2, 6, 12):
            raise PanguError("EarthAttention3D: the kernels are specialised for head_dim 32 and (2,6,12) windows")
        self.type_of_windows = (input_shape[0] // window_size[0]) * (input_shape[1] // window_size[1])
        n = window_size[0] * window_size[1] * window_size[2]
        bias = torch.zeros(1, self.type_of_windows, heads, n, n, device=self.device)
        self.earth_specific_bias = nn.Parameter(bias)
        trunc_normal_(self.earth_specific_bias, std=0.02)
        self._construct_index()

    def _construct_index(self):
        """models/layers.py:371-411 in closed form; unused by forward, kept as an attribute."""
        wz, wh, ww = self.window_size
        k = torch.arange(wz * wh * ww)
        z, h, w = k // (wh * ww), (k // ww) % wh, k % ww
        idx = (z[:, None] + wz * z[None, :]) * ((2 * ww - 1) * wh * wh) + \
              (h[:, None] + wh * h[None, :]) * (2 * ww - 1) + (w[:, None] - w[None, :] + ww - 1)
        self.position_index = idx.flatten().to(self.device) if self.device is not None else idx.flatten()

    def forward(self, x, mask):
        _no_training_graph(self, x)
        xw = _need_cuda(x, "EarthAttention3D")
        return PF.attention_windows_forward(self, xw, mask, self._mode())


class EarthSpecificBlock(_B200Module):
    """models/layers.py:158-299."""

    def __init__(self, dim, drop_path_ratio, heads, device):
        super().__init__()
        self.device = device
        self.window_size = (2, 6, 12)
        self.drop_path = DropPath(drop_path_ratio) if drop_path_ratio > 0. else nn.Identity()
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.linear = Mlp(dim, 0)
        self.attention = EarthAttention3D(dim, heads, 0, self.window_size, device=self.device)
        self.padding_front, self.padding_back = 0, 5
        input_shape = [8, 186] if dim == 192 else [8, 96]
        self.type_of_windows = (input_shape[0] // self.window_size[0]) * (input_shape[1] // self.window_size[1])

    def gen_mask(self, x):
        """models/layers.py:187-216: additive shift mask for a rolled padded tensor x [1, Z, Hp, W, C]."""
        from pangu_b200 import ops
        Z, Hp, W = x.shape[1], x.shape[2], x.shape[3]
        m = ops.shift_mask(Z, Hp - 5, W, x.device)
        return m.unsqueeze(0).expand(W // 12, -1, -1, -1)

    def _check_grid(self, Z, H, W):
        if (Z // 2) * ((H + 5) // 6) != self.type_of_windows or Z % 2 or (H + 5) % 6 or W % 12:
            raise PanguError(f"EarthSpecificBlock(dim={self.attention.dim}): grid ({Z},{H},{W}) does not give "
                             f"{self.type_of_windows} window types")

    def branch_scales(self):
        """DropPath factors of the attention and the Mlp branch for ONE sample (two independent draws, as the two
        self.drop_path(...) calls of models/layers.py:296-297 make): 1 in eval, else 0 or 1/keep."""
        if isinstance(self.drop_path, DropPath):
            return self.drop_path.branch_scale(), self.drop_path.branch_scale()
        return 1.0, 1.0

    def forward_sample(self, x, Z, H, W, roll, xb=None, graph=False):
        if graph:
            return AG.block_apply(self, x, xb, Z, H, W, roll)
        s1, s2 = self.branch_scales()
        return PF.block_forward(self, x, Z, H, W, roll, self._mode(), xb, s1, s2)

    def forward(self, x, Z, H, W, roll):
        self._check_grid(Z, H, W)
        xs = _need_cuda(x, "EarthSpecificBlock")
        graph = AG.wants_graph(self, x)
        if graph:
            AG.require_bf16(self)
        return torch.stack([self.forward_sample(xs[b], Z, H, W, roll, graph=graph)[0] for b in range(xs.shape[0])], 0)


class EarthSpecificLayer(_B200Module):
    """models/layers.py:123-155: `depth` blocks, the odd ones shifted."""

    def __init__(self, depth, dim, drop_path_ratio_list, heads, use_checkpoint, device):
        super().__init__()
        self.device = device
        self.depth = depth
        block_list = OrderedDict()
        for i_layer in range(depth):
            block_list['EarthSpecificBlock{}'.format(i_layer)] = EarthSpecificBlock(
                dim, drop_path_ratio_list[i_layer], heads, device=self.device)
        self.blocks = nn.Sequential(block_list)
        self.use_checkpoint = use_checkpoint     # API parity; BlockFn always re-computes the block in its backward

    def forward_sample(self, x, Z, H, W, xb=None, graph=False):
        for i, blk in enumerate(self.blocks):
            x, xb = blk.forward_sample(x, Z, H, W, i % 2 == 1, xb, graph)
        return x, xb

    def forward(self, x, Z, H, W):
        xs = _need_cuda(x, "EarthSpecificLayer")
        for blk in self.blocks:
            blk._check_grid(Z, H, W)
        graph = AG.wants_graph(self, x)
        if graph:
            AG.require_bf16(self)
        return torch.stack([self.forward_sample(xs[b], Z, H, W, graph=graph)[0] for b in range(xs.shape[0])], 0)


class DownSample(_B200Module):
    """models/layers.py:487-524."""

    def __init__(self, dim):
        super().__init__()
        self.linear = nn.Linear(in_features=4 * dim, out_features=2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)

    def forward_sample(self, x, Z, H, W, graph=False):
        if graph:
            return AG.downsample_apply(self, x, Z, H, W)
        return PF.downsample_forward(self, x, Z, H, W, self._mode())

    def forward(self, x, Z, H, W):
        xs = _need_cuda(x, "DownSample")
        graph = AG.wants_graph(self, x)
        if graph:
            AG.require_bf16(self)
        return torch.stack([self.forward_sample(xs[b], Z, H, W, graph)[0] for b in range(xs.shape[0])], 0)


class UpSample(_B200Module):
    """models/layers.py:527-567 (grid sizes 8 x 91 x 180 -> 8 x 181 x 360 hard-coded there)."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.linear1 = nn.Linear(input_dim, output_dim * 4, bias=False)
        self.linear2 = nn.Linear(output_dim, output_dim, bias=False)
        self.norm = nn.LayerNorm(output_dim)

    def forward_sample(self, x, xb=None, graph=False):
        if graph:
            return AG.upsample_apply(self, x, xb)
        return PF.upsample_forward(self, x, self._mode(), xb)

    def forward(self, x):
        xs = _need_cuda(x, "UpSample")
        if xs.shape[1] != 8 * 91 * 180:
            raise PanguError("UpSample is hard-wired to 8x91x180 tokens, like the reference (models/layers.py:546)")
        graph = AG.wants_graph(self, x)
        if graph:
            AG.require_bf16(self)
        return torch.stack([self.forward_sample(xs[b], graph=graph)[0] for b in range(xs.shape[0])], 0)


class PatchRecovery_pretrain(_B200Module):
    """models/layers.py:570-621."""

    def __init__(self, dim):
        super().__init__()
        self.patch_size = (2, 4, 4)
        self.dim = dim
        self.conv = nn.Conv1d(in_channels=dim, out_channels=160, kernel_size=1, stride=1)
        self.conv_surface = nn.Conv1d(in_channels=dim, out_channels=64, kernel_size=1, stride=1)

    def forward_sample(self, x, Z, H, W, skip=None, denorm=None, xb=None, skip_b=None, graph=False):
        if graph:
            if denorm is not None:
                raise PanguError("PatchRecovery: the fused de-normalisation is an inference feature")
            if (Z, H, W) != (8, 181, 360):
                raise PanguError("PatchRecovery_pretrain is hard-wired to the (8,181,360) grid, like the reference")
            return AG.recover_apply(self, x, xb, Z, H, W, skip, skip_b)
        return PF.patch_recover_forward(self, x, Z, H, W, self._mode(), skip, denorm=denorm, xb=xb, skip_b=skip_b)

    def forward(self, x, Z, H, W):
        xs = _need_cuda(x, "PatchRecovery")
        graph = AG.wants_graph(self, x)
        if graph:
            AG.require_bf16(self)
        outs = [self.forward_sample(xs[b], Z, H, W, graph=graph) for b in range(xs.shape[0])]
        return torch.cat([o[0] for o in outs], 0), torch.cat([o[1] for o in outs], 0)


PatchRecovery = PatchRecovery_pretrain
