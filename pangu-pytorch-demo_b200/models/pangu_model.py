"""B200-native PanguModel with the API of the reference's models/pangu_model.py:18-104."""
import os
import sys
from collections import OrderedDict

import torch
from torch import nn

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from models.layers import *  # noqa: E402,F401,F403
from models.layers import (DownSample, EarthSpecificLayer, PatchEmbedding_pretrain, PatchRecovery_pretrain,  # noqa: E402
                           UpSample, _B200Module, _need_cuda, set_compute_dtype, trunc_normal_)
from pangu_b200 import autograd as AG  # noqa: E402
from pangu_b200 import functional as PF  # noqa: E402


class PanguModel(_B200Module):
    """Same constructor, forward signature, sub-module names and state_dict keys (keys_all.csv) as the
    reference.  forward(input, input_surface, statistics, maps, const_h) -> (output, output_surface),
    normalised units, fp32, on the input's CUDA device."""

    def __init__(self, depths=[2, 6, 6, 2], num_heads=[6, 12, 12, 6], dims=[192, 384, 384, 192],
                 patch_size=(2, 4, 4), device=None):
        super(PanguModel, self).__init__()
        self.device = device
        self._input_layer = PatchEmbedding_pretrain(patch_size, dims[0])
        self.downsample = DownSample(dims[0])
        dpr = [x.item() for x in torch.linspace(0, 0.2, sum(depths))]
        self.num_layers = len(depths)
        layer_list = OrderedDict()
        for i_layer in range(self.num_layers):
            layer_list['EarthSpecificLayer{}'.format(i_layer)] = EarthSpecificLayer(
                depth=depths[i_layer],
                dim=dims[i_layer],
                drop_path_ratio_list=dpr[sum(depths[:i_layer]):sum(depths[:i_layer + 1])],
                heads=num_heads[i_layer],
                use_checkpoint=self.training,
                device=self.device)
        self.layers = nn.Sequential(layer_list)
        self.upsample = UpSample(dims[-2], dims[-1])
        self._output_layer = PatchRecovery_pretrain(dims[-2])
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def set_compute_dtype(self, mode):
        return set_compute_dtype(self, mode)

    def forward_sample(self, inp, inp_s, stats, maps, const_h, denorm=None, graph=False):
        """One sample through models/pangu_model.py:61-104; the bf16 shadow of the residual stream is
        handed from kernel to kernel so that no separate cast pass is needed.  graph=True chains the autograd
        Functions of pangu_b200/autograd.py instead (fine-tuning; the skip connection's two gradient streams are
        summed by autograd)."""
        mode = self._mode()
        if graph:
            x, xb = AG.embed_apply(self._input_layer, inp, inp_s, stats, maps, const_h)
        else:
            x, xb = PF.patch_embed_forward(self._input_layer, inp, inp_s, stats, maps, const_h, mode)
        x, xb = self.layers[0].forward_sample(x, 8, 181, 360, xb, graph)
        skip, skip_b = x, xb
        x, xb = self.downsample.forward_sample(x, 8, 181, 360, graph)
        x, xb = self.layers[1].forward_sample(x, 8, 91, 180, xb, graph)
        x, xb = self.layers[2].forward_sample(x, 8, 91, 180, xb, graph)
        x, xb = self.upsample.forward_sample(x, xb, graph)
        x, xb = self.layers[3].forward_sample(x, 8, 181, 360, xb, graph)
        return self._output_layer.forward_sample(x, 8, 181, 360, skip=skip, denorm=denorm, xb=xb, skip_b=skip_b, graph=graph)

    def forward(self, input, input_surface, statistics, maps, const_h):
        inp, inp_s = _need_cuda(input, "PanguModel"), _need_cuda(input_surface, "PanguModel")
        stats = tuple(s.to(inp.device) for s in statistics)
        maps_c = maps.to(inp.device).float().contiguous()
        ch = const_h.to(inp.device).float().contiguous()
        graph = AG.wants_graph(self)
        if graph:
            AG.require_bf16(self)
        outs = [self.forward_sample(inp[b], inp_s[b], stats, maps_c, ch, graph=graph) for b in range(inp.shape[0])]
        if len(outs) == 1:
            return outs[0]
        return torch.cat([o[0] for o in outs], 0), torch.cat([o[1] for o in outs], 0)
