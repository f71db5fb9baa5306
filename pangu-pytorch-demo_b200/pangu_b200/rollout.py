"""Iterative rollout: the 24 h model chained k times on the device (BASELINE configs[2]).

The reference chains forecasts by feeding the de-normalised output of one step back as the next input
(`inference/inference_multiOutput.py:171-187` for the ONNX model, `inference/inference_mix_multiOutput.py:201-238`
for the torch model: `normBackData(output, output_surface, weather_statistics_last)`, era5_data/utils_data.py:540-546)
through host numpy arrays.  Here the state never leaves the GPU: the de-normalisation is folded into the
patch-recover scatter kernel (`pangu_patch_recover_scatter_denorm`) and, with `graph=True`, every step is one
CUDA-graph replay whose static output buffers are copied into its static input buffers.

    ro = Rollout(model, statistics, statistics_last, maps, const_h)
    for upper, surface in ro.run(input, input_surface, steps=7):      # physical units, device tensors
        ...
"""
import torch

from .abi import PanguError
from .graph import GraphedForward


def _flat_stats_last(statistics_last, device):
    """weatherStatistics_output layout: surface [1,4,1,1] x2, upper [1,5,13,1,1] x2 -> flat fp32 device vectors."""
    sm, ss, um, us = statistics_last
    out = tuple(t.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous() for t in (sm, ss, um, us))
    if [t.numel() for t in out] != [4, 4, 65, 65]:
        raise PanguError("statistics_last must be (surface_mean [4], surface_std [4], upper_mean [5*13], upper_std [5*13])")
    return out


class Rollout:
    def __init__(self, model, statistics, statistics_last, maps, const_h, graph=True):
        self.model = model
        p = next(model.parameters())
        if not p.is_cuda:
            raise PanguError("Rollout: the model must be on a CUDA device")
        self.dev = p.device
        self.stats = tuple(s.to(self.dev) for s in statistics)
        self.denorm = _flat_stats_last(statistics_last, self.dev)
        self.maps = maps.to(self.dev).float().contiguous()
        self.const_h = const_h.to(self.dev).float().contiguous()
        self.use_graph = graph
        self._graphed = None

    def step(self, inp, inp_s):
        """One forecast step in physical units: [5,13,721,1440], [4,721,1440] -> [1,5,13,721,1440], [1,4,721,1440]."""
        with torch.no_grad():
            return self.model.forward_sample(inp, inp_s, self.stats, self.maps, self.const_h, denorm=self.denorm)

    def run(self, input, input_surface, steps=7):
        """Yields (upper, surface) after each of `steps` chained forecasts.  With graph replay the yielded tensors are
        the static output buffers: consume (or clone) them before advancing the generator."""
        inp = input.to(self.dev).float().reshape(5, 13, 721, 1440).contiguous()
        inp_s = input_surface.to(self.dev).float().reshape(4, 721, 1440).contiguous()
        if self.use_graph:
            if self._graphed is None:
                self._graphed = GraphedForward(self.step, (inp, inp_s))
            g = self._graphed
            g.load(inp, inp_s)
            for _ in range(steps):
                out, out_s = g.replay()
                yield out, out_s
                g.load(out, out_s)                         # device-to-device: next step's input
        else:
            for _ in range(steps):
                out, out_s = self.step(inp, inp_s)
                yield out, out_s
                inp, inp_s = out[0], out_s[0]
