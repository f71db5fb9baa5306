"""CUDA-graph replay of a fixed-shape forward.

One Pangu forward is ~100 kernel launches issued from Python through ctypes (~10 us each) plus, in the
latitude-band mode, 16 NCCL neighbour exchanges.  On one GPU the 20 ms of device work hides that; on 4-8 GPUs the
per-rank device time drops to a few ms and the host becomes the limiter.  Capturing the whole step (kernels +
NCCL point-to-point) in a CUDA graph removes the host from the loop: one `cudaGraphLaunch` per step.

    fwd = GraphedForward(model, (input, input_surface, statistics, maps, const_h))    # warm-up + capture
    out, out_surface = fwd(input, input_surface)      # copies the two fields into the static buffers, replays

Outputs are STATIC tensors that the next call overwrites (clone them to keep a result).  Shapes, dtypes and the
model's weights' addresses must not change; after an in-place weight update call `fwd.recapture()` (the bf16
weight caches are refreshed outside the graph).
"""
import torch

from . import abi, ops
from .abi import PanguError


def _clone_static(x):
    if torch.is_tensor(x):
        if not x.is_cuda:
            raise PanguError("GraphedForward: every tensor argument must already be on the CUDA device")
        return x.clone()
    if isinstance(x, (tuple, list)):
        return type(x)(_clone_static(v) for v in x)
    return x


class GraphedForward:
    def __init__(self, fn, example_args, warmup=2, dynamic=(0, 1)):
        """fn(*args) -> tensor or tuple of tensors.  `dynamic` = positions of the tensor arguments that change
        from call to call (the rest -- statistics, constant maps -- are captured as constants)."""
        self.fn = fn
        self.args = [_clone_static(a) for a in example_args]
        self.dynamic = tuple(dynamic)
        self.warmup = warmup
        self.graph = None
        self.out = None
        self.launches_per_replay = 0
        self.recapture()

    def recapture(self):
        # programmatic dependent launch (include/pangu_b200.h: pangu_set_pdl) is baked into the captured kernel nodes: the
        # replayed step is alone on the device, which is where it pays (emulated 8-band step 23.05 -> 22.44 ms)
        prev = abi.lib().pangu_set_pdl(1)
        try:
            return self._recapture()
        finally:
            abi.lib().pangu_set_pdl(prev)

    def _recapture(self):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):                       # builds weight caches, NCCL communicators, smem attributes
                self.fn(*self.args)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = ops.LAUNCHES
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = self.fn(*self.args)
        self.launches_per_replay = ops.LAUNCHES - before
        return self

    def replay(self):
        """Re-run on whatever the static input buffers hold."""
        self.graph.replay()
        ops.LAUNCHES += self.launches_per_replay
        return self.out

    def load(self, *inputs):
        """Copy this call's dynamic inputs (device or pinned-host tensors) into the static buffers."""
        if len(inputs) != len(self.dynamic):
            raise PanguError(f"GraphedForward: expected {len(self.dynamic)} dynamic inputs, got {len(inputs)}")
        for pos, t in zip(self.dynamic, inputs):
            dst = self.args[pos]
            if t.shape != dst.shape:
                if t.numel() != dst.numel():
                    raise PanguError(f"GraphedForward: input {pos} has shape {tuple(t.shape)}, captured {tuple(dst.shape)}")
                t = t.reshape(dst.shape)
            dst.copy_(t, non_blocking=True)

    def __call__(self, *inputs):
        self.load(*inputs)
        return self.replay()
