"""Host <-> device streaming around the forecast step.

One 24 h step reads 287 MB of fields and writes 287 MB (fp32); over PCIe that is ~11 ms each way -- as long as
the 18 ms forward.  `StreamedForecaster` keeps three CUDA streams busy so that, in steady state, the H2D copy
of sample i+1, the forward of sample i and the D2H copy of sample i-1 overlap:

    copy-in stream   pinned host input  -> the static inputs of graph (i % 2)
    compute stream   CUDA-graph replay of graph (i % 2)            (one captured graph per slot: no staging copies)
    copy-out stream  static outputs of graph (i % 2) -> pinned host output

    sf = StreamedForecaster(model_or_banded, (input, input_surface, statistics, maps, const_h))
    for k, (x, xs) in enumerate(samples):          # pinned host tensors
        done = sf.submit(x, xs)                    # returns the (host) result of an EARLIER sample, or None
    results = sf.flush()

Each sample still pays its own two transfers; only the idle time between them is removed.
"""
import torch

from .abi import PanguError
from .graph import GraphedForward


class StreamedForecaster:
    def __init__(self, forward, example_args, depth=2):
        """forward(*args) as for GraphedForward; example_args[0:2] are device tensors shaped like one sample's fields.
        One captured graph PER SLOT: the H2D copy lands directly in that graph's static inputs and the D2H copy reads
        its static outputs, so no device-to-device staging copy runs on the compute stream (2 x 287 MB per sample
        saved; costs a second activation pool, ~3 GB of 180 GB)."""
        self.depth = depth
        self.graphs = [GraphedForward(forward, example_args) for _ in range(depth)]
        self.g = self.graphs[0]
        dev = self.g.args[0].device
        self.dev = dev
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        outs = self.g.out if isinstance(self.g.out, (tuple, list)) else (self.g.out,)
        self.host_out = [tuple(torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs) for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]          # H2D into slot's static inputs finished
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]        # replay of slot finished (inputs consumed, outputs valid)
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]         # D2H of slot's static outputs finished
        self.n = 0
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.g.args[0:2])
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in outs)

    def submit(self, inp_host, inp_s_host):
        """Queue one sample (pinned host tensors).  Returns the host outputs of the sample submitted `depth` calls
        earlier once its slot is about to be reused (tuple of pinned tensors, valid until the next submit), else None."""
        if not (inp_host.is_pinned() and inp_s_host.is_pinned()):
            raise PanguError("StreamedForecaster: inputs must be pinned host tensors")
        k = self.n % self.depth
        g = self.graphs[k]
        ready = None
        if self.n >= self.depth:                       # slot k is being reused: its previous result must be on the host
            self.ev_out[k].synchronize()
            ready = self.host_out[k]
        cur = torch.cuda.current_stream(self.dev)
        a, b = g.args[0], g.args[1]
        with torch.cuda.stream(self.s_in):
            if self.n >= self.depth:
                self.s_in.wait_event(self.ev_done[k])  # the previous replay of this slot has read its inputs
            a.copy_(inp_host.reshape(a.shape), non_blocking=True)
            b.copy_(inp_s_host.reshape(b.shape), non_blocking=True)
            self.ev_in[k].record(self.s_in)
        cur.wait_event(self.ev_in[k])
        if self.n >= self.depth:
            cur.wait_event(self.ev_out[k])             # the slot's static outputs have been copied to the host
        out = g.replay()
        out = out if isinstance(out, (tuple, list)) else (out,)
        self.ev_done[k].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_done[k])
            for dst, src in zip(self.host_out[k], out):
                dst.copy_(src, non_blocking=True)
            self.ev_out[k].record(self.s_out)
        self.n += 1
        return ready

    def flush(self):
        """Wait for everything in flight; returns the host outputs of the last min(n, depth) samples, oldest first."""
        torch.cuda.synchronize(self.dev)
        m = min(self.n, self.depth)
        return [self.host_out[(self.n - m + i) % self.depth] for i in range(m)]
