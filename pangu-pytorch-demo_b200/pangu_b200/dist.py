"""Latitude-band spatial sharding of the Pangu forward (BASELINE configs[3]) -- new in this repo,
the reference has no spatial parallelism.

One process per GPU; rank r of P in {1, 2, 4, 8} owns a contiguous band of latitude rows in every stage:

    stage A (C = 192)  24 * (8 / P) token rows      (4 * (8 / P) six-row windows)
    stage B (C = 384)  12 * (8 / P) token rows      (2 * (8 / P) windows)
    pixels             96 * (8 / P) rows

the last rank taking the remainder (up to row 180 / 90 / 720 plus the reference's zero padding).  Everything
except the shifted-window attention is row-local: patch embed / recover are 4x4 pixel patches, down / up-sample
pair rows (2i, 2i+1), linears / LayerNorm / Mlp are per token, un-rolled windows (models/layers.py:253-262) are
aligned to the band edges.  In a rolled block (torch.roll by (-1,-3,-6), models/layers.py:237) the h-window hw
reads rows [6 hw + 3, 6 hw + 9), so one window straddles every band edge.  Two exchange schemes:

  "redundant" (default)  ONE bidirectional exchange per rolled block: each rank receives the first 3 qkv rows of
                         its southern neighbour and the last 3 qkv rows of its northern neighbour, BOTH ranks run
                         the straddling window and each keeps only its own rows;
  "sendback"             each rank runs the windows whose first source row it owns: receives the southern
                         neighbour's first 3 qkv rows, then returns the attention output of those 3 rows
                         (two dependent exchanges per rolled block, no duplicated window).

The wrap-around window (3 pad rows + global rows 0..2) is isolated by the -100 shift mask
(models/layers.py:187-216), so rank 0 computes it locally.

The exchanges are NCCL point-to-point (`torch.distributed.batch_isend_irecv`) of 3*8*W*3C bf16 (about 10 MB per
direction); a `LocalComm` runs all ranks of a plan inside ONE process (tests on a single GPU: the banded result
must equal the un-sharded forward bit for bit).
"""
import os

import torch
import torch.distributed as dist

from . import functional as PF
from . import ops
from .abi import Band, PanguError

TOK_H = {"A": 181, "B": 91}
TOK_W = {"A": 360, "B": 180}
PIX_H, Z = 721, 8


class BandPlan:
    """Row ranges of one rank.  All ranges are [begin, end) in GLOBAL row coordinates."""

    def __init__(self, world, rank):
        if world not in (1, 2, 4, 8):
            raise PanguError(f"latitude-band sharding supports 1, 2, 4 or 8 ranks, not {world}")
        if not 0 <= rank < world:
            raise PanguError(f"rank {rank} outside world {world}")
        self.world, self.rank = world, rank
        self.first, self.last = rank == 0, rank == world - 1
        per_a = 24 * (8 // world)                      # stage-A rows per rank; a multiple of the 6-row window
        a0 = per_a * rank
        a1 = TOK_H["A"] if self.last else per_a * (rank + 1)
        self.rows = {"A": (a0, a1), "B": (a0 // 2, (a1 + 1) // 2)}
        self.pix = (4 * a0, min(4 * a1, PIX_H))
        self.map_rows = (4 * a0, 4 * a1)               # the constant maps are stored already padded (724 rows)

    def nrows(self, stage):
        r = self.rows[stage]
        return r[1] - r[0]

    def band(self, stage, roll, scheme="redundant"):
        """pangu_band of this rank for one block of `stage` (models/layers.py:228: 5 pad rows, 6-row windows)."""
        h0, h1 = self.rows[stage]
        nH = (TOK_H[stage] + 5) // 6
        hw0 = h0 // 6
        if not roll:
            hw1 = nH if self.last else h1 // 6
            return Band(h0, h1 - h0, hw0, hw1 - hw0, 0, 0, 0)
        # rolled: regular windows [hw0, hw1); the global window nH-1 wraps around to rows 0..2 -> rank 0
        hw1 = nH - 1 if self.last else h1 // 6
        wrap = 1 if self.first else 0
        halo = 0 if self.last else 3
        halo_lo = 0
        if scheme == "redundant" and not self.first:      # also run the window straddling the northern edge
            hw0, halo_lo = hw0 - 1, 3
        return Band(h0, h1 - h0, hw0, hw1 - hw0 + wrap, wrap, halo, halo_lo)

    def slice_inputs(self, input, input_surface, maps, const_h):
        """Band of the full-grid arrays (contiguous copies): what this rank is handed in a sharded deployment."""
        p0, p1 = self.pix
        m0, m1 = self.map_rows
        return (input[..., p0:p1, :].contiguous(), input_surface[..., p0:p1, :].contiguous(),
                maps[..., m0:m1, :].contiguous(), const_h[..., p0:p1, :].contiguous())


# ----------------------------------------------------------------------------------------------
# communicators: shift_up (rank r receives what rank r+1 sends), shift_down (r receives from r-1)
# ----------------------------------------------------------------------------------------------
class DistComm:
    """torch.distributed point-to-point neighbour exchange (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise PanguError("DistComm needs an initialised torch.distributed process group")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def _peer(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def _exchange(self, send, recv_like, to_rank, from_rank):
        opsl, out = [], None
        if send is not None and 0 <= to_rank < self.world:
            opsl.append(dist.P2POp(dist.isend, send, self._peer(to_rank), self.group))
        if recv_like is not None and 0 <= from_rank < self.world:
            out = torch.empty_like(recv_like)
            opsl.append(dist.P2POp(dist.irecv, out, self._peer(from_rank), self.group))
        if opsl:
            for w in dist.batch_isend_irecv(opsl):
                w.wait()
        return out

    def shift_up(self, sends, recv_like):
        """sends[0] goes to rank-1; returns what rank+1 sent (shaped like recv_like[0]) or None on the last rank."""
        return [self._exchange(sends[0], recv_like[0], self.rank - 1, self.rank + 1)]

    def shift_down(self, sends, recv_like):
        return [self._exchange(sends[0], recv_like[0], self.rank + 1, self.rank - 1)]

    def swap_edges(self, first_rows, last_rows, like):
        """One batched exchange: my first rows go north, my last rows go south.
        -> [(rows received from the south, rows received from the north)] (None at the grid edges)."""
        opsl, from_s, from_n = [], None, None
        if self.rank > 0:
            opsl.append(dist.P2POp(dist.isend, first_rows[0], self._peer(self.rank - 1), self.group))
            from_n = torch.empty_like(like[0])
            opsl.append(dist.P2POp(dist.irecv, from_n, self._peer(self.rank - 1), self.group))
        if self.rank < self.world - 1:
            opsl.append(dist.P2POp(dist.isend, last_rows[0], self._peer(self.rank + 1), self.group))
            from_s = torch.empty_like(like[0])
            opsl.append(dist.P2POp(dist.irecv, from_s, self._peer(self.rank + 1), self.group))
        if opsl:
            for w in dist.batch_isend_irecv(opsl):
                w.wait()
        return [(from_s, from_n)]


class LocalComm:
    """All ranks of a plan in one process: lists are indexed by rank."""

    def __init__(self, world):
        self.world, self.rank = world, None

    def shift_up(self, sends, recv_like):
        return [sends[r + 1] if r + 1 < self.world and recv_like[r] is not None else None for r in range(self.world)]

    def shift_down(self, sends, recv_like):
        return [sends[r - 1] if r >= 1 and recv_like[r] is not None else None for r in range(self.world)]

    def swap_edges(self, first_rows, last_rows, like):
        return [(first_rows[r + 1] if r + 1 < self.world else None, last_rows[r - 1] if r >= 1 else None)
                for r in range(self.world)]


# ----------------------------------------------------------------------------------------------
class _Worker:
    """State of one rank's band while it moves through the network."""

    def __init__(self, model, plan):
        self.m, self.plan = model, plan
        self.x = self.xb = self.skip = self.skip_b = self.qkv = self.o = self.halo_qkv = self.halo_lo_qkv = self.halo_o = None

    # --- row-local stages
    def embed(self, inp, inp_s, stats, maps, const_h):
        self.x, self.xb = PF.patch_embed_forward(self.m._input_layer, inp, inp_s, stats, maps, const_h, "bf16")

    def downsample(self):
        self.skip, self.skip_b = self.x, self.xb
        self.x, self.xb = self.m.downsample.forward_sample(self.x, Z, self.plan.nrows("A"), TOK_W["A"])

    def upsample(self):
        self.x, self.xb = PF.upsample_forward(self.m.upsample, self.x, "bf16", self.xb, Z=Z, H2=self.plan.nrows("B"),
                                              W2=TOK_W["B"], H=self.plan.nrows("A"))

    def recover(self):
        lat = self.plan.pix[1] - self.plan.pix[0]
        return PF.patch_recover_forward(self.m._output_layer, self.x, Z, self.plan.nrows("A"), TOK_W["A"], "bf16",
                                        skip=self.skip, lat=lat, xb=self.xb, skip_b=self.skip_b)

    # --- one EarthSpecificBlock in three phases around the two neighbour exchanges
    def block_qkv(self, blk, stage):
        att, wc = blk.attention, blk._wcache
        if self.xb is None:
            self.xb = ops.cast_bf16(self.x)
        w_qkv, b_qkv, _ = PF.attention_operands(att, wc)
        self.qkv = ops.linear(self.xb, w_qkv, b_qkv)

    def first_rows(self, t, stage, rows=3, col0=0):
        """The first `rows` latitude rows of a band tensor [Z*hrows*W, F] as a contiguous [Z*rows*W, F - col0] block
        (col0 = C on a qkv tensor: only the K and V columns)."""
        hr, W = self.plan.nrows(stage), TOK_W[stage]
        return t.view(Z, hr, W, t.shape[-1])[:, :rows, :, col0:].reshape(Z * rows * W, t.shape[-1] - col0).contiguous()

    def last_rows(self, t, stage, rows=3, col0=0):
        hr, W = self.plan.nrows(stage), TOK_W[stage]
        return t.view(Z, hr, W, t.shape[-1])[:, hr - rows:, :, col0:].reshape(Z * rows * W, t.shape[-1] - col0).contiguous()

    def block_attend(self, blk, stage, roll, scheme):
        att, wc = blk.attention, blk._wcache
        band = self.plan.band(stage, roll, scheme)
        _, b_qkv, eb = PF.attention_operands(att, wc)
        self.o, self.halo_o = ops.window_attention_band(
            self.qkv, self.halo_qkv, b_qkv, eb, Z, TOK_H[stage], TOK_W[stage], att.head_number, band, roll,
            halo_lo_qkv=self.halo_lo_qkv, return_halo=(scheme == "sendback"), prescaled=True,
            exact_max=PF.attention_exact_max(att, wc), halo_kv=(scheme == "redundant" and HALO_KV_ONLY))
        self.qkv = self.halo_qkv = self.halo_lo_qkv = None

    def block_finish(self, blk, stage, o_first):
        att, mlp, wc = blk.attention, blk.linear, blk._wcache
        if o_first is not None:                          # rows 0..2 of this band were computed by the northern neighbour
            hr, W = self.plan.nrows(stage), TOK_W[stage]
            self.o.view(Z, hr, W, self.o.shape[-1])[:, :3].copy_(o_first.view(Z, 3, W, self.o.shape[-1]))
        if PF.FUSED_MLP and PF.FUSED_PROJ and self.o.shape[-1] == 384:
            self.x, self.xb = ops.attn_proj_mlp_ln_bf16(
                self.o, wc.bf16("a2", PF.lin_w(att.linear2)), PF._f(PF.lin_b(att.linear2)), PF._f(blk.norm1.weight),
                PF._f(blk.norm1.bias), self.x, wc.bf16("m1", PF.lin_w(mlp.linear1)), PF._f(PF.lin_b(mlp.linear1)),
                wc.f16("m2h", PF.lin_w(mlp.linear2)), PF._f(PF.lin_b(mlp.linear2)), PF._f(blk.norm2.weight),
                PF._f(blk.norm2.bias), eps1=blk.norm1.eps, eps2=blk.norm2.eps)
            self.o = self.halo_o = None
            return
        x1, x1b = ops.linear_ln_residual_bf16(self.o, wc.bf16("a2", PF.lin_w(att.linear2)), PF._f(PF.lin_b(att.linear2)),
                                              PF._f(blk.norm1.weight), PF._f(blk.norm1.bias), self.x, eps=blk.norm1.eps)
        self.o = self.halo_o = None
        self.x, self.xb = ops.mlp_ln_residual_bf16(x1b, wc.bf16("m1", PF.lin_w(mlp.linear1)), PF._f(PF.lin_b(mlp.linear1)),
                                                   wc.f16("m2h", PF.lin_w(mlp.linear2)), PF._f(PF.lin_b(mlp.linear2)),
                                                   PF._f(blk.norm2.weight), PF._f(blk.norm2.bias), x1, eps=blk.norm2.eps)


SCHEMES = ("redundant", "sendback")
# "redundant" scheme: exchange only the K and V columns of the halo rows (0 = whole qkv rows, for A/B measurements)
HALO_KV_ONLY = os.environ.get("PANGU_B200_HALO_KV", "1") != "0"


def _run(model, workers, comm, inputs, scheme="redundant"):
    """Drive the workers (one per local rank) through the network; `comm` moves the halos between them."""
    if scheme not in SCHEMES:
        raise PanguError(f"exchange scheme {scheme!r}: expected one of {SCHEMES}")
    n = len(workers)
    for w, (inp, inp_s, stats, maps, const_h) in zip(workers, inputs):
        w.embed(inp, inp_s, stats, maps, const_h)
    stages = ["A", "B", "B", "A"]
    for li, layer in enumerate(model.layers):
        stage = stages[li]
        if li == 1:
            for w in workers:
                w.downsample()
        if li == 3:
            for w in workers:
                w.upsample()
        for bi, blk in enumerate(layer.blocks):
            roll = 1 if bi % 2 == 1 else 0
            for w in workers:
                w.block_qkv(blk, stage)
            exchange = roll and comm.world > 1
            if exchange and scheme == "redundant":
                # each rank runs the straddling window for its OWN rows only: the neighbours' queries are never needed, so only
                # the K and V columns travel (2/3 of the bytes)
                c0 = workers[0].qkv.shape[-1] // 3 if HALO_KV_ONLY else 0
                firsts_ = [None if w.plan.first else w.first_rows(w.qkv, stage, col0=c0) for w in workers]
                lasts_ = [None if w.plan.last else w.last_rows(w.qkv, stage, col0=c0) for w in workers]
                like = [_halo_like(w.qkv, stage, cols=w.qkv.shape[-1] - c0) for w in workers]
                for w, (south, north) in zip(workers, comm.swap_edges(firsts_, lasts_, like)):
                    w.halo_qkv, w.halo_lo_qkv = south, north
            elif exchange:
                sends = [None if w.plan.first else w.first_rows(w.qkv, stage) for w in workers]
                like = [None if w.plan.last else _halo_like(w.qkv, stage) for w in workers]
                halos = comm.shift_up(sends, like)
                for w, h in zip(workers, halos):
                    w.halo_qkv = h
            for w in workers:
                w.block_attend(blk, stage, roll, scheme)
            firsts = [None] * n
            if exchange and scheme == "sendback":
                sends = [w.halo_o for w in workers]
                like = [None if w.plan.first else _halo_like(w.o, stage) for w in workers]
                firsts = comm.shift_down(sends, like)
            for w, f in zip(workers, firsts):
                w.block_finish(blk, stage, f)
    return [w.recover() for w in workers]


def _halo_like(t, stage, rows=3, cols=None):
    return torch.empty((Z * rows * TOK_W[stage], t.shape[-1] if cols is None else cols), dtype=t.dtype, device=t.device)


def _check_model(model):
    if getattr(model, "training", False) and torch.is_grad_enabled():
        pass                                             # forward only; gradients are never recorded here
    if model._mode() != "bf16":
        raise PanguError("latitude-band sharding runs the bf16 tensor-core path (set_compute_dtype('bf16'))")


class BandedPangu:
    """This rank's band of `PanguModel.forward` inside a torch.distributed job (weights replicated).

        plan = BandPlan(world, rank); band_inputs = plan.slice_inputs(input[0], input_surface[0], maps, const_h)
        out, out_surface = BandedPangu(model)(*band_inputs[:2], statistics, *band_inputs[2:])

    Inputs are ONE sample's band: input [5,13,rows,1440], input_surface [4,rows,1440], maps [.,3,4*tok_rows,1440],
    const_h [...,13,rows,1440]; outputs [1,5,13,rows,1440] and [1,4,rows,1440] are the same rows of the forecast."""

    def __init__(self, model, group=None, comm=None, scheme="redundant"):
        _check_model(model)
        self.model = model
        self.scheme = scheme
        self.comm = comm if comm is not None else DistComm(group)
        self.plan = BandPlan(self.comm.world, self.comm.rank)

    @torch.no_grad()
    def __call__(self, inp, inp_s, statistics, maps, const_h):
        dev = inp.device
        stats = tuple(s.to(dev) for s in statistics)
        w = _Worker(self.model, self.plan)
        return _run(self.model, [w], self.comm, [(inp.float().contiguous(), inp_s.float().contiguous(), stats,
                                                  maps.float().contiguous(), const_h.float().contiguous())],
                    self.scheme)[0]


@torch.no_grad()
def emulate_bands(model, world, input, input_surface, statistics, maps, const_h, scheme="redundant"):
    """Run all `world` bands of one full-grid sample in THIS process (LocalComm) and stitch the outputs:
    (output [1,5,13,721,1440], output_surface [1,4,721,1440]).  Used by the single-GPU parity tests."""
    _check_model(model)
    dev = input.device
    stats = tuple(s.to(dev) for s in statistics)
    plans = [BandPlan(world, r) for r in range(world)]
    inputs = []
    for p in plans:
        a, b, m, c = p.slice_inputs(input.float(), input_surface.float(), maps.float(), const_h.float())
        inputs.append((a.reshape(5, 13, -1, 1440), b.reshape(4, -1, 1440), stats, m, c))
    outs = _run(model, [_Worker(model, p) for p in plans], LocalComm(world), inputs, scheme)
    return torch.cat([o[0] for o in outs], dim=3), torch.cat([o[1] for o in outs], dim=2)


# ------------------------------------------------------------------------------------------ fine-tune data parallelism
class GradientAllReducer:
    """Data-parallel fine-tuning (BASELINE configs[4]; the reference wraps the model in DDP, finetune/finetune_fully.py:220):
    one sample per rank, mean of the 223 fp32 gradients over the ranks.

    The gradients are packed into a few flat fp32 buckets in REVERSE parameter order (the order in which the
    backward produces them) and each bucket is all-reduced with one NCCL call as soon as its last gradient exists
    (post-accumulate-grad hooks), on the side stream NCCL uses, so that the reduction of the late blocks overlaps the
    backward kernels of the early ones.  The 16 Earth-specific bias tables are 91 % of the 1.1 GB and get buckets of
    their own.  `finish()` waits for the reductions and copies the means back into `p.grad` (call it before the
    optimiser step).  Works with any backend (gloo on CPU in the tests).

    Gradient accumulation (the reference divides the loss by `accumulation_steps` and runs several backward passes per
    optimiser step): run every micro-step but the last under `with reducer.no_sync():` -- the hooks then leave the
    gradients to accumulate locally in `p.grad`, exactly like DDP's no_sync -- and the last one normally: its hooks see
    the accumulated sums, which is what gets reduced.  A second synchronising backward without `finish()` in between
    raises instead of silently dropping or double-averaging gradients."""

    def __init__(self, model, group=None, bucket_mb=64.0):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.buckets, cur, cur_n = [], [], 0
        limit = int(bucket_mb * (1 << 20) / 4)
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > limit:
                self.buckets.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device) for b in self.buckets]
        self.where = {}
        for bi, b in enumerate(self.buckets):
            off = 0
            for p in b:
                self.where[p] = (bi, off)
                off += p.numel()
        self.pending = [len(b) for b in self.buckets]
        self.works = [None] * len(self.buckets)
        self.sync = True
        self.hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def no_sync(self):
        """Context manager: backward passes inside accumulate into `p.grad` without any communication."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            old, self.sync = self.sync, False
            try:
                yield self
            finally:
                self.sync = old
        return ctx()

    def _on_grad(self, p):
        if not self.sync:
            return
        bi, off = self.where[p]
        if self.pending[bi] <= 0:
            raise RuntimeError("GradientAllReducer: a second synchronising backward() before finish(); run the earlier "
                               "micro-steps under `with reducer.no_sync():` (gradient accumulation) or call finish() "
                               "after every backward()")
        self.flat[bi][off:off + p.numel()].copy_(p.grad.reshape(-1))
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self.flat[bi].div_(self.world)
            self.works[bi] = self.dist.all_reduce(self.flat[bi], group=self.group, async_op=True)

    def finish(self):
        if any(n != 0 for n in self.pending):
            missing = [i for i, n in enumerate(self.pending) if n != 0]
            raise RuntimeError(f"GradientAllReducer.finish(): buckets {missing} are incomplete -- some parameters got no gradient")
        for bi, b in enumerate(self.buckets):
            self.works[bi].wait()
            off = 0
            for p in b:
                p.grad.copy_(self.flat[bi][off:off + p.numel()].view_as(p.grad))
                off += p.numel()
        self.pending = [len(b) for b in self.buckets]
        self.works = [None] * len(self.buckets)

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []
