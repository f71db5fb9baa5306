"""Host-side logic of the latitude-band sharding (pangu_b200/dist.py) on CPU: band plans tile every stage
exactly, every rolled window's source rows are own / halo / pad rows, and the neighbour exchange moves the
right rows between ranks (world_size 2, gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pangu_oracle as orc
from pangu_b200.dist import TOK_H, BandPlan, DistComm, LocalComm


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_band_plans_tile_the_grid(world):
    plans = [BandPlan(world, r) for r in range(world)]
    for stage in ("A", "B"):
        H = TOK_H[stage]
        nH = (H + 5) // 6
        rows = [p.rows[stage] for p in plans]
        assert rows[0][0] == 0 and rows[-1][1] == H
        assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
        assert all(r[0] % 6 == 0 for r in rows), "bands start on a window edge"
        for roll in (0, 1):
            seen = []
            for p in plans:
                b = p.band(stage, roll, "sendback")
                regular = b.nhw - b.wrap
                seen += list(range(b.hw0, b.hw0 + regular)) + ([nH - 1] if b.wrap else [])
                assert (b.h0, b.h0 + b.hrows) == p.rows[stage]
                assert b.halo == (3 if (roll and not p.last) else 0) and b.halo_lo == 0
                r = p.band(stage, roll, "redundant")      # additionally the window straddling the northern edge
                extra = 1 if (roll and not p.first) else 0
                assert (r.hw0, r.nhw, r.halo_lo, r.halo, r.wrap) == (b.hw0 - extra, b.nhw + extra, 3 * extra, b.halo, b.wrap)
            assert sorted(seen) == list(range(nH)), (stage, roll, seen)
    pix = [p.pix for p in plans]
    assert pix[0][0] == 0 and pix[-1][1] == 721 and all(pix[i][1] == pix[i + 1][0] for i in range(world - 1))
    for p in plans:                                     # stage transitions are row-local
        a0, a1 = p.rows["A"]
        assert p.rows["B"] == (a0 // 2, (a1 + 1) // 2) and a0 % 2 == 0
        assert p.pix[0] == 4 * a0 and (p.pix[1] + 3) // 4 == a1


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("stage,W", [("A", 24), ("B", 12)])
def test_rolled_windows_only_need_own_halo_or_pad_rows(world, stage, W):
    """Against the oracle's closed form of pad + roll + partition (models/layers.py:224-262)."""
    H = TOK_H[stage]
    nH = (H + 5) // 6
    for roll in (False, True):
        src = orc.window_source_index(8, H, W, roll)               # [nLon, T, 144] token index or -1
        rows = np.where(src >= 0, (src // W) % H, -1).reshape(src.shape[0], 4, nH, 144)
        owner_of_output = {}
        for r in range(world):
            p = BandPlan(world, r)
            b = p.band(stage, int(roll), "sendback")
            hws = list(range(b.hw0, b.hw0 + b.nhw - b.wrap)) + ([nH - 1] if b.wrap else [])
            for hw in hws:
                need = np.unique(rows[:, :, hw])
                need = need[need >= 0]
                ok = (need >= b.h0) & (need < b.h0 + b.hrows + b.halo)
                assert ok.all(), (world, stage, roll, r, hw, need)
                for h in need:
                    owner_of_output.setdefault(int(h), []).append(r)
            # "redundant" scheme: own rows are covered by windows whose rows are own / 3-row halos / pad
            rb = p.band(stage, int(roll), "redundant")
            covered = set()
            for hw in list(range(rb.hw0, rb.hw0 + rb.nhw - rb.wrap)) + ([nH - 1] if rb.wrap else []):
                need = np.unique(rows[:, :, hw])
                need = need[need >= 0]
                assert ((need >= rb.h0 - rb.halo_lo) & (need < rb.h0 + rb.hrows + rb.halo)).all()
                covered |= set(int(h) for h in need)
            assert set(range(rb.h0, rb.h0 + rb.hrows)) <= covered
        # every real row is produced exactly once (8 z-planes share the same h)
        assert sorted(owner_of_output) == list(range(H))
        assert all(len(set(v)) == 1 for v in owner_of_output.values())


def test_local_comm_shifts():
    c = LocalComm(3)
    sends = [torch.full((2,), float(r)) for r in range(3)]
    up = c.shift_up(sends, [sends[0], sends[0], None])
    assert up[0][0] == 1 and up[1][0] == 2 and up[2] is None
    down = c.shift_down(sends, [None, sends[0], sends[0]])
    assert down[0] is None and down[1][0] == 0 and down[2][0] == 1


def test_local_comm_swap_edges():
    c = LocalComm(3)
    first = [torch.full((1,), 10.0 + r) for r in range(3)]
    last = [torch.full((1,), 20.0 + r) for r in range(3)]
    got = c.swap_edges(first, last, first)
    assert got[0][0][0] == 11 and got[0][1] is None          # rank 0: south = rank 1's first rows, no north
    assert got[1][0][0] == 12 and got[1][1][0] == 20
    assert got[2][0] is None and got[2][1][0] == 21


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = DistComm()
        plan = BandPlan(world, rank)
        hr, W, F = plan.nrows("A"), 12, 4
        band = (torch.arange(8 * hr * W * F, dtype=torch.float32) + 1e6 * rank).view(8 * hr * W, F)
        first = band.view(8, hr, W, F)[:, :3].reshape(8 * 3 * W, F).contiguous()
        like = torch.empty(8 * 3 * W, F)
        up = comm.shift_up([None if plan.first else first], [None if plan.last else like])[0]
        down = comm.shift_down([None if plan.last else first * 2], [None if plan.first else like])[0]
        south, north = comm.swap_edges([first], [first + 7], [like])[0]
        q.put((rank, None if up is None else up.clone(), None if down is None else down.clone(), first.clone(),
               None if south is None else south.clone(), None if north is None else north.clone()))
    finally:
        dist.destroy_process_group()


def test_dist_comm_neighbour_exchange_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, up, down, first, south, north = q.get(timeout=120)
        got[rank] = (up, down, first, south, north)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(got[0][0], got[1][2]), "rank 0 receives rank 1's first three rows"
    assert got[1][0] is None and got[0][1] is None
    assert torch.equal(got[1][1], got[0][2] * 2), "rank 1 receives what rank 0 computed for it"
    assert torch.equal(got[0][3], got[1][2]) and got[0][4] is None      # swap_edges: south halo of rank 0
    assert torch.equal(got[1][4], got[0][2] + 7) and got[1][3] is None  # north halo of rank 1


# ------------------------------------------------------------------------------------------ fine-tune DP (configs[4])
def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pangu_b200.dist import GradientAllReducer
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4))
        red = GradientAllReducer(net, bucket_mb=16 * 4 / (1 << 20) * 4)      # several tiny buckets
        outs = []
        for it in range(2):                                                   # buckets are reusable step after step
            net.zero_grad(set_to_none=True)
            x = torch.full((3, 8), float(rank + 1 + it))
            net(x).square().sum().backward()
            local = [p.grad.numpy().copy() for p in net.parameters()]           # numpy: pickled by value
            red.finish()
            outs.append((local, [p.grad.numpy().copy() for p in net.parameters()]))
        # gradient accumulation (the reference's loss / accumulation_steps, several backward passes per optimiser step)
        net.zero_grad(set_to_none=True)
        xs = [torch.full((3, 8), float(rank + 1 + k)) for k in range(2)]
        with red.no_sync():
            (net(xs[0]).square().sum() / 2).backward()
        (net(xs[1]).square().sum() / 2).backward()
        local = [p.grad.numpy().copy() for p in net.parameters()]               # accumulated local sums
        red.finish()
        outs.append((local, [p.grad.numpy().copy() for p in net.parameters()]))
        # a second synchronising backward without finish() must raise, not mis-reduce
        net.zero_grad(set_to_none=True)
        net(xs[0]).square().sum().backward()
        try:
            net(xs[1]).square().sum().backward()
            raised = False
        except RuntimeError as e:
            raised = "no_sync" in str(e)
        q.put((rank, len(red.buckets), outs, raised))
    finally:
        dist.destroy_process_group()


def test_gradient_allreducer_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, nb, outs, raised = q.get(timeout=120)
        got[rank] = outs
        assert nb > 1 and raised
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for it in range(3):                                                       # it == 2: two accumulated micro-steps
        mean = [(a + b) / 2 for a, b in zip(got[0][it][0], got[1][it][0])]
        for r in range(world):
            for m, g in zip(mean, got[r][it][1]):
                assert np.allclose(m, g, rtol=1e-6, atol=1e-7)
