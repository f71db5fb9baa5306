"""Multi-GPU check of the latitude-band path (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_bands_check.py

Every rank runs its band through NCCL halo exchanges; rank 0 also runs the un-sharded forward and compares the
stitched result (gathered over NCCL) with it."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pangu-pytorch-demo_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pangu_oracle as orc  # noqa: E402
from models.pangu_model import PanguModel  # noqa: E402
from pangu_b200.dist import BandedPangu, BandPlan  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    model = PanguModel(device="cpu")
    model.load_state_dict(orc.synth_params(seed=0), strict=True)
    model = model.to(dev).eval().set_compute_dtype("bf16")
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    plan = BandPlan(world, rank)
    a, b, m, c = plan.slice_inputs(inp[0], inp_s[0], maps, const_h)
    banded = BandedPangu(model)
    out, out_s = banded(a.to(dev), b.to(dev), stats, m.to(dev), c.to(dev))
    torch.cuda.synchronize()
    # gather the bands on rank 0 (pad to the largest band, all_gather, crop)
    rows = [BandPlan(world, r).pix for r in range(world)]
    mx = max(r1 - r0 for r0, r1 in rows)
    pad = torch.zeros((1, 5, 13, mx, 1440), device=dev)
    pad[:, :, :, :out.shape[3]] = out
    pad_s = torch.zeros((1, 4, mx, 1440), device=dev)
    pad_s[:, :, :out_s.shape[2]] = out_s
    g = [torch.empty_like(pad) for _ in range(world)]
    gs = [torch.empty_like(pad_s) for _ in range(world)]
    dist.all_gather(g, pad)
    dist.all_gather(gs, pad_s)
    ok = True
    if rank == 0:
        full = torch.cat([g[r][:, :, :, :rows[r][1] - rows[r][0]] for r in range(world)], dim=3)
        full_s = torch.cat([gs[r][:, :, :rows[r][1] - rows[r][0]] for r in range(world)], dim=2)
        with torch.no_grad():
            want, want_s = model(inp.to(dev), inp_s.to(dev), tuple(s.to(dev) for s in stats), maps.to(dev), const_h.to(dev))
        e0, e1 = orc.rel_l2(full.cpu(), want.cpu()), orc.rel_l2(full_s.cpu(), want_s.cpu())
        ok = e0 <= 1e-6 and e1 <= 1e-6
        print(f"bands x{world} over NCCL vs un-sharded on rank 0: rel-L2 output {e0:.2e} surface {e1:.2e} -> {'OK' if ok else 'MISMATCH'}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
