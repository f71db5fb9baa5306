"""Neighbour exchange micro-benchmark (torchrun, 2+ ranks): the band mode's halo swap = one batched NCCL send/recv pair per
direction of `--mb` MB (stage A / B qkv halo: 9.95 MB).  CUDA-event time per exchange, max over ranks."""
import argparse
import os

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=9.95)
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(a.mb * 1e6 / 2)
send_n, send_s = torch.randn(n, device="cuda").bfloat16(), torch.randn(n, device="cuda").bfloat16()
recv_n, recv_s = torch.empty_like(send_n), torch.empty_like(send_n)


def swap():
    opsl = []
    if rank > 0:
        opsl += [dist.P2POp(dist.isend, send_n, rank - 1), dist.P2POp(dist.irecv, recv_n, rank - 1)]
    if rank < world - 1:
        opsl += [dist.P2POp(dist.isend, send_s, rank + 1), dist.P2POp(dist.irecv, recv_s, rank + 1)]
    for w in dist.batch_isend_irecv(opsl):
        w.wait()


for _ in range(5):
    swap()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    swap()
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / a.iters], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    us = float(t) * 1000
    print(f"NCCL_NCHANNELS_PER_PEER={os.environ.get('NCCL_NCHANNELS_PER_PEER')} NCCL_MIN_P2P_NCHANNELS={os.environ.get('NCCL_MIN_P2P_NCHANNELS')} "
          f"NCCL_P2P_USE_CUDA_MEMCPY={os.environ.get('NCCL_P2P_USE_CUDA_MEMCPY')}: {us:.1f} us per swap of {a.mb} MB each way = {a.mb * 1e6 / us / 1e3:.0f} GB/s per direction")
dist.barrier()
dist.destroy_process_group()
