"""Fine-tune step timing on one GPU: forward + backward of PanguModel.train() at full resolution with the reference's
L1 loss (models/pangu_sample.py:205-218), CUDA events, plus the per-kernel-class breakdown of ops.py."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pangu-pytorch-demo_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pangu_oracle as orc  # noqa: E402
from models.pangu_model import PanguModel  # noqa: E402
from pangu_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
model = PanguModel(device="cpu")
model.load_state_dict(orc.synth_params(seed=0), strict=True)
model = model.to(dev).train()
inp, inp_s, stats, maps, const_h = (t.to(dev) if torch.is_tensor(t) else tuple(s.to(dev) for s in t) for t in orc.synth_inputs(seed=1))
tgt, tgt_s = torch.randn_like(inp), torch.randn_like(inp_s)


def step():
    model.zero_grad(set_to_none=True)
    o, os_ = model(inp, inp_s, stats, maps, const_h)
    loss = (o - tgt).abs().mean() + 0.25 * (os_ - tgt_s).abs().mean()
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record()
for _ in range(n):
    loss = step()
e1.record()
torch.cuda.synchronize()
print("fwd+bwd ms/step %.2f  loss %.5f  peak mem %.1f GB" % (e0.elapsed_time(e1) / n, float(loss), torch.cuda.max_memory_allocated() / 2**30))
ops.start_kernel_timing()
step()
t = ops.stop_kernel_timing()
tot = sum(v[1] for v in t.values())
for k, v in sorted(t.items(), key=lambda kv: -kv[1][1]):
    print("%-44s n=%3d %8.3f ms %5.1f%%  %7.1f TF/s %7.1f GB/s" % (k, v[0], v[1], 100 * v[1] / tot, v[2] / v[1] / 1e9 if v[1] else 0, v[3] / v[1] / 1e6 if v[1] else 0))
print("sum of kernel classes %.2f ms" % tot)
