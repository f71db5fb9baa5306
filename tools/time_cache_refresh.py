"""What the operand-cache refresh after an optimizer step costs (bf16 / fp16 / transposed weight shadows, pre-scaled bias tables):
fwd+bwd with valid caches vs fwd+bwd+Adam (every cache stale at the next forward) vs Adam alone."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pangu-pytorch-demo_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pangu_oracle as orc  # noqa: E402
from models.pangu_model import PanguModel  # noqa: E402

dev = torch.device("cuda:0")
model = PanguModel(device="cpu")
model.load_state_dict(orc.synth_params(seed=0), strict=True)
model = model.to(dev).train()
inp, inp_s, stats, maps, const_h = (t.to(dev) if torch.is_tensor(t) else tuple(s.to(dev) for s in t) for t in orc.synth_inputs(seed=1))
tgt, tgt_s = torch.randn_like(inp), torch.randn_like(inp_s)
opt = torch.optim.Adam(model.parameters(), lr=1e-6, fused=True)


def fwd_bwd():
    model.zero_grad(set_to_none=True)
    o, os_ = model(inp, inp_s, stats, maps, const_h)
    ((o - tgt).abs().mean() + 0.25 * (os_ - tgt_s).abs().mean()).backward()


def timeit(fn, n=4):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


a = timeit(fwd_bwd)
b = timeit(lambda: (fwd_bwd(), opt.step()))
c = timeit(opt.step)
print("fwd+bwd (caches valid) %.2f ms | fwd+bwd+Adam %.2f ms | Adam alone %.2f ms | cache refresh = %.2f ms per step" % (a, b, c, b - a - c))
