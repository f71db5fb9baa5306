import os, sys, torch
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "pangu-pytorch-demo_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pangu_oracle as orc
from models.pangu_model import PanguModel
from pangu_b200.dist import emulate_bands
from pangu_b200.graph import GraphedForward
model = PanguModel(device="cpu"); model.load_state_dict(orc.synth_params(seed=0), strict=True)
model = model.cuda().eval().set_compute_dtype("bf16")
inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
args = (inp.cuda(), inp_s.cuda(), tuple(s.cuda() for s in stats), maps.cuda(), const_h.cuda())
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
with torch.no_grad():
    g1 = GraphedForward(lambda *a: model(*a), args)
    print("graph un-sharded %.3f ms" % timeit(g1.replay))
    for w in (4, 8):
        gw = GraphedForward(lambda *a, w=w: emulate_bands(model, w, *a), args)
        print("graph bands x%d total %.3f ms" % (w, timeit(gw.replay)))
