"""Single-GPU estimate of the band mode's compute cost: all bands of a plan run back to back in one process
(LocalComm, no communication), CUDA-event timed.  sum(bands) / un-sharded = kernel-efficiency loss from the
smaller per-rank problem; the rest of the multi-GPU gap is exchange latency and rank skew."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pangu-pytorch-demo_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pangu_oracle as orc  # noqa: E402
from models.pangu_model import PanguModel  # noqa: E402
from pangu_b200.dist import emulate_bands  # noqa: E402

model = PanguModel(device="cpu")
model.load_state_dict(orc.synth_params(seed=0), strict=True)
model = model.cuda().eval().set_compute_dtype("bf16")
inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
args = (inp.cuda(), inp_s.cuda(), tuple(s.cuda() for s in stats), maps.cuda(), const_h.cuda())


def timeit(fn, n=5):
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


with torch.no_grad():
    base = timeit(lambda: model(*args))
    print(f"un-sharded forward: {base:.2f} ms")
    for world in (2, 4, 8):
        ms = timeit(lambda: emulate_bands(model, world, *args))
        print(f"bands x{world} back to back on one GPU: {ms:.2f} ms total = {ms / world:.2f} ms per band on average "
              f"({ms / base:.3f} x un-sharded)")

    # per kernel class: where the small per-rank problem loses (eager launches with CUDA events around each)
    from pangu_b200 import ops  # noqa: E402

    def classes(fn):
        fn()
        torch.cuda.synchronize()
        ops.start_kernel_timing()
        fn()
        torch.cuda.synchronize()
        return ops.stop_kernel_timing()

    t1 = classes(lambda: model(*args))
    t8 = classes(lambda: emulate_bands(model, 8, *args))
    print(f"{'kernel class':42s} {'x1 ms':>8s} {'x8 ms':>8s}  ratio")
    for k in sorted(t1, key=lambda k: -t1[k][1]):
        if k in t8:
            print(f"{k:42s} {t1[k][1]:8.3f} {t8[k][1]:8.3f}  {t8[k][1] / t1[k][1]:.2f}")
    print("only in bands:", {k: round(v[1], 3) for k, v in t8.items() if k not in t1})
