"""Print the fused-Mlp kernel's pipeline timeline (see tc_mlp.cu g_mlp_trace)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pangu-pytorch-demo_b200"))
os.environ["PANGU_MLP_DBG"] = str(16 | int(os.environ.get("PANGU_MLP_DBG", "0")))
from pangu_b200 import abi, ops  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 384
M = 148 * 128 * 4
g = torch.Generator(device="cuda").manual_seed(0)
xb = torch.randn(M, C, device="cuda", generator=g).bfloat16()
x = torch.randn(M, C, device="cuda", generator=g)
w1 = (torch.randn(4 * C, C, device="cuda", generator=g) * 0.05).bfloat16()
w2 = (torch.randn(C, 4 * C, device="cuda", generator=g) * 0.05).half()
b1, b2 = torch.zeros(4 * C, device="cuda"), torch.zeros(C, device="cuda")
ga, be = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
PROJ = len(sys.argv) > 2 and sys.argv[2] == "proj"
wp = (torch.randn(C, C, device="cuda", generator=g) * 0.05).bfloat16()
for _ in range(3):
    if PROJ:
        ops.attn_proj_mlp_ln_bf16(xb, wp, b2, ga, be, x, w1, b1, w2, b2, ga, be)
    else:
        ops.mlp_ln_residual_bf16(xb, w1, b1, w2, b2, ga, be, x)
torch.cuda.synchronize()
buf = (ctypes.c_int64 * 512)()
abi.check(abi.lib().pangu_debug_mlp_trace(ctypes.cast(buf, ctypes.c_void_p), 512), "trace")
nch = 4 * C // 64
t0 = min(v for v in buf[:nch * 8] if v > 0)
print("chunk | mma: p_full  g2_issued g1_issued | epi: h_full  loaded  gelu_done  p_stored   (cycles since first event)")
for j in range(nch):
    r = [buf[j * 8 + k] - t0 if buf[j * 8 + k] else -1 for k in range(8)]
    print(f"{j:3d}   | {r[0]:8d} {r[1]:8d} {r[2]:8d} | {r[4]:8d} {r[5]:8d} {r[6]:8d} {r[7]:8d}")

print("tile-level events of CTA 0, row tiles 1 and 2 (cycles since tile 1's x load was issued):")
names = ["x load issued", "x_full passed (MMA)", "y_empty passed (MMA)", "y_full committed (MMA issue)", "LN: x_empty passed", "LN: y_full passed",
         "LN: stats done", "LN: units done", "LN: stores read, xs_free", "GELU: first h_full", "GELU: last h_full",
         "PROJ: proj MMAs issued", "PROJ: x1_ready passed (MMA)", "PROJ: LN1 y_full passed", "PROJ: LN1 done"]
base = buf[256 + 16]
for n in (1, 2):
    for k, nm in enumerate(names):
        v = buf[256 + n * 16 + k]
        print(f"  tile {n}  {nm:32s} {v - base if v else -1:9d}")
nu = 3 if C == 192 else 12
print("LN units of warp 12, row tile 1: start, tile landed, prefetch issued, tmem loaded, math done, staged, fenced, stores issued (cycles since unit 0 start)")
b0 = buf[400]
for u in range(nu):
    if 400 + u * 8 + 7 >= 512:
        break
    r = [buf[400 + u * 8 + k] - b0 for k in (0, 1, 2, 3, 6, 7, 4, 5)]
    print(f"  unit {u:2d}: " + " ".join(f"{x:8d}" for x in r))
