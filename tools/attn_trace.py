"""Pipeline timeline of the tcgen05 window-attention kernel (CTA (0,0,0), first windows): run with PANGU_ATTN_DBG=1."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pangu-pytorch-demo_b200"))
from pangu_b200 import abi, ops  # noqa: E402

os.environ.setdefault("PANGU_ATTN_DBG", "1")
stage = sys.argv[1] if len(sys.argv) > 1 else "B"
Z, H, W, C, heads = (8, 181, 360, 192, 6) if stage == "A" else (8, 91, 180, 384, 12)
N = Z * H * W
T = (Z // 2) * ((H + 5) // 6)
dev = torch.device("cuda:0")
qkv = torch.randn(N, 3 * C, device=dev).bfloat16()
bq = torch.zeros(3 * C, device=dev)
eb = torch.randn(T, heads, 144, 144, device=dev).bfloat16()
for roll in (0, 1):
    for _ in range(3):
        o, _h = ops.window_attention_band(qkv, None, bq, eb, Z, H, W, heads, ops.full_band(H), roll, prescaled=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.window_attention_band(qkv, None, bq, eb, Z, H, W, heads, ops.full_band(H), roll, prescaled=True)
    e1.record()
    torch.cuda.synchronize()
    print("stage %s roll %d: %.3f ms / launch" % (stage, roll, e0.elapsed_time(e1) / 10))
buf = (ctypes.c_int64 * 128)()
abi.check(abi.lib().pangu_debug_attn_trace(buf, 128), "trace")
names = ["tma", "S issued", "p_full", "PV issued", "s_full", "max pass", "max exch", "P stored", "epi(i-2)", "-", "tail start",
         "tail end", "swap start", "swap end"]
t00 = min(buf[w * 16] for w in range(8))
for w in range(8):
    row = [buf[w * 16 + i] for i in range(14)]
    print("slot %d: " % w + "  ".join("%s %d" % (n, v - t00) for n, v in zip(names, row) if n != "-"))
