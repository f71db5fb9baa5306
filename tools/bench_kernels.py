"""Per-kernel timings at the model's two stage geometries (CUDA events, warm, inputs > L2).

    python tools/bench_kernels.py [mlp] [gemm] [attn]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pangu-pytorch-demo_b200"))
from pangu_b200 import ops  # noqa: E402
from pangu_b200.abi import ACT_GELU  # noqa: E402

STAGES = {"A": (8, 181, 360, 192, 6), "B": (8, 91, 180, 384, 12)}


def timeit(fn, iters=10, warm=3):
    if os.environ.get("BK_ONCE"):                # one launch per kernel: the ncu capture order = the print order
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    which = set(sys.argv[1:]) or {"mlp", "gemm", "attn"}
    g = torch.Generator(device="cuda").manual_seed(0)
    for tag, (Z, H, W, C, heads) in STAGES.items():
        M = Z * H * W
        xb = torch.randn(M, C, device="cuda", generator=g).bfloat16()
        x = torch.randn(M, C, device="cuda", generator=g)
        if "mlp" in which:
            w1 = (torch.randn(4 * C, C, device="cuda", generator=g) * 0.05).bfloat16()
            w2 = (torch.randn(C, 4 * C, device="cuda", generator=g) * 0.05).half()
            b1, b2 = torch.zeros(4 * C, device="cuda"), torch.zeros(C, device="cuda")
            ga, be = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
            ms = timeit(lambda: ops.mlp_ln_residual_bf16(xb, w1, b1, w2, b2, ga, be, x))
            print(f"[{tag}] mlp_fused            {ms:7.3f} ms  {16.0 * M * C * C / ms / 1e9:7.0f} TF/s  {M * C * 12 / ms / 1e6:6.0f} GB/s")
            w2b = w2.bfloat16()

            def unf():
                h = ops.linear(xb, w1, b1, act=ACT_GELU)
                return ops.linear_ln_residual_bf16(h, w2b, b2, ga, be, x)
            ms = timeit(unf)
            print(f"[{tag}] mlp two kernels      {ms:7.3f} ms  {16.0 * M * C * C / ms / 1e9:7.0f} TF/s")
        if "aux" in which:                      # fine-tune epilogues of the CTA-pair GEMM (Mlp.linear1 fwd, linear2 dgrad)
            F = 4 * C
            w1 = (torch.randn(F, C, device="cuda", generator=g) * 0.05).bfloat16()
            b1 = torch.zeros(F, device="cuda")
            dy = torch.randn(M, C, device="cuda", generator=g).bfloat16()
            hp = torch.randn(M, F, device="cuda", generator=g).bfloat16()
            col = torch.zeros(F, device="cuda")
            gb = lambda ms, nb: f"{ms:7.3f} ms  {2.0 * M * C * F / ms / 1e9:7.0f} TF/s  {nb / ms / 1e6:6.0f} GB/s"
            print(f"[{tag}] linear1 plain        " + gb(timeit(lambda: ops.linear(xb, w1, b1)), M * C * 2 + M * F * 2))
            print(f"[{tag}] linear1 gelu         " + gb(timeit(lambda: ops.linear(xb, w1, b1, act=ACT_GELU)), M * C * 2 + M * F * 2))
            print(f"[{tag}] linear1 gelu+pre     " + gb(timeit(lambda: ops.linear_gelu_pre(xb, w1, b1)), M * C * 2 + M * F * 4))
            print(f"[{tag}] linear2 dgrad*gelu'  " + gb(timeit(lambda: ops.linear_gelu_backward(dy, w1, hp, col)), M * C * 2 + M * F * 4))
            dh = torch.randn(M, F, device="cuda", generator=g).bfloat16()
            print(f"[{tag}] gelu_backward kernel " + gb(timeit(lambda: ops.gelu_backward_bf16(dh, hp, col)), M * F * 6))
            del hp, dh, dy
        if "gemm" in which:
            wq = (torch.randn(3 * C, C, device="cuda", generator=g) * 0.05).bfloat16()
            bq = torch.zeros(3 * C, device="cuda")
            ms = timeit(lambda: ops.linear(xb, wq, bq))
            print(f"[{tag}] qkv gemm             {ms:7.3f} ms  {6.0 * M * C * C / ms / 1e9:7.0f} TF/s  {M * C * 8 / ms / 1e6:6.0f} GB/s")
            wp = (torch.randn(C, C, device="cuda", generator=g) * 0.05).bfloat16()
            bp = torch.zeros(C, device="cuda")
            ga, be = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
            ms = timeit(lambda: ops.linear_ln_residual_bf16(xb, wp, bp, ga, be, x))
            print(f"[{tag}] proj+ln gemm         {ms:7.3f} ms  {2.0 * M * C * C / ms / 1e9:7.0f} TF/s  {M * C * 12 / ms / 1e6:6.0f} GB/s")
        if "attn" in which:
            T = (Z // 2) * ((H + 5) // 6)
            qkv = torch.randn(M, 3 * C, device="cuda", generator=g).bfloat16()
            qb = torch.zeros(3 * C, device="cuda")
            eb = (torch.randn(T, heads, 144, 144, device="cuda", generator=g) * 0.02).bfloat16()
            nwin = (W // 12) * T
            for roll in (0, 1):
                ms = timeit(lambda: ops.window_attention_band(qkv, None, qb, eb, Z, H, W, heads, ops.full_band(H), roll, prescaled=True))
                print(f"[{tag}] attention roll={roll} pre {ms:7.3f} ms  {nwin * heads * 4.0 * 144 * 144 * 32 / ms / 1e9:7.0f} TF/s  "
                      f"{(M * C * 8 + eb.numel() * 2) / ms / 1e6:6.0f} GB/s")
                ms = timeit(lambda: ops.window_attention(qkv, qb, eb, Z, H, W, heads, roll))
                print(f"[{tag}] attention roll={roll}     {ms:7.3f} ms  {nwin * heads * 4.0 * 144 * 144 * 32 / ms / 1e9:7.0f} TF/s  "
                      f"{(M * C * 8 + eb.numel() * 2) / ms / 1e6:6.0f} GB/s")
        if "attnbwd" in which:
            T = (Z // 2) * ((H + 5) // 6)
            qkv = (torch.randn(M, 3 * C, device="cuda", generator=g) * 0.5).bfloat16()
            qb = torch.zeros(3 * C, device="cuda")
            eb = (torch.randn(T, heads, 144, 144, device="cuda", generator=g) * 0.02).bfloat16()
            nwin = (W // 12) * T
            d_eb = torch.zeros(T, heads, 144, 144, device="cuda")
            d_pad = torch.zeros(3 * C, device="cuda")
            for roll in (0, 1):
                o, lse = ops.window_attention_train(qkv, qb, eb, Z, H, W, heads, roll)
                do = torch.randn(M, C, device="cuda", generator=g).bfloat16()
                ms = timeit(lambda: ops.window_attention_backward(qkv, qb, eb, o, do, lse, Z, H, W, heads, roll, d_eb, d_pad))
                print(f"[{tag}] attention bwd roll={roll} {ms:7.3f} ms  {nwin * heads * 14.0 * 144 * 144 * 32 / ms / 1e9:7.0f} TF/s  "
                      f"{(M * C * 16 + eb.numel() * 6) / ms / 1e6:6.0f} GB/s")


if __name__ == "__main__":
    main()
