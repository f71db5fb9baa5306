"""GPU bring-up probe for the tcgen05 GEMM / tensor-core attention kernels.  Each case runs in its own
process (a trapped kernel poisons the CUDA context) and prints a compact report.

    python tools/tc_debug.py            # all cases
    python tools/tc_debug.py CASE       # one case in-process
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pangu-pytorch-demo_b200"))

CASES = ["ident_64", "ident_k128", "rand_128x192x192", "rand_tail", "rand_big", "gelu", "ln192", "ln384",
         "attn_a", "attn_b_roll", "mlp192_256", "mlp192", "mlp384_256", "mlp384", "mlp384_big"]


def report(name, got, want):
    import torch
    got, want = got.double().cpu(), want.double().cpu()
    err = float((got - want).norm() / want.norm().clamp_min(1e-30))
    bad = (got - want).abs() > 1e-2 * want.abs().max()
    print(f"[{name}] rel-L2 {err:.3e}  bad {int(bad.sum())}/{bad.numel()}  nan {int(torch.isnan(got).sum())}")
    if err > 1e-2:
        torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
        print(" got [0,:16]", got[0, :16])
        print(" want[0,:16]", want[0, :16])
        print(" got [1,:16]", got[1, :16])
        print(" want[1,:16]", want[1, :16])
        rows_bad = bad.any(1).nonzero().flatten()
        cols_bad = bad.any(0).nonzero().flatten()
        print(" bad rows (first 16):", rows_bad[:16].tolist(), " count", rows_bad.numel())
        print(" bad cols (first 16):", cols_bad[:16].tolist(), " count", cols_bad.numel())
    return err


def run_case(name):
    import torch
    from pangu_b200 import ops
    from pangu_b200.abi import ACT_GELU
    g = torch.Generator().manual_seed(0)
    if name in ("ident_64", "ident_k128"):
        K = 64 if name == "ident_64" else 128
        a = torch.randint(-8, 9, (128, K), generator=g).float()
        w = torch.eye(64, K)
        if K == 128:
            w[:, 64:] = torch.eye(64) * 2
        got = ops.linear(a.bfloat16().cuda(), w.bfloat16().cuda(), None, out_dtype=torch.float32)
        report(name, got, a @ w.t())
    elif name.startswith("rand"):
        M, K, N = {"rand_128x192x192": (128, 192, 192), "rand_tail": (333, 112, 576), "rand_big": (20000, 384, 1536)}[name]
        a = torch.randn(M, K, generator=g).bfloat16()
        w = (torch.randn(N, K, generator=g) * 0.05).bfloat16()
        b = torch.randn(N, generator=g)
        want = a.double() @ w.double().t() + b.double()
        report(name + ".f32", ops.linear(a.cuda(), w.cuda(), b.cuda(), out_dtype=torch.float32), want)
        report(name + ".bf16", ops.linear(a.cuda(), w.cuda(), b.cuda()), want)
    elif name == "gelu":
        a = torch.randn(700, 192, generator=g).bfloat16()
        w = (torch.randn(768, 192, generator=g) * 0.2).bfloat16()
        b = torch.randn(768, generator=g)
        want = torch.nn.functional.gelu(a.double() @ w.double().t() + b.double())
        report(name, ops.linear(a.cuda(), w.cuda(), b.cuda(), act=ACT_GELU, out_dtype=torch.float32), want)
    elif name in ("ln192", "ln384"):
        C = int(name[2:])
        M, K = 777, 2 * C
        a = torch.randn(M, K, generator=g).bfloat16()
        w = (torch.randn(C, K, generator=g) * 0.05).bfloat16()
        b, gamma, beta = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
        res = torch.randn(M, C, generator=g)
        y = a.double() @ w.double().t() + b.double()
        want = res.double() + torch.nn.functional.layer_norm(y, (C,), gamma.double(), beta.double(), 1e-5)
        x, xb = ops.linear_ln_residual_bf16(a.cuda(), w.cuda(), b.cuda(), gamma.cuda(), beta.cuda(), res.cuda())
        report(name + ".f32", x, want)
        report(name + ".bf16", xb, want)
    elif name.startswith("mlp"):
        C = int(name[3:6])
        M = {"256": 256, "": 777, "big": 40000}[name[7:]]
        a = torch.randn(M, C, generator=g).bfloat16()
        w1 = (torch.randn(4 * C, C, generator=g) * 0.08).bfloat16()
        w2 = (torch.randn(C, 4 * C, generator=g) * 0.05).half()
        b1, b2 = torch.randn(4 * C, generator=g) * 0.5, torch.randn(C, generator=g)
        gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
        res = torch.randn(M, C, generator=g)
        h = torch.nn.functional.gelu(a.double() @ w1.double().t() + b1.double())
        hb = h.float().half().double()                          # the kernel feeds GEMM2 with an fp16 hidden
        y = hb @ w2.double().t() + b2.double()
        want = res.double() + torch.nn.functional.layer_norm(y, (C,), gamma.double(), beta.double(), 1e-5)
        x, xb = ops.mlp_ln_residual_bf16(a.cuda(), w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), gamma.cuda(),
                                         beta.cuda(), res.cuda())
        torch.cuda.synchronize()
        report(name + ".f32", x, want)
        report(name + ".bf16", xb, want)
    elif name.startswith("attn"):
        Z, H, W, C, heads, roll = (8, 181, 24, 192, 6, 0) if name == "attn_a" else (8, 91, 24, 384, 12, 1)
        T = (Z // 2) * ((H + 5) // 6)
        qkv = torch.randn(Z * H * W, 3 * C, generator=g).bfloat16()
        qb = (torch.randn(3 * C, generator=g) * 0.1).bfloat16().float()
        eb = (torch.randn(T, heads, 144, 144, generator=g) * 0.5).bfloat16()
        want = ops.window_attention(qkv.float().cuda(), qb.cuda(), eb.float().cuda(), Z, H, W, heads, roll)
        got = ops.window_attention(qkv.cuda(), qb.cuda(), eb.cuda(), Z, H, W, heads, roll)
        report(name, got, want)
    torch.cuda.synchronize()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for c in CASES:
            r = subprocess.run(["timeout", "120", sys.executable, os.path.abspath(__file__), c], capture_output=True, text=True)
            out = (r.stdout + r.stderr).strip().splitlines()
            keep = [l for l in out if l.startswith("[") or l.startswith(" ") or "rror" in l or "timed out" in l]
            print(f"== {c}: exit {r.returncode}")
            print("\n".join(keep[-30:]))
