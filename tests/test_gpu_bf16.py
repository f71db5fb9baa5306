"""bf16 tensor-core path (tcgen05 GEMMs, tensor-core window attention) against the oracle / goldens.
Tolerance: relative L2 <= 2e-2 for model-level outputs (north_star); unit tests are much tighter
because they compare against the same computation on bf16-rounded operands."""
import pytest
import torch

import pangu_oracle as orc
from util_gpu import check_digest, load_params

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.fixture(scope="module")
def L():
    import models.layers as layers
    return layers


@pytest.mark.parametrize("M,K,N", [(128, 64, 64), (128, 192, 192), (1000, 192, 576), (333, 112, 192),
                                   (4133, 384, 1152), (640, 768, 384), (257, 1536, 256), (129, 384, 160),
                                   (300, 384, 64), (521, 192, 768),
                                   # M >= 2048, K in {192, 384}, N % 192 == 0: the A-resident CTA-pair kernel (tc_gemm2.cu)
                                   (2048, 192, 192), (3001, 192, 576), (2500, 384, 768), (40000, 384, 1152)])
def test_tc_linear(M, K, N):
    from pangu_b200 import ops
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.05).bfloat16()
    b = torch.randn(N, generator=g)
    want = a.double() @ w.double().t() + b.double()
    got = ops.linear(a.cuda(), w.cuda(), b.cuda(), out_dtype=torch.float32).cpu()
    assert orc.rel_l2(got, want) <= 3e-5, "fp32-out GEMM on bf16 operands must be exact to fp32 accumulation"
    got = ops.linear(a.cuda(), w.cuda(), b.cuda()).cpu()
    assert got.dtype == torch.bfloat16 and orc.rel_l2(got, want) <= 3e-3


def test_tc_linear_gelu():
    from pangu_b200 import ops
    from pangu_b200.abi import ACT_GELU
    g = torch.Generator().manual_seed(5)
    a = torch.randn(700, 192, generator=g).bfloat16()
    w = (torch.randn(768, 192, generator=g) * 0.2).bfloat16()
    b = torch.randn(768, generator=g)
    want = torch.nn.functional.gelu(a.double() @ w.double().t() + b.double())
    got = ops.linear(a.cuda(), w.cuda(), b.cuda(), act=ACT_GELU, out_dtype=torch.float32).cpu()
    assert float((got.double() - want).abs().max()) <= 2e-3      # tanh-form GELU fit + MUFU.TANH
    assert orc.rel_l2(got, want) <= 1e-3


@pytest.mark.parametrize("C,K", [(192, 192), (192, 768), (384, 384), (384, 1536)])
def test_tc_linear_ln_residual(C, K):
    from pangu_b200 import ops
    g = torch.Generator().manual_seed(C + K)
    M = 777
    a = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(C, K, generator=g) * 0.05).bfloat16()
    b, gamma, beta = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    res = torch.randn(M, C, generator=g)
    y = a.double() @ w.double().t() + b.double()
    want = res.double() + torch.nn.functional.layer_norm(y, (C,), gamma.double(), beta.double(), 1e-5)
    x, xb = ops.linear_ln_residual_bf16(a.cuda(), w.cuda(), b.cuda(), gamma.cuda(), beta.cuda(), res.cuda())
    assert orc.rel_l2(x.cpu(), want) <= 3e-5
    assert orc.rel_l2(xb.cpu(), want) <= 3e-3


@pytest.mark.parametrize("C,M", [(192, 256), (192, 1000), (384, 128), (384, 777), (384, 20000)])
def test_tc_fused_mlp_ln_residual(C, M):
    """Fused Mlp + norm2 + residual kernel (tcgen05 cta_group::2, hidden kept in TMEM as fp16) against
    fp64 on the same bf16/fp16 operands."""
    from pangu_b200 import ops
    g = torch.Generator().manual_seed(C + M)
    a = torch.randn(M, C, generator=g).bfloat16()
    w1 = (torch.randn(4 * C, C, generator=g) * 0.08).bfloat16()
    w2 = (torch.randn(C, 4 * C, generator=g) * 0.05).half()
    b1, b2 = torch.randn(4 * C, generator=g) * 0.5, torch.randn(C, generator=g)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    res = torch.randn(M, C, generator=g)
    h = torch.nn.functional.gelu(a.double() @ w1.double().t() + b1.double())
    y = h @ w2.double().t() + b2.double()
    want = res.double() + torch.nn.functional.layer_norm(y, (C,), gamma.double(), beta.double(), 1e-5)
    x, xb = ops.mlp_ln_residual_bf16(a.cuda(), w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), gamma.cuda(),
                                     beta.cuda(), res.cuda())
    assert orc.rel_l2(x.cpu(), want) <= 2e-3            # fp16 GELU / hidden rounding only
    assert orc.rel_l2(xb.cpu(), want) <= 4e-3


@pytest.mark.parametrize("Z,H,W,C,heads", [(8, 181, 24, 192, 6), (8, 91, 24, 384, 12)])
@pytest.mark.parametrize("roll", [0, 1])
def test_attention_bf16_vs_fp32_kernel(Z, H, W, C, heads, roll):
    """Tensor-core attention against the fp32 SIMT kernel on the same bf16-rounded q/k/v."""
    from pangu_b200 import ops
    g = torch.Generator().manual_seed(11)
    T = (Z // 2) * ((H + 5) // 6)
    qkv = torch.randn(Z * H * W, 3 * C, generator=g).bfloat16()
    qb = torch.randn(3 * C, generator=g) * 0.1
    eb = (torch.randn(T, heads, 144, 144, generator=g) * 0.5).bfloat16()
    want = ops.window_attention(qkv.float().cuda(), qb.bfloat16().float().cuda(), eb.float().cuda(), Z, H, W, heads, roll)
    got = ops.window_attention(qkv.cuda(), qb.cuda(), eb.cuda(), Z, H, W, heads, roll)
    assert orc.rel_l2(got.float().cpu(), want.cpu()) <= 6e-3      # P and the output are rounded to bf16


def test_attention_truncated_probabilities_are_unbiased():
    """The tcgen05 kernel packs P to bf16 by truncation (PRMT instead of F2FP, tc_attention2.cu) and multiplies 1 / rowsum by the
    mean truncation loss (kTruncFix = 1.002826): the output must carry no magnitude bias against the fp32 kernel -- a missing
    correction shows up as -2.8e-3 here, a doubled one as +2.8e-3."""
    from pangu_b200 import ops
    Z, H, W, C, heads = 8, 91, 24, 384, 12
    g = torch.Generator().manual_seed(12)
    T = (Z // 2) * ((H + 5) // 6)
    qkv = torch.randn(Z * H * W, 3 * C, generator=g).bfloat16()
    qb = torch.randn(3 * C, generator=g) * 0.1
    eb = (torch.randn(T, heads, 144, 144, generator=g) * 0.5).bfloat16()
    for roll in (0, 1):
        want = ops.window_attention(qkv.float().cuda(), qb.bfloat16().float().cuda(), eb.float().cuda(), Z, H, W, heads, roll).double()
        got = ops.window_attention(qkv.cuda(), qb.cuda(), eb.cuda(), Z, H, W, heads, roll).double()
        bias = float(((got - want) * torch.sign(want)).sum() / want.abs().sum())
        assert abs(bias) <= 5e-4, f"roll {roll}: signed magnitude bias {bias:.2e}"


@pytest.mark.parametrize("tag,dim,heads,Z,H,W,pfx", [
    ("blockA", 192, 6, 8, 181, 24, "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."),
    ("blockB", 384, 12, 8, 91, 24, "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock3."),
])
def test_block_bf16(L, goldens, tag, dim, heads, Z, H, W, pfx):
    params = orc.synth_params(seed=0, only_prefix=pfx)
    blk = load_params(L.EarthSpecificBlock(dim, 0.0, heads, "cpu"), params, pfx)
    L.set_compute_dtype(blk, "bf16")
    g = torch.Generator().manual_seed(7)
    if tag == "blockB":
        torch.randn(1, 8 * 181 * 24, 192, generator=g)
    x = torch.randn(1, Z * H * W, dim, generator=g)
    for roll in (False, True):
        with torch.no_grad():
            y = blk(x.cuda(), Z, H, W, roll)
        err = check_digest(goldens, f"{tag}.roll{int(roll)}", y, TOL)
        print(f"{tag} roll={roll} bf16 rel-L2 {err:.2e}")


def test_full_model_bf16_vs_reference_golden(goldens):
    if "output.val" not in goldens:
        pytest.skip("goldens were generated with --skip-full")
    from models.pangu_model import PanguModel
    params = orc.synth_params(seed=0)
    model = PanguModel(device="cpu")
    model.load_state_dict(params, strict=True)
    model = model.cuda().eval().set_compute_dtype("bf16")
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
    stats = tuple(s.cuda() for s in stats)
    with torch.no_grad():
        out, out_s = model(inp.cuda(), inp_s.cuda(), stats, maps.cuda(), const_h.cuda())
    e1 = check_digest(goldens, "output", out, TOL)
    e2 = check_digest(goldens, "output_surface", out_s, TOL)
    print(f"bf16 full model rel-L2: output {e1:.2e} surface {e2:.2e}")


def test_batched_input_is_a_loop_over_samples(L):
    """The reference is only correct for batch 1 (its window reverse assumes B = 1); the B200 modules take
    [B, N, C] like the reference's signatures and process the samples one at a time: row b of the batched result is
    bit-identical to the single-sample call."""
    pfx = "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."
    Z, H, W, dim, heads = 8, 181, 24, 192, 6
    blk = load_params(L.EarthSpecificBlock(dim, 0.0, heads, "cpu"), orc.synth_params(seed=0, only_prefix=pfx), pfx)
    L.set_compute_dtype(blk, "bf16")
    x = torch.randn(2, Z * H * W, dim, generator=torch.Generator().manual_seed(11)).cuda()
    with torch.no_grad():
        y2 = blk(x, Z, H, W, True)
        y0, y1 = blk(x[0:1], Z, H, W, True), blk(x[1:2], Z, H, W, True)
    assert y2.shape == x.shape
    assert torch.equal(y2[0], y0[0]) and torch.equal(y2[1], y1[0])
    mlp = blk.linear
    with torch.no_grad():                                      # stand-alone Mlp is forward-only (it raises when a graph is wanted)
        m2 = mlp(x[:, :4096])
        assert torch.equal(m2[1], mlp(x[1:2, :4096])[0])
    with pytest.raises(L.PanguError):
        mlp(x[1:2, :4096])


def test_repeated_launches_are_bit_identical(L):
    """compute-sanitizer is closed on this GPU pool (profiles/r2_sanitizer_closed.md), so the hand-rolled mbarrier / TMEM
    hand-over protocols of the tcgen05 kernels get a cheaper race check: a data race between warp roles (a tile read
    before its load landed, an accumulator drained while still being written, a staging tile refilled under a pending bulk
    store) shows up as run-to-run differences.  Every forward kernel of a block -- QKV GEMM, tcgen05 attention (rolled:
    masked tiles, tail rows), GEMM + LayerNorm epilogue, fused MLP with its GELU / LayerNorm warps -- is run 6 times on
    the same inputs, full row-tile counts so that every CTA walks several tiles, and must return identical bits."""
    for dim, heads, Z, H, W, pfx in ((192, 6, 8, 181, 72, "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."),
                                     (384, 12, 8, 91, 180, "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock3.")):
        blk = load_params(L.EarthSpecificBlock(dim, 0.0, heads, "cpu"), orc.synth_params(seed=0, only_prefix=pfx), pfx)
        L.set_compute_dtype(blk, "bf16")
        x = torch.randn(1, Z * H * W, dim, generator=torch.Generator().manual_seed(3)).cuda()
        with torch.no_grad():
            ref = blk(x, Z, H, W, True)
            for _ in range(5):
                assert torch.equal(blk(x, Z, H, W, True), ref), f"C={dim}: run-to-run difference"
        assert torch.isfinite(ref).all()


@pytest.mark.parametrize("M", [131040, 256 * 74 + 77, 4000, 100])
def test_fused_block_tail_equals_the_two_kernels(M):
    """pangu_attn_proj_mlp_bf16 (attention.linear2 + norm1 + shortcut + Mlp + norm2 + shortcut in ONE kernel, C = 384) against
    pangu_linear_ln_residual_bf16 followed by pangu_mlp_ln_residual_bf16: the same arithmetic (the fused kernel only keeps
    x1 on chip); the LayerNorm row sums are grouped over 4 instead of 2 warps, so the last bits of mean / variance may
    differ by an ulp, which flips the bf16 rounding of x1 in 2^-16 of the elements: rel-L2 <= 2e-5 between the two (measured 7e-6), run-to-run bit identity of the fused kernel itself -- at the full stage-B token
    count (several row tiles per CTA pair, odd pairs staggered), a ragged count, fewer tiles than CTA pairs, and less than
    one row tile."""
    from pangu_b200 import ops
    C = 384
    g = torch.Generator().manual_seed(M)
    dev = "cuda"
    o = torch.randn(M, C, generator=g).to(dev).bfloat16()
    x = (torch.randn(M, C, generator=g) * 2).to(dev)
    wp = (torch.randn(C, C, generator=g) * C ** -0.5).to(dev).bfloat16()
    w1 = (torch.randn(4 * C, C, generator=g) * C ** -0.5).to(dev).bfloat16()
    w2 = (torch.randn(C, 4 * C, generator=g) * (4 * C) ** -0.5).to(dev).half()
    bp, b1, b2 = (torch.randn(n, generator=g).to(dev) * 0.1 for n in (C, 4 * C, C))
    g1, be1, g2, be2 = (1 + 0.1 * torch.randn(C, generator=g)).to(dev), 0.1 * torch.randn(C, generator=g).to(dev), \
        (1 + 0.1 * torch.randn(C, generator=g)).to(dev), 0.1 * torch.randn(C, generator=g).to(dev)
    x1, x1b = ops.linear_ln_residual_bf16(o, wp, bp, g1, be1, x)
    want, want_b = ops.mlp_ln_residual_bf16(x1b, w1, b1, w2, b2, g2, be2, x1)
    first = None
    for _ in range(3):                                       # repeated: the scratch tiles / barrier phases are reused correctly
        got, got_b = ops.attn_proj_mlp_ln_bf16(o, wp, bp, g1, be1, x, w1, b1, w2, b2, g2, be2)
        e = orc.rel_l2(got, want)
        assert e <= 2e-5, e
        assert orc.rel_l2(got_b.float(), want_b.float()) <= 1e-4          # a few 1-ulp bf16 rounding flips
        assert torch.equal(got_b, got.bfloat16())                        # the shadow is the rounded fp32 result
        if first is None:
            first = got.clone()
        assert torch.equal(got, first)
    # fp32 torch reference of the whole tail on the same operands (bounds the pair itself): rel-L2 <= 5e-3
    ref1 = x + torch.nn.functional.layer_norm(o.float() @ wp.float().t() + bp, (C,), g1, be1, 1e-5)
    h = torch.nn.functional.gelu(ref1.bfloat16().float() @ w1.float().t() + b1)
    ref = ref1 + torch.nn.functional.layer_norm(h @ w2.float().t() + b2, (C,), g2, be2, 1e-5)
    assert orc.rel_l2(got, ref) <= 5e-3
