"""torch.Tensor wrappers over the C ABI.  PyTorch is used for device memory and streams only."""
import torch

from . import abi
from .abi import ACT_GELU, ACT_NONE, BF16, F32, Band, Geom

_DT = {torch.float32: F32, torch.bfloat16: BF16}

LAUNCHES = 0          # kernels launched through this module (bench.py reads it as `gpu_launches`)
_TIMING = None        # list of (op-class, start event, end event, flops, bytes) while kernel timing is on


def start_kernel_timing():
    global _TIMING
    _TIMING = []


def stop_kernel_timing():
    """-> {op-class: (calls, total_ms, flops, bytes)} measured with CUDA events on the launching stream."""
    global _TIMING
    rec, _TIMING = _TIMING or [], None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, fl, by in rec:
        c = out.setdefault(name, [0, 0.0, 0.0, 0.0])
        c[0] += 1
        c[1] += e0.elapsed_time(e1)
        c[2] += fl
        c[3] += by
    return {k: tuple(v) for k, v in out.items()}


_DEV = None           # device of the tensors of the op being issued (set by _chk / _on); the library itself never calls
                      # cudaSetDevice, so the launch is made with that device current and on ITS current stream


def _on(device):
    """Declare the device of the op being issued (ops without tensor inputs)."""
    global _DEV
    _DEV = torch.device(device)
    if _DEV.type != "cuda":
        raise abi.PanguError(f"device {device}: the B200 path runs on CUDA only (no CPU fallback)")
    if _DEV.index is None:
        _DEV = torch.device("cuda", torch.cuda.current_device())
    return _DEV


def _launch(fn, args):
    if _DEV is not None and _DEV.index != torch.cuda.current_device():
        with torch.cuda.device(_DEV):            # model on cuda:1 while cuda:0 is current (reference: torch.device('cuda:%d'))
            abi.check(getattr(abi.lib(), fn)(*args), fn)
    else:
        abi.check(getattr(abi.lib(), fn)(*args), fn)


def _call(name, fn, args, kernels=1, flops=0.0, nbytes=0.0):
    global LAUNCHES
    LAUNCHES += kernels
    if _TIMING is None:
        _launch(fn, args)
        return
    st = torch.cuda.current_stream(_DEV)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    _launch(fn, args)
    e1.record(st)
    _TIMING.append((name, e0, e1, flops, nbytes))


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    """The current stream of the device the op's tensors live on (not of the current device)."""
    return torch.cuda.current_stream(_DEV).cuda_stream


def _chk(t, dtype=None, name="tensor"):
    global _DEV
    if not t.is_cuda:
        raise abi.PanguError(f"{name} must be a CUDA tensor: the B200 path has no CPU fallback")
    _DEV = t.device
    if not t.is_contiguous():
        raise abi.PanguError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise abi.PanguError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def geom(Z, H, W, C, heads=None):
    return Geom(Z, H, W, C, heads if heads is not None else C // 32)


def window_counts(Z, H, W):
    Hp = H + 5
    if Z % 2 or Hp % 6 or W % 12:
        raise abi.PanguError(f"grid ({Z},{H},{W}) is not tileable by (2,6,12) windows after padding H by 5")
    return W // 12, (Z // 2) * (Hp // 6)


# ------------------------------------------------------------------ index kernels
def window_partition(x, Z, H, W, roll):
    """x [Z*H*W, C] -> [nLon, T, 144, C]  (models/layers.py:224-262)."""
    _chk(x, name="x")
    C = x.shape[-1]
    nLon, T = window_counts(Z, H, W)
    out = torch.empty((nLon, T, 144, C), dtype=x.dtype, device=x.device)
    g = geom(Z, H, W, C)
    _call("window_partition", "pangu_window_partition", (_ptr(x), _ptr(out), g, int(roll), x.element_size(), _stream(),))
    return out


def window_reverse(win, Z, H, W, roll):
    """[nLon, T, 144, C] -> x [Z*H*W, C]  (models/layers.py:269-293)."""
    _chk(win, name="win")
    C = win.shape[-1]
    out = torch.empty((Z * H * W, C), dtype=win.dtype, device=win.device)
    g = geom(Z, H, W, C)
    _call("window_reverse", "pangu_window_reverse", (_ptr(win), _ptr(out), g, int(roll), win.element_size(), _stream(),))
    return out


def window_source_index(Z, H, W, roll, device):
    nLon, T = window_counts(Z, H, W)
    device = _on(device)
    out = torch.empty((nLon, T, 144), dtype=torch.int64, device=device)
    g = geom(Z, H, W, 32)
    _call("window_source_index", "pangu_window_source_index", (_ptr(out), g, int(roll), _stream(),))
    return out


def shift_mask(Z, H, W, device):
    nLon, T = window_counts(Z, H, W)
    device = _on(device)
    out = torch.empty((T, 144, 144), dtype=torch.float32, device=device)
    g = geom(Z, H, W, 32)
    _call("shift_mask", "pangu_shift_mask", (_ptr(out), g, _stream(),))
    return out


def position_index(device):
    device = _on(device)
    out = torch.empty((144 * 144,), dtype=torch.int64, device=device)
    _call("position_index", "pangu_position_index", (_ptr(out), _stream(),))
    return out


# ------------------------------------------------------------------ dense
def linear(a, w, bias=None, act=ACT_NONE, out_dtype=None, out=None):
    """out = act(a @ w.T + bias); a [M,K], w [N,K] (fp32 both, or bf16 both)."""
    _chk(a, name="a")
    _chk(w, a.dtype, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    out_dtype = out_dtype or a.dtype
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    tag = "gemm_%s[K=%d,N=%d%s]" % ("bf16" if a.dtype == torch.bfloat16 else "f32", K, N, ",gelu" if act == ACT_GELU else "")
    _call(tag, "pangu_linear", (_ptr(a), K, _ptr(w), _ptr(bias), _ptr(out), N, M, K, N, act, _DT[a.dtype],
                                     _DT[out.dtype], _stream(),), flops=2.0 * M * K * N,
          nbytes=float(a.numel() * a.element_size() + out.numel() * out.element_size() + w.numel() * w.element_size()))
    return out


def linear_ex(a, w, bias=None, a2=None, act=ACT_NONE, out_dtype=torch.float32, out=None, shadow=None, want_shadow=False):
    """bf16 tensor-core linear: out = act(cat(a, a2) @ w.T + bias); optionally also a bf16 shadow of an fp32 output.
    a [M,K1], a2 [M,K-K1] or None, w [N,K].  -> out, or (out, shadow) when want_shadow / shadow is given."""
    _chk(a, torch.bfloat16, "a")
    _chk(w, torch.bfloat16, "w")
    M, K1 = a.shape
    N, K = w.shape
    if a2 is not None:
        _chk(a2, torch.bfloat16, "a2")
        if a2.shape != (M, K - K1):
            raise abi.PanguError(f"linear_ex: a2 must be [{M}, {K - K1}]")
    elif K1 != K:
        raise abi.PanguError("linear_ex: a and w disagree on K")
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    if want_shadow and shadow is None:
        shadow = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    tag = "gemm_bf16[K=%d,N=%d%s%s]" % (K, N, ",cat" if a2 is not None else "", ",shadow" if shadow is not None else "")
    _call(tag, "pangu_linear_bf16_ex",
          (_ptr(a), K1, _ptr(a2), (K - K1) if a2 is not None else 0, K1 if a2 is not None else 0, _ptr(w), _ptr(bias),
           _ptr(out), _ptr(shadow), N, M, K, N, act, _DT[out.dtype], _stream(),), flops=2.0 * M * K * N,
          nbytes=float(M * K * 2 + out.numel() * out.element_size() + w.numel() * 2 + (M * N * 2 if shadow is not None else 0)))
    return (out, shadow) if shadow is not None else out


def ln_residual(y, gamma, beta, residual=None, want_f32=True, want_bf16=False, eps=1e-5):
    """residual + LayerNorm(y)*gamma + beta -> (fp32 or None, bf16 or None)."""
    _chk(y, name="y")
    M, C = y.shape
    x_out = torch.empty((M, C), dtype=torch.float32, device=y.device) if want_f32 else None
    xb = torch.empty((M, C), dtype=torch.bfloat16, device=y.device) if want_bf16 else None
    if residual is not None:
        _chk(residual, torch.float32, "residual")
    _call("ln_residual[C=%d]" % C, "pangu_ln_residual", (_ptr(y), _DT[y.dtype], _ptr(gamma), _ptr(beta), _ptr(residual), _ptr(x_out),
                                          _ptr(xb), M, C, eps, _stream(),),
          nbytes=float(M * C * (y.element_size() + 4 * (residual is not None) + 4 * want_f32 + 2 * want_bf16)))
    return x_out, xb


def linear_ln_residual_bf16(a, w, bias, gamma, beta, residual, want_bf16=True, eps=1e-5):
    """residual + LN(a @ w.T + bias)*gamma + beta with the LayerNorm fused in the GEMM epilogue."""
    _chk(a, torch.bfloat16, "a")
    _chk(w, torch.bfloat16, "w")
    M, K = a.shape
    C = w.shape[0]
    x_out = torch.empty((M, C), dtype=torch.float32, device=a.device)
    xb = torch.empty((M, C), dtype=torch.bfloat16, device=a.device) if want_bf16 else None
    _call("gemm_bf16_ln[K=%d,N=%d]" % (K, C), "pangu_linear_ln_residual_bf16", (_ptr(a), K, _ptr(w), _ptr(bias), _ptr(gamma), _ptr(beta),
                                                      _ptr(residual), _ptr(x_out), _ptr(xb), M, K, C, eps, _stream(),),
          flops=2.0 * M * K * C, nbytes=float(M * K * 2 + C * K * 2 + M * C * (8 + 2 * want_bf16)))
    return x_out, xb


def mlp_ln_residual_bf16(xb, w1, b1, w2, b2, gamma, beta, residual, want_bf16=True, eps=1e-5):
    """residual + LN(gelu(xb @ w1.T + b1) @ w2.T + b2)*gamma + beta in ONE kernel (hidden stays in TMEM)."""
    _chk(xb, torch.bfloat16, "x")
    _chk(w1, torch.bfloat16, "w1")
    _chk(w2, torch.float16, "w2 (the GELU output / second GEMM run in fp16)")
    _chk(residual, torch.float32, "residual")
    M, C = xb.shape
    assert w1.shape == (4 * C, C) and w2.shape == (C, 4 * C)
    x_out = torch.empty((M, C), dtype=torch.float32, device=xb.device)
    xo_b = torch.empty((M, C), dtype=torch.bfloat16, device=xb.device) if want_bf16 else None
    _call("mlp_fused_bf16[C=%d]" % C, "pangu_mlp_ln_residual_bf16",
          (_ptr(xb), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(gamma), _ptr(beta), _ptr(residual), _ptr(x_out),
           _ptr(xo_b), M, C, eps, _stream(),),
          flops=16.0 * M * C * C, nbytes=float(M * C * (2 + 4 + 4 + 2 * want_bf16) + 16 * C * C))
    return x_out, xo_b


_SCRATCH = {}         # (device, C) -> fp32 [128 * SMs, C] x1 scratch of attn_proj_mlp_ln_bf16 (29 MB at C = 384: L2-resident)
SCRATCH_ROWS = 128 * 160


def attn_proj_mlp_ln_bf16(o, w_proj, b_proj, gamma1, beta1, x_in, w1, b1, w2, b2, gamma2, beta2, eps1=1e-5, eps2=1e-5,
                          want_bf16=True):
    """x1 = x_in + LN(o @ w_proj.T + b_proj) * gamma1 + beta1;  out = x1 + LN(gelu(x1 @ w1.T + b1) @ w2.T + b2) * gamma2 + beta2
    in ONE kernel (C = 384): linear_ln_residual_bf16 + mlp_ln_residual_bf16 without the HBM round trip of x1."""
    _chk(o, torch.bfloat16, "o")
    _chk(w_proj, torch.bfloat16, "w_proj")
    _chk(w1, torch.bfloat16, "w1")
    _chk(w2, torch.float16, "w2 (the GELU output / second GEMM run in fp16)")
    _chk(x_in, torch.float32, "x_in")
    M, C = o.shape
    assert w_proj.shape == (C, C) and w1.shape == (4 * C, C) and w2.shape == (C, 4 * C) and x_in.shape == (M, C)
    key = (o.device, C)
    scratch = _SCRATCH.get(key)
    if scratch is None:
        scratch = _SCRATCH[key] = torch.empty((SCRATCH_ROWS, C), dtype=torch.float32, device=o.device)
    x_out = torch.empty((M, C), dtype=torch.float32, device=o.device)
    xo_b = torch.empty((M, C), dtype=torch.bfloat16, device=o.device) if want_bf16 else None
    _call("attn_proj_mlp_bf16[C=%d]" % C, "pangu_attn_proj_mlp_bf16",
          (_ptr(o), _ptr(w_proj), _ptr(b_proj), _ptr(gamma1), _ptr(beta1), _ptr(x_in), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2),
           _ptr(gamma2), _ptr(beta2), _ptr(scratch), SCRATCH_ROWS, _ptr(x_out), _ptr(xo_b), M, C, eps1, eps2, _stream(),),
          flops=18.0 * M * C * C, nbytes=float(M * C * (2 + 4 + 4 + 2 * want_bf16) + 18 * C * C))
    return x_out, xo_b


def window_attention(qkv, qkv_bias, earth_bias, Z, H, W, heads, mode):
    """qkv [N, 3C] (token order, or window order when mode == WINDOWED) -> [N, C]."""
    _chk(qkv, name="qkv")
    _chk(qkv_bias, torch.float32, "qkv_bias")
    _chk(earth_bias, name="earth_bias")
    N, C3 = qkv.shape
    C = C3 // 3
    out = torch.empty((N, C), dtype=qkv.dtype, device=qkv.device)
    g = geom(Z, H, W, C, heads)
    nwin = (W // 12) * (Z // 2) * ((H + 5) // 6)
    tag = "attention_%s[C=%d]" % ("bf16" if qkv.dtype == torch.bfloat16 else "f32", C)
    _call(tag, "pangu_window_attention", (_ptr(qkv), _ptr(qkv_bias), _ptr(earth_bias), _DT[earth_bias.dtype],
                                               _ptr(out), g, int(mode), _DT[qkv.dtype], _stream(),),
          flops=nwin * heads * 4.0 * 144 * 144 * 32,
          nbytes=float(qkv.numel() * qkv.element_size() + out.numel() * out.element_size() + earth_bias.numel() * earth_bias.element_size()))
    return out


def full_band(H):
    """The whole grid as one band."""
    return Band(0, H, 0, (H + 5) // 6, 0, 0, 0)


def window_attention_band(qkv, halo_qkv, qkv_bias, earth_bias, Z, H, W, heads, band, roll, halo_lo_qkv=None,
                          return_halo=True, prescaled=False, exact_max=False, halo_kv=False):
    """Band-sharded window attention (bf16): qkv [Z*hrows*W, 3C] holds the band's own rows of the GLOBAL
    (Z, H, W) grid, halo_qkv [Z*halo*W, 3C] the southern neighbour's first rows, halo_lo_qkv the northern
    neighbour's last rows.  -> (out, halo_out): halo_out (attention output of the southern halo rows, to be sent
    back) only when return_halo, else None (the neighbour computes those rows itself).
    halo_kv: the halo tensors hold only the K and V columns, [Z*halo*W, 2C] (PANGU_ATTN_HALO_KV: no halo output possible)."""
    _chk(qkv, torch.bfloat16, "qkv")
    _chk(qkv_bias, torch.float32, "qkv_bias")
    _chk(earth_bias, name="earth_bias")
    C = qkv.shape[1] // 3
    if qkv.shape[0] != Z * band.hrows * W:
        raise abi.PanguError(f"qkv has {qkv.shape[0]} rows, band expects {Z * band.hrows * W}")
    out = torch.empty((qkv.shape[0], C), dtype=qkv.dtype, device=qkv.device)
    halo_out = None
    for t, n, nm in ((halo_qkv, band.halo, "halo_qkv"), (halo_lo_qkv, band.halo_lo, "halo_lo_qkv")):
        if n:
            _chk(t, torch.bfloat16, nm)
            if t.shape[0] != Z * n * W or t.shape[1] != (2 if halo_kv else 3) * C:
                raise abi.PanguError(f"{nm} has the wrong shape {tuple(t.shape)}")
    if halo_kv and return_halo:
        raise abi.PanguError("window_attention_band: K/V-only halos cannot return the halo rows' output")
    if band.halo and return_halo:
        halo_out = torch.empty((Z * band.halo * W, C), dtype=qkv.dtype, device=qkv.device)
    g = geom(Z, H, W, C, heads)
    nwin = (W // 12) * (Z // 2) * band.nhw
    _call("attention_bf16[C=%d]" % C, "pangu_window_attention_band",
          (_ptr(qkv), _ptr(halo_qkv) if band.halo else None, _ptr(halo_lo_qkv) if band.halo_lo else None, _ptr(qkv_bias),
           _ptr(earth_bias), _DT[earth_bias.dtype], _ptr(out), _ptr(halo_out), g, band, int(roll),
           int(bool(prescaled)) | (2 if exact_max else 0) | (4 if halo_kv else 0),
           _stream(),),
          flops=nwin * heads * 4.0 * 144 * 144 * 32,
          nbytes=float(qkv.numel() * 2 + out.numel() * 2 + earth_bias.numel() * earth_bias.element_size() * band.nhw / ((H + 5) // 6)))
    return out, halo_out


# ------------------------------------------------------------------ layout
def patch_embed_gather(inp, inp_s, statistics, maps, const_h, out_dtype):
    """inp [5,13,lat,1440], inp_s [4,lat,1440], maps [3,map_rows,1440], const_h [13,lat,1440] (any leading 1s);
    lat = 721 for the full grid or the pixel rows of a latitude band.  -> (surface patches, upper patches)."""
    sm, ss, um, us = statistics
    dev = inp.device
    for t, n in ((inp, "input"), (inp_s, "input_surface"), (maps, "maps"), (const_h, "const_h")):
        _chk(t, torch.float32, n)
    lat, map_rows = inp.shape[-2], maps.shape[-2]
    tok_rows = (lat + 3) // 4
    if inp_s.shape[-2] != lat or const_h.shape[-2] != lat or inp.shape[-1] != 1440:
        raise abi.PanguError("patch_embed_gather: inconsistent latitude extents")
    sm, ss = sm.reshape(-1).contiguous().float(), ss.reshape(-1).contiguous().float()
    um, us = um.reshape(13, 5).contiguous().float(), us.reshape(13, 5).contiguous().float()
    ps = torch.empty((tok_rows * 360, 112), dtype=out_dtype, device=dev)
    pu = torch.empty((7 * tok_rows * 360, 192), dtype=out_dtype, device=dev)
    _call("patch_embed_gather", "pangu_patch_embed_gather_rows",
          (_ptr(inp), _ptr(inp_s), _ptr(sm), _ptr(ss), _ptr(um), _ptr(us), _ptr(maps), _ptr(const_h), _ptr(ps), _ptr(pu),
           _DT[out_dtype], lat, tok_rows, map_rows, _stream(),), kernels=2,
          nbytes=float((inp.numel() + inp_s.numel() + maps.numel() + const_h.numel()) * 4 + (ps.numel() + pu.numel()) * ps.element_size()))
    return ps, pu


def patch_recover_scatter(y_upper, y_surface, lat=721, denorm=None):
    """-> output [1,5,13,lat,1440], output_surface [1,4,lat,1440]; y_* hold ceil(lat/4) token rows.
    denorm = (surface_mean [4], surface_std [4], upper_mean [5*13], upper_std [5*13]) fp32 device tensors in the
    order of era5_data.utils_data.weatherStatistics_output: outputs are x * std + mean (physical units)."""
    _chk(y_upper, torch.float32, "y_upper")
    _chk(y_surface, torch.float32, "y_surface")
    dev = y_upper.device
    tok_rows = y_surface.shape[0] // 360
    if y_surface.shape[0] != tok_rows * 360 or y_upper.shape[0] != 7 * tok_rows * 360 or (lat + 3) // 4 != tok_rows:
        raise abi.PanguError("patch_recover_scatter: token rows do not match the latitude extent")
    out = torch.empty((1, 5, 13, lat, 1440), dtype=torch.float32, device=dev)
    out_s = torch.empty((1, 4, lat, 1440), dtype=torch.float32, device=dev)
    sm = ss = um = us = None
    if denorm is not None:
        sm, ss, um, us = (_chk(t.reshape(-1), torch.float32, "statistics") for t in denorm)
        if sm.numel() != 4 or ss.numel() != 4 or um.numel() != 65 or us.numel() != 65:
            raise abi.PanguError("patch_recover_scatter: statistics must have 4 / 4 / 65 / 65 elements")
    _call("patch_recover_scatter", "pangu_patch_recover_scatter_denorm",
          (_ptr(y_upper), _ptr(y_surface), _ptr(out), _ptr(out_s), lat, tok_rows, _ptr(us), _ptr(um), _ptr(ss), _ptr(sm),
           _stream(),),
          kernels=2, nbytes=float((y_upper.numel() + y_surface.numel() + out.numel() + out_s.numel()) * 4))
    return out, out_s


def downsample_merge_ln(x, gamma, beta, Z, H, W, out_dtype, eps=1e-5):
    _chk(x, torch.float32, "x")
    C = x.shape[-1]
    rows = Z * ((H + 1) // 2) * (W // 2)
    out = torch.empty((rows, 4 * C), dtype=out_dtype, device=x.device)
    _call("downsample_merge_ln", "pangu_downsample_merge_ln", (_ptr(x), _ptr(gamma), _ptr(beta), _ptr(out), _DT[out_dtype], Z, H, W,
                                                  C, eps, _stream(),),
          nbytes=float(x.numel() * 4 + out.numel() * out.element_size()))
    return out


def upsample_shuffle_ln(y, gamma, beta, Z, H2, W2, H, out_dtype, eps=1e-5):
    _chk(y, name="y")
    Co = y.shape[-1] // 4
    out = torch.empty((Z * H * 2 * W2, Co), dtype=out_dtype, device=y.device)
    _call("upsample_shuffle_ln", "pangu_upsample_shuffle_ln", (_ptr(y), _DT[y.dtype], _ptr(gamma), _ptr(beta), _ptr(out),
                                                  _DT[out_dtype], Z, H2, W2, H, Co, eps, _stream(),),
          nbytes=float(y.numel() * y.element_size() + out.numel() * out.element_size()))
    return out


def cast_bf16(x):
    _chk(x, torch.float32, "x")
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _call("cast_f32_bf16", "pangu_cast_f32_bf16", (_ptr(x), _ptr(out), x.numel(), _stream(),), nbytes=float(x.numel() * 6))
    return out


def concat_cast_bf16(a, b):
    _chk(a, torch.float32, "a")
    _chk(b, torch.float32, "b")
    n, C1 = a.shape
    C2 = b.shape[1]
    out = torch.empty((n, C1 + C2), dtype=torch.bfloat16, device=a.device)
    _call("concat_cast_bf16", "pangu_concat_cast_bf16", (_ptr(a), _ptr(b), _ptr(out), n, C1, C2, _stream(),),
          nbytes=float(n * (C1 + C2) * 6))
    return out


# ------------------------------------------------------------------ fine-tune backward (bf16 operands, fp32 accumulation)
def linear_add(a, w, bias=None, addend=None, out=None):
    """out (fp32) = a @ w.T + bias + addend -- dgrad GEMM fused with the residual-gradient add."""
    _chk(a, torch.bfloat16, "a")
    _chk(w, torch.bfloat16, "w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise abi.PanguError("linear_add: a and w disagree on K")
    if addend is not None:
        _chk(addend, torch.float32, "addend")
        if addend.shape != (M, N):
            raise abi.PanguError("linear_add: addend must be [M, N]")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _call("gemm_bf16_dgrad[K=%d,N=%d]" % (K, N), "pangu_linear_bf16_add",
          (_ptr(a), K, _ptr(w), _ptr(bias), _ptr(addend), _ptr(out), N, M, K, N, _stream(),), flops=2.0 * M * K * N,
          nbytes=float(M * K * 2 + N * K * 2 + M * N * (4 + 4 * (addend is not None))))
    return out


def linear_wgrad(dy, x, dw=None):
    """dw [n_out, k_in] (fp32) += dy[M, n_out].T @ x[M, k_in]; dy / x bf16 row-major (row slices allowed)."""
    for t, n in ((dy, "dy"), (x, "x")):
        if not t.is_cuda or t.dtype != torch.bfloat16 or t.stride(-1) != 1:
            raise abi.PanguError(f"linear_wgrad: {n} must be a CUDA bf16 tensor with unit inner stride")
    M, n_out = dy.shape
    k_in = x.shape[1]
    if x.shape[0] != M:
        raise abi.PanguError("linear_wgrad: dy and x disagree on the token count")
    if dw is None:
        dw = torch.zeros((n_out, k_in), dtype=torch.float32, device=dy.device)
    if dw.dtype != torch.float32 or dw.stride(-1) != 1 or dw.shape != (n_out, k_in):
        raise abi.PanguError("linear_wgrad: dw must be fp32 [n_out, k_in] with unit inner stride")
    _call("wgrad_bf16[N=%d,K=%d]" % (n_out, k_in), "pangu_linear_wgrad_bf16",
          (_ptr(dy), dy.stride(0), _ptr(x), x.stride(0), _ptr(dw), dw.stride(0), M, n_out, k_in, _stream(),),
          flops=2.0 * M * n_out * k_in, nbytes=float(M * (n_out + k_in) * 2))
    return dw


def colsum(x, out=None):
    """out [C] (fp32) += x.sum(0); x bf16 or fp32 [M, C]."""
    if not x.is_cuda or x.stride(-1) != 1:
        raise abi.PanguError("colsum: x must be a CUDA tensor with unit inner stride")
    M, C = x.shape
    if out is None:
        out = torch.zeros((C,), dtype=torch.float32, device=x.device)
    _call("colsum", "pangu_colsum", (_ptr(x), _DT[x.dtype], x.stride(0), M, C, _ptr(out), _stream(),),
          nbytes=float(M * C * x.element_size()))
    return out


def ln_backward(dout, y, gamma, scale=1.0, dout2=None, dgamma=None, dbeta=None, dcolsum=None, eps=1e-5):
    """Backward of scale * (LN(y) * gamma + beta) -> dy bf16; dgamma / dbeta / dcolsum (fp32 [C]) are accumulated."""
    _chk(dout, torch.float32, "dout")
    _chk(y, name="y")
    if dout2 is not None:
        _chk(dout2, torch.float32, "dout2")
    M, C = y.shape
    dy = torch.empty((M, C), dtype=torch.bfloat16, device=y.device)
    _call("ln_backward[C=%d]" % C, "pangu_ln_backward",
          (_ptr(dout), _ptr(dout2), _ptr(y), _DT[y.dtype], _ptr(gamma), float(scale), _ptr(dy), _ptr(dgamma), _ptr(dbeta),
           _ptr(dcolsum), M, C, eps, _stream(),),
          nbytes=float(M * C * (4 + 4 * (dout2 is not None) + y.element_size() + 2)))
    return dy


def upsample_shuffle_ln_backward(dout, y, gamma, dgamma, dbeta, Z, H2, W2, H, eps=1e-5):
    _chk(dout, torch.float32, "dout")
    _chk(y, torch.bfloat16, "y")
    Co = y.shape[-1] // 4
    dy = torch.zeros_like(y)                              # the cropped row gets no gradient
    _call("upsample_shuffle_ln_backward", "pangu_upsample_shuffle_ln_backward",
          (_ptr(dout), _ptr(y), _ptr(gamma), _ptr(dy), _ptr(dgamma), _ptr(dbeta), Z, H2, W2, H, Co, eps, _stream(),),
          nbytes=float(dout.numel() * 4 + y.numel() * 4))
    return dy


def downsample_merge_ln_backward(dout, x, gamma, dgamma, dbeta, Z, H, W, eps=1e-5):
    _chk(dout, torch.float32, "dout")
    _chk(x, torch.float32, "x")
    C = x.shape[-1]
    dx = torch.empty_like(x)
    _call("downsample_merge_ln_backward", "pangu_downsample_merge_ln_backward",
          (_ptr(dout), _ptr(x), _ptr(gamma), _ptr(dx), _ptr(dgamma), _ptr(dbeta), Z, H, W, C, eps, _stream(),),
          nbytes=float(dout.numel() * 4 + x.numel() * 8))
    return dx


def gelu_bf16(h_pre):
    _chk(h_pre, torch.bfloat16, "h_pre")
    h = torch.empty_like(h_pre)
    _call("gelu_bf16", "pangu_gelu_bf16", (_ptr(h_pre), _ptr(h), h_pre.numel(), _stream(),), nbytes=float(h_pre.numel() * 4))
    return h


def gelu_backward_bf16(dh, h_pre, dcolsum=None):
    """dh <- dh * GELU'(h_pre) in place; dcolsum [F] += column sums of the result."""
    _chk(dh, torch.bfloat16, "dh")
    _chk(h_pre, torch.bfloat16, "h_pre")
    M, F = dh.shape
    _call("gelu_backward_bf16", "pangu_gelu_backward_bf16", (_ptr(dh), _ptr(h_pre), _ptr(dh), _ptr(dcolsum), M, F, _stream(),),
          nbytes=float(dh.numel() * 6))
    return dh


def pair_gemm_covers(M, K, N):
    """Shapes of the A-resident CTA-pair GEMM (csrc/tc_gemm2.cu), the kernel that carries the aux epilogues."""
    return K in (192, 384) and N % 192 == 0 and N <= 1536 and M >= 2048


def linear_gelu_pre(a, w, bias):
    """Mlp.linear1 for the fine-tune step: (h_pre, h) = (a @ w.T + bias, GELU(h_pre)), both bf16, from ONE GEMM pass
    (PANGU_AUX_PRE_OUT); shapes the CTA-pair GEMM does not cover run the GEMM twice (still CUDA)."""
    _chk(a, torch.bfloat16, "a")
    _chk(w, torch.bfloat16, "w")
    _chk(bias, torch.float32, "bias")
    M, K = a.shape
    N = w.shape[0]
    if not pair_gemm_covers(M, K, N):
        return linear(a, w, bias), linear(a, w, bias, act=ACT_GELU)
    h_pre = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    h = torch.empty_like(h_pre)
    _call("gemm_bf16[K=%d,N=%d,gelu+pre]" % (K, N), "pangu_linear_bf16_aux",
          (_ptr(a), K, _ptr(w), _ptr(bias), _ptr(h), _ptr(h_pre), N, M, K, N, ACT_GELU, 1, None, _stream(),),
          flops=2.0 * M * K * N, nbytes=float(a.numel() * 2 + 4 * M * N + w.numel() * 2))
    return h_pre, h


def linear_gelu_backward(dy, w_t, h_pre, dcolsum=None):
    """dh_pre = (dy @ w_t.T) * GELU'(h_pre) (bf16) and dcolsum [F] += its column sums: the dgrad of Mlp.linear2 with the
    GELU backward in its epilogue (PANGU_AUX_GELU_BWD); other shapes: GEMM + gelu_backward_bf16."""
    _chk(dy, torch.bfloat16, "dy")
    _chk(w_t, torch.bfloat16, "w_t")
    _chk(h_pre, torch.bfloat16, "h_pre")
    M, K = dy.shape
    N = w_t.shape[0]
    if not pair_gemm_covers(M, K, N):
        return gelu_backward_bf16(linear(dy, w_t, None), h_pre, dcolsum)
    dh = torch.empty((M, N), dtype=torch.bfloat16, device=dy.device)
    _call("gemm_bf16[K=%d,N=%d,gelu_bwd]" % (K, N), "pangu_linear_bf16_aux",
          (_ptr(dy), K, _ptr(w_t), None, _ptr(dh), _ptr(h_pre), N, M, K, N, ACT_NONE, 2, _ptr(dcolsum), _stream(),),
          flops=2.0 * M * K * N, nbytes=float(dy.numel() * 2 + 4 * M * N + w_t.numel() * 2))
    return dh


def window_attention_train(qkv, qkv_bias, earth_bias, Z, H, W, heads, roll, exact_max=False):
    """Pre-scaled bf16 window attention that also returns the log2-sum-exp rows for the backward kernel."""
    _chk(qkv, torch.bfloat16, "qkv")
    _chk(qkv_bias, torch.float32, "qkv_bias")
    _chk(earth_bias, torch.bfloat16, "earth_bias")
    N, C3 = qkv.shape
    C = C3 // 3
    nLon, T = window_counts(Z, H, W)
    out = torch.empty((N, C), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((nLon, T, heads, 144), dtype=torch.float32, device=qkv.device)
    g = geom(Z, H, W, C, heads)
    _call("attention_bf16[C=%d]" % C, "pangu_window_attention_train",
          (_ptr(qkv), _ptr(qkv_bias), _ptr(earth_bias), _ptr(out), _ptr(lse), g, int(roll) | (0x100 if exact_max else 0), _stream(),),
          flops=nLon * T * heads * 4.0 * 144 * 144 * 32, nbytes=float(qkv.numel() * 2 + out.numel() * 2 + earth_bias.numel() * 2))
    return out, lse


def window_attention_backward(qkv, qkv_bias, earth_bias, out, d_out, lse, Z, H, W, heads, roll, d_earth_bias, d_pad):
    """-> d_qkv bf16 [N, 3C]; d_earth_bias fp32 [T, heads, 144, 144] and d_pad fp32 [3C] are accumulated."""
    for t, n in ((qkv, "qkv"), (earth_bias, "earth_bias"), (out, "out"), (d_out, "d_out")):
        _chk(t, torch.bfloat16, n)
    _chk(lse, torch.float32, "lse")
    _chk(d_earth_bias, torch.float32, "d_earth_bias")
    _chk(d_pad, torch.float32, "d_pad")
    N, C3 = qkv.shape
    C = C3 // 3
    nLon, T = window_counts(Z, H, W)
    d_qkv = torch.empty_like(qkv)
    g = geom(Z, H, W, C, heads)
    _call("attention_bwd_bf16[C=%d]" % C, "pangu_window_attention_backward",
          (_ptr(qkv), _ptr(qkv_bias), _ptr(earth_bias), _ptr(out), _ptr(d_out), _ptr(lse), _ptr(d_qkv), _ptr(d_earth_bias),
           _ptr(d_pad), g, int(roll), _stream(),),
          flops=nLon * T * heads * 14.0 * 144 * 144 * 32, nbytes=float(qkv.numel() * 4 + out.numel() * 4 + earth_bias.numel() * 6))
    return d_qkv


def patch_recover_gather_backward(d_output, d_output_surface, lat=721):
    _chk(d_output, torch.float32, "d_output")
    _chk(d_output_surface, torch.float32, "d_output_surface")
    tok_rows = (lat + 3) // 4
    dev = d_output.device
    dyu = torch.empty((7 * tok_rows * 360, 160), dtype=torch.bfloat16, device=dev)
    dys = torch.empty((tok_rows * 360, 64), dtype=torch.bfloat16, device=dev)
    _call("patch_recover_gather_backward", "pangu_patch_recover_gather_backward",
          (_ptr(d_output), _ptr(d_output_surface), _ptr(dyu), _ptr(dys), lat, tok_rows, _stream(),), kernels=2,
          nbytes=float((d_output.numel() + d_output_surface.numel()) * 4 + (dyu.numel() + dys.numel()) * 2))
    return dyu, dys
