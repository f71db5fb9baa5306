// Bandwidth kernels of the fine-tune backward pass (the autograd of models/layers.py that
// finetune/finetune_fully.py runs through loss.backward(), models/pangu_sample.py:226):
// LayerNorm backward (plain, and fused with the up-sample pixel-shuffle / down-sample merge index maps),
// exact-erf GELU forward/backward, column sums (bias gradients), the inverse of the patch-recover scatter.
// All are HBM-bound: one warp per token row with 128-byte coalesced accesses, per-lane column partial sums
// kept in registers across a grid-stride row loop and flushed with one atomic per column per warp.
#include "common.cuh"

namespace pangu {

constexpr int kLat = 721, kLon = 1440, kLev = 13;
constexpr int kTokW = 360;
constexpr int kTileTok = 32;

__device__ __forceinline__ float gelu_erf_fwd(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// d/dx [x Phi(x)] = Phi(x) + x phi(x)   (nn.GELU(), models/layers.py:313)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return fmaf(x, pdf, cdf);
}

enum { LNB_PLAIN = 0, LNB_UP = 1, LNB_DOWN = 2 };
struct LnBwdMap { int Z, H, W, H2, W2, C; };

// out = scale * (LN(y) * gamma + beta)  ->  dy, dgamma += scale * sum dout * yhat, dbeta += scale * sum dout,
// dcolsum += sum dy (the bias gradient of the linear that produced y).  dout = dout_a (+ dout_b).
//   LNB_PLAIN: y, dy [M, F] row-major.
//   LNB_UP   : rows are fine-grid tokens (z,h,w) of [Z,H,2*W2]; y / dy live in the coarse tensor [Z*H2*W2, 4F] at
//              feature block (h&1)*2+(w&1) (models/layers.py:546-556); dy of the cropped row is pre-zeroed by the caller.
//   LNB_DOWN : rows are coarse tokens (z,h2,w2); feature f = q*C + c gathers fine token (2h2+(q>>1), 2w2+(q&1)), zero
//              when that row is the pad row (models/layers.py:506-516); dy is the gradient of the fine tensor [Z*H*W, C].
template <int MODE, typename TY, typename TD, int kPerLane>
__global__ void __launch_bounds__(256)
ln_backward_kernel(const float* __restrict__ dout_a, const float* __restrict__ dout_b, const TY* __restrict__ y,
                   const float* __restrict__ gamma, float scale, TD* __restrict__ dy, float* __restrict__ dgamma,
                   float* __restrict__ dbeta, float* __restrict__ dcolsum, long long M, float eps, LnBwdMap mp) {
  constexpr int F = kPerLane * 32;
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  float g_acc[kPerLane], b_acc[kPerLane], c_acc[kPerLane], gam[kPerLane];
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) { g_acc[i] = b_acc[i] = c_acc[i] = 0.f; gam[i] = __ldg(gamma + i * 32 + lane) * scale; }

  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += nwarps) {
    long long base = row * F;                       // PLAIN
    int z = 0, hh = 0, ww = 0;
    if (MODE == LNB_UP) {
      const int W = 2 * mp.W2;
      const int w = (int)(row % W);
      const long long zh = row / W;
      const int h = (int)(zh % mp.H);
      z = (int)(zh / mp.H);
      base = (((long long)z * mp.H2 + (h >> 1)) * mp.W2 + (w >> 1)) * (4 * F) + ((h & 1) * 2 + (w & 1)) * F;
    } else if (MODE == LNB_DOWN) {
      ww = (int)(row % mp.W2);
      const long long zh = row / mp.W2;
      hh = (int)(zh % mp.H2);
      z = (int)(zh / mp.H2);
    }
    float v[kPerLane], d[kPerLane];
    int addr[kPerLane];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int f = i * 32 + lane;
      if (MODE == LNB_DOWN) {
        const int q = f / mp.C, c = f - q * mp.C;
        const int h = 2 * hh + (q >> 1), w = 2 * ww + (q & 1);
        addr[i] = h < mp.H ? (int)((((long long)z * mp.H + h) * mp.W + w) * mp.C + c) : -1;
      } else {
        addr[i] = (int)(base + f);
      }
      v[i] = addr[i] >= 0 ? to_f32<TY>(y[addr[i]]) : 0.f;
      s += v[i];
      float dd = __ldg(dout_a + row * F + f);
      if (dout_b != nullptr) dd += __ldg(dout_b + row * F + f);
      d[i] = dd;
    }
    const float mean = warp_sum(s) * (1.0f / F);
    float qv = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) { v[i] -= mean; qv = fmaf(v[i], v[i], qv); }
    const float rstd = rsqrtf(warp_sum(qv) * (1.0f / F) + eps);
    float sg = 0.f, sgy = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      v[i] *= rstd;                                  // yhat
      g_acc[i] = fmaf(d[i], v[i], g_acc[i]);
      b_acc[i] += d[i];
      d[i] *= gam[i];                                // g = dout * scale * gamma
      sg += d[i];
      sgy = fmaf(d[i], v[i], sgy);
    }
    sg = warp_sum(sg) * (1.0f / F);
    sgy = warp_sum(sgy) * (1.0f / F);
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const float r = rstd * (d[i] - sg - v[i] * sgy);
      if (MODE == LNB_PLAIN) c_acc[i] += r;
      if (addr[i] >= 0) dy[addr[i]] = from_f32<TD>(r);
    }
  }
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    const int f = i * 32 + lane;
    if (dgamma != nullptr) atomicAdd(dgamma + f, g_acc[i] * scale);
    if (dbeta != nullptr) atomicAdd(dbeta + f, b_acc[i] * scale);
    if (MODE == LNB_PLAIN && dcolsum != nullptr) atomicAdd(dcolsum + f, c_acc[i]);
  }
}

// h = GELU(h_pre) (bf16, exact erf form), 8 elements per thread.
__global__ void __launch_bounds__(256)
gelu_fwd_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  uint4 v = __ldg(in + i);
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(p[j]);
    p[j] = __floats2bfloat162_rn(gelu_erf_fwd(f.x), gelu_erf_fwd(f.y));
  }
  out[i] = v;
}

// dh_pre = dh * GELU'(h_pre) (bf16, may alias dh) and dcolsum[f] += sum_rows dh_pre (bias gradient of Mlp.linear1).
// Thread = 8 consecutive columns; a CTA covers blockDim/(F/8) rows per step of a grid-stride loop.
__global__ void __launch_bounds__(192)
gelu_bwd_bf16_kernel(const uint4* __restrict__ dh, const uint4* __restrict__ h_pre, uint4* __restrict__ out,
                     float* __restrict__ dcolsum, long long M, int F) {
  const int cg = F >> 3;
  const int c = threadIdx.x % cg, r0 = threadIdx.x / cg, rpb = blockDim.x / cg;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (r0 < rpb) {
    for (long long row = (long long)blockIdx.x * rpb + r0; row < M; row += (long long)gridDim.x * rpb) {
      uint4 g = __ldg(dh + row * cg + c);
      const uint4 x = __ldg(h_pre + row * cg + c);
      __nv_bfloat162* gp = reinterpret_cast<__nv_bfloat162*>(&g);
      const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 gf = __bfloat1622float2(gp[j]), xf = __bfloat1622float2(xp[j]);
        const __nv_bfloat162 o = __floats2bfloat162_rn(gf.x * gelu_erf_grad(xf.x), gf.y * gelu_erf_grad(xf.y));
        gp[j] = o;
        const float2 of = __bfloat1622float2(o);     // sum what the wgrad GEMM will see
        acc[2 * j] += of.x; acc[2 * j + 1] += of.y;
      }
      out[row * cg + c] = g;
    }
    if (dcolsum != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(dcolsum + c * 8 + j, acc[j]);
    }
  }
}

// out[c] += sum_rows x[row, c]; grid (row chunks, ceil(C/64)); a warp reads 64 columns (2 per lane) of one row.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long ld, long long M, int C, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.y * 64 + 2 * lane;
  if (col >= C) return;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  float a0 = 0.f, a1 = 0.f;
  constexpr int U = 8;                               // rows in flight per warp (one 4- / 8-byte load per lane each)
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += nwarps * U) {
    float2 f[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = row + u * nwarps;
      f[u] = make_float2(0.f, 0.f);
      if (r < M) {
        const T* p = x + r * ld + col;
        if (sizeof(T) == 2) f[u] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
        else f[u] = *reinterpret_cast<const float2*>(p);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { a0 += f[u].x; a1 += f[u].y; }
  }
  atomicAdd(out + col, a0);
  atomicAdd(out + col + 1, a1);
}

// Inverse of patch_recover_kernel (layout_kernels.cu): gradient of the output fields -> gradient of the conv outputs
// y [tokens, kF] as bf16 (the A operand of the dgrad / wgrad GEMMs).  Cropped positions (level 13, rows >= lat;
// models/layers.py:603,619) get zero.
template <int kF, int kRows, bool kUpper>
__global__ void __launch_bounds__(256)
patch_recover_bwd_kernel(const float* __restrict__ dout, __nv_bfloat16* __restrict__ dy, int lat, int tokH) {
  constexpr int PITCH = kF + 1;
  __shared__ float tile[kTileTok * PITCH];
  const int w0 = blockIdx.x * kTileTok, hp = blockIdx.y, zp = blockIdx.z;
  const int ntok = min(kTileTok, kTokW - w0);
  for (int i = threadIdx.x; i < kRows * kTileTok; i += blockDim.x) {
    const int tok = i & (kTileTok - 1);
    const int r = i / kTileTok;
    int v, lev, yy;
    if (kUpper) { v = r >> 3; lev = 2 * zp + ((r >> 2) & 1); yy = 4 * hp + (r & 3); }
    else { v = r >> 2; lev = 0; yy = 4 * hp + (r & 3); }
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tok < ntok && lev < kLev && yy < lat) {
      const long long plane = kUpper ? ((long long)v * kLev + lev) : (long long)v;
      o = __ldg(reinterpret_cast<const float4*>(dout + (plane * lat + yy) * kLon + 4 * (w0 + tok)));
    }
    float* s = tile + tok * PITCH + r * 4;
    s[0] = o.x; s[1] = o.y; s[2] = o.z; s[3] = o.w;
  }
  __syncthreads();
  const long long tok0 = ((long long)zp * tokH + hp) * kTokW + w0;
  __nv_bfloat162* d2 = reinterpret_cast<__nv_bfloat162*>(dy + tok0 * kF);
  constexpr int F2 = kF / 2;
  for (int i = threadIdx.x; i < ntok * F2; i += blockDim.x) {
    const int r = i / F2, c = (i - r * F2) * 2;
    d2[i] = __floats2bfloat162_rn(tile[r * PITCH + c], tile[r * PITCH + c + 1]);
  }
}


// Training loss of the reference in one pass (models/pangu_sample.py:163-218, default branch): the target is normalised
// (era5_data/utils_data.py normData: (x - mean) / std per variable and level), then
//   loss += scale * sum_i w[var(i)] * |out_i - target_i|     and     d_out_i = scale * w[var(i)] * sign(out_i - target_i)
// with scale = loss_weight / numel, i.e. loss_weight * mean(L1(out, target) * w) and its gradient for d loss = 1.
__global__ void __launch_bounds__(256)
weighted_l1_loss_kernel(const float4* __restrict__ out, const float4* __restrict__ target, const float* __restrict__ mean,
                        const float* __restrict__ stdv, const float* __restrict__ weight, const float4* __restrict__ mask,
                        int planes_per_var, long long plane_elems4, long long total4, float scale, float* __restrict__ loss_sum,
                        float4* __restrict__ d_out) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int plane = (int)(i / plane_elems4);
    const float w = __ldg(weight + plane / planes_per_var) * scale;
    float m = 0.f, rs = 1.f;
    if (mean != nullptr) { m = __ldg(mean + plane); rs = 1.0f / __ldg(stdv + plane); }
    const float4 o = __ldg(out + i), t = __ldg(target + i);
    // custom mask [H][W] (models/pangu_sample.py:196-199): the L1 term of every plane is multiplied by it
    float4 k = make_float4(w, w, w, w);
    if (mask != nullptr) { const float4 mk = __ldg(mask + (i - (long long)plane * plane_elems4)); k.x *= mk.x; k.y *= mk.y; k.z *= mk.z; k.w *= mk.w; }
    const float d0 = o.x - (t.x - m) * rs, d1 = o.y - (t.y - m) * rs, d2 = o.z - (t.z - m) * rs, d3 = o.w - (t.w - m) * rs;
    acc += k.x * fabsf(d0) + k.y * fabsf(d1) + k.z * fabsf(d2) + k.w * fabsf(d3);
    if (d_out != nullptr)
      d_out[i] = make_float4(d0 > 0.f ? k.x : (d0 < 0.f ? -k.x : 0.f), d1 > 0.f ? k.y : (d1 < 0.f ? -k.y : 0.f),
                             d2 > 0.f ? k.z : (d2 < 0.f ? -k.z : 0.f), d3 > 0.f ? k.w : (d3 < 0.f ? -k.w : 0.f));
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = part[threadIdx.x];
    v += __shfl_xor_sync(0xffu, v, 4); v += __shfl_xor_sync(0xffu, v, 2); v += __shfl_xor_sync(0xffu, v, 1);
    if (threadIdx.x == 0) atomicAdd(loss_sum, v);
  }
}

// Wind-speed L1 loss and its gradient in one pass (models/pangu_sample.py:74-93 get_wind_speed, :184-193):
//   ws(u, v) = sqrt(u^2 + v^2);  loss += scale * sum mask * | ws(out_u, out_v) - ws(norm(tgt_u), norm(tgt_v)) |
//   d_u = scale * mask * sign(.) * out_u / ws(out),  d_v likewise  (0 where ws(out) == 0).
// u / v are plane-aligned: plane p of u and plane p of v are the same level ([planes][plane_elems] each).
__global__ void __launch_bounds__(256)
wind_speed_l1_loss_kernel(const float4* __restrict__ out_u, const float4* __restrict__ out_v, const float4* __restrict__ tgt_u,
                          const float4* __restrict__ tgt_v, const float* __restrict__ mean_u, const float* __restrict__ std_u,
                          const float* __restrict__ mean_v, const float* __restrict__ std_v, const float4* __restrict__ mask,
                          long long plane_elems4, long long total4, float scale, float* __restrict__ loss_sum,
                          float4* __restrict__ d_u, float4* __restrict__ d_v) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int plane = (int)(i / plane_elems4);
    float mu = 0.f, ru = 1.f, mv = 0.f, rv = 1.f;
    if (mean_u != nullptr) {
      mu = __ldg(mean_u + plane); ru = 1.0f / __ldg(std_u + plane);
      mv = __ldg(mean_v + plane); rv = 1.0f / __ldg(std_v + plane);
    }
    const float4 ou = __ldg(out_u + i), ov = __ldg(out_v + i), tu = __ldg(tgt_u + i), tv = __ldg(tgt_v + i);
    float4 k = make_float4(scale, scale, scale, scale);
    if (mask != nullptr) { const float4 mk = __ldg(mask + (i - (long long)plane * plane_elems4)); k.x *= mk.x; k.y *= mk.y; k.z *= mk.z; k.w *= mk.w; }
    const float a[4] = {ou.x, ou.y, ou.z, ou.w}, b[4] = {ov.x, ov.y, ov.z, ov.w};
    const float c[4] = {(tu.x - mu) * ru, (tu.y - mu) * ru, (tu.z - mu) * ru, (tu.w - mu) * ru};
    const float d[4] = {(tv.x - mv) * rv, (tv.y - mv) * rv, (tv.z - mv) * rv, (tv.w - mv) * rv};
    const float kk[4] = {k.x, k.y, k.z, k.w};
    float gu[4], gv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float wo = sqrtf(a[j] * a[j] + b[j] * b[j]), wt = sqrtf(c[j] * c[j] + d[j] * d[j]);
      const float diff = wo - wt;
      acc += kk[j] * fabsf(diff);
      const float s = diff > 0.f ? kk[j] : (diff < 0.f ? -kk[j] : 0.f);
      const float inv = wo > 0.f ? 1.0f / wo : 0.f;
      gu[j] = s * a[j] * inv;
      gv[j] = s * b[j] * inv;
    }
    if (d_u != nullptr) {
      d_u[i] = make_float4(gu[0], gu[1], gu[2], gu[3]);
      d_v[i] = make_float4(gv[0], gv[1], gv[2], gv[3]);
    }
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = part[threadIdx.x];
    v += __shfl_xor_sync(0xffu, v, 4); v += __shfl_xor_sync(0xffu, v, 2); v += __shfl_xor_sync(0xffu, v, 1);
    if (threadIdx.x == 0) atomicAdd(loss_sum, v);
  }
}


// Latitude-weighted verification scores in one pass over prediction and target (era5_data/score.py:126-161
// weighted_rmse_torch_channels incl. its optional mask, :181-201 weighted_acc_torch_channels; the reference makes ~10
// full-tensor passes per score and calls them once per variable, models/pangu_sample.py:531-569).  For every plane p
// (one variable at one level, [H][W]) five sums, with w = latitude weight of the row, m = mask (or 1), a = pred - clim[p],
// b = target - clim[p]:   { sum w m (pred-target)^2,  sum w m,  sum w a b,  sum w a^2,  sum w b^2 }.
// Grid (planes, row chunks); fp32 partial sums per thread (a few hundred terms), fp64 from the warp reduction on.
template <int VEC>
__global__ void __launch_bounds__(256)
lat_weighted_score_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ mask,
                          const float* __restrict__ clim, const float* __restrict__ lat_w, int H, int W,
                          double* __restrict__ sums) {
  const int plane = blockIdx.x;
  const int rows_per = (H + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(H, r0 + rows_per);
  const float c = clim != nullptr ? __ldg(clim + plane) : 0.f;
  const float* pp = pred + (long long)plane * H * W;
  const float* tp = target + (long long)plane * H * W;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = r0; r < r1; ++r) {
    const float w = __ldg(lat_w + r);
    float e2 = 0.f, wm = 0.f, ab = 0.f, aa = 0.f, bb = 0.f;
    for (int x = threadIdx.x * VEC; x < W; x += blockDim.x * VEC) {
      float pv[VEC], tv[VEC], mv[VEC];
      if constexpr (VEC == 4) {                                 // W % 4 == 0 and 16-byte aligned bases (checked by the host)
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(pp + (long long)r * W + x));
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(tp + (long long)r * W + x));
        pv[0] = p4.x; pv[1] = p4.y; pv[2] = p4.z; pv[3] = p4.w;
        tv[0] = t4.x; tv[1] = t4.y; tv[2] = t4.z; tv[3] = t4.w;
        if (mask != nullptr) {
          const float4 m4 = __ldg(reinterpret_cast<const float4*>(mask + (long long)r * W + x));
          mv[0] = m4.x; mv[1] = m4.y; mv[2] = m4.z; mv[3] = m4.w;
        } else { mv[0] = mv[1] = mv[2] = mv[3] = 1.f; }
      } else {
        pv[0] = __ldg(pp + (long long)r * W + x); tv[0] = __ldg(tp + (long long)r * W + x);
        mv[0] = mask != nullptr ? __ldg(mask + (long long)r * W + x) : 1.f;
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float d = pv[i] - tv[i], a = pv[i] - c, b = tv[i] - c;
        e2 = fmaf(mv[i] * d, d, e2); wm += mv[i];
        ab = fmaf(a, b, ab); aa = fmaf(a, a, aa); bb = fmaf(b, b, bb);
      }
    }
    acc[0] = fmaf(w, e2, acc[0]); acc[1] = fmaf(w, wm, acc[1]);
    acc[2] = fmaf(w, ab, acc[2]); acc[3] = fmaf(w, aa, acc[3]); acc[4] = fmaf(w, bb, acc[4]);
  }
  __shared__ double part[8][5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    double v = (double)acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0.0;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) v += part[wq][threadIdx.x];
    atomicAdd(sums + plane * 5 + threadIdx.x, v);
  }
}

static unsigned row_grid(long long M, int warps_per_cta) {
  long long want = (M + warps_per_cta - 1) / warps_per_cta;
  const long long cap = 148LL * 8;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace pangu

using namespace pangu;

extern "C" int pangu_ln_backward(const float* dout, const float* dout2, const void* y, int y_dtype, const float* gamma,
                                 float scale, void* dy, float* dgamma, float* dbeta, float* dcolsum, int64_t M,
                                 int32_t C, float eps, void* stream) {
  if (!dout || !y || !gamma || !dy || M < 0) { set_error("ln_backward: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (M == 0) return PANGU_OK;
  cudaStream_t st = as_stream(stream);
  const unsigned grid = row_grid(M, 8);
  const LnBwdMap mp{};
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dy);
  if (y_dtype == PANGU_F32 && C == 192)
    ln_backward_kernel<LNB_PLAIN, float, __nv_bfloat16, 6><<<grid, 256, 0, st>>>(dout, dout2, (const float*)y, gamma, scale, o, dgamma, dbeta, dcolsum, M, eps, mp);
  else if (y_dtype == PANGU_F32 && C == 384)
    ln_backward_kernel<LNB_PLAIN, float, __nv_bfloat16, 12><<<grid, 256, 0, st>>>(dout, dout2, (const float*)y, gamma, scale, o, dgamma, dbeta, dcolsum, M, eps, mp);
  else if (y_dtype == PANGU_BF16 && C == 192)
    ln_backward_kernel<LNB_PLAIN, __nv_bfloat16, __nv_bfloat16, 6><<<grid, 256, 0, st>>>(dout, dout2, (const __nv_bfloat16*)y, gamma, scale, o, dgamma, dbeta, dcolsum, M, eps, mp);
  else if (y_dtype == PANGU_BF16 && C == 384)
    ln_backward_kernel<LNB_PLAIN, __nv_bfloat16, __nv_bfloat16, 12><<<grid, 256, 0, st>>>(dout, dout2, (const __nv_bfloat16*)y, gamma, scale, o, dgamma, dbeta, dcolsum, M, eps, mp);
  else { set_error("ln_backward: C=%d / dtype %d unsupported", C, y_dtype); return PANGU_ERR_UNSUPPORTED; }
  return check_launch("ln_backward");
}

extern "C" int pangu_upsample_shuffle_ln_backward(const float* dout, const void* y, const float* gamma, void* dy,
                                                  float* dgamma, float* dbeta, int32_t Z, int32_t H2, int32_t W2,
                                                  int32_t H, int32_t Cout, float eps, void* stream) {
  if (!dout || !y || !gamma || !dy || Z <= 0 || H2 <= 0 || W2 <= 0 || H > 2 * H2) { set_error("upsample_shuffle_ln_backward: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (Cout != 192) { set_error("upsample_shuffle_ln_backward: Cout=%d unsupported (192)", Cout); return PANGU_ERR_UNSUPPORTED; }
  const long long rows = (long long)Z * H * 2 * W2;
  const LnBwdMap mp{Z, H, 2 * W2, H2, W2, Cout};
  ln_backward_kernel<LNB_UP, __nv_bfloat16, __nv_bfloat16, 6><<<row_grid(rows, 8), 256, 0, as_stream(stream)>>>(
      dout, nullptr, (const __nv_bfloat16*)y, gamma, 1.0f, (__nv_bfloat16*)dy, dgamma, dbeta, nullptr, rows, eps, mp);
  return check_launch("upsample_shuffle_ln_backward");
}

extern "C" int pangu_downsample_merge_ln_backward(const float* dout, const float* x, const float* gamma, float* dx,
                                                  float* dgamma, float* dbeta, int32_t Z, int32_t H, int32_t W,
                                                  int32_t C, float eps, void* stream) {
  if (!dout || !x || !gamma || !dx || Z <= 0 || H <= 0 || W <= 0 || (W & 1)) { set_error("downsample_merge_ln_backward: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (C != 192) { set_error("downsample_merge_ln_backward: C=%d unsupported (192)", C); return PANGU_ERR_UNSUPPORTED; }
  const int H2 = (H + 1) / 2, W2 = W / 2;
  const long long rows = (long long)Z * H2 * W2;
  const LnBwdMap mp{Z, H, W, H2, W2, C};
  ln_backward_kernel<LNB_DOWN, float, float, 24><<<row_grid(rows, 8), 256, 0, as_stream(stream)>>>(
      dout, nullptr, x, gamma, 1.0f, dx, dgamma, dbeta, nullptr, rows, eps, mp);
  return check_launch("downsample_merge_ln_backward");
}

extern "C" int pangu_gelu_bf16(const void* h_pre, void* h, int64_t n, void* stream) {
  if (!h_pre || !h || n < 0 || (n & 7)) { set_error("gelu_bf16: n must be a multiple of 8"); return PANGU_ERR_BAD_ARG; }
  if (n == 0) return PANGU_OK;
  const long long n8 = n / 8;
  gelu_fwd_bf16_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, as_stream(stream)>>>((const uint4*)h_pre, (uint4*)h, n8);
  return check_launch("gelu_bf16");
}

extern "C" int pangu_gelu_backward_bf16(const void* dh, const void* h_pre, void* dh_pre, float* dcolsum, int64_t M,
                                        int32_t F, void* stream) {
  if (!dh || !h_pre || !dh_pre || M < 0 || F <= 0 || (F & 7) || F / 8 > 192) { set_error("gelu_backward_bf16: bad argument (F=%d)", F); return PANGU_ERR_BAD_ARG; }
  if (M == 0) return PANGU_OK;
  const int rpb = 192 / (F / 8);
  long long want = (M + rpb - 1) / rpb;
  const unsigned grid = (unsigned)(want < 148LL * 16 ? want : 148LL * 16);
  gelu_bwd_bf16_kernel<<<grid, 192, 0, as_stream(stream)>>>((const uint4*)dh, (const uint4*)h_pre, (uint4*)dh_pre, dcolsum, M, F);
  return check_launch("gelu_backward_bf16");
}

extern "C" int pangu_colsum(const void* x, int dtype, int64_t ld, int64_t M, int32_t C, float* out, void* stream) {
  if (!x || !out || M < 0 || C <= 0 || (C & 1) || (ld & 1)) { set_error("colsum: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (M == 0) return PANGU_OK;
  dim3 grid(row_grid(M, 8 * 16), (unsigned)((C + 63) / 64));
  if (dtype == PANGU_BF16) colsum_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, ld, M, C, out);
  else if (dtype == PANGU_F32) colsum_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)x, ld, M, C, out);
  else { set_error("colsum: unknown dtype"); return PANGU_ERR_BAD_ARG; }
  return check_launch("colsum");
}

extern "C" int pangu_patch_recover_gather_backward(const float* d_output, const float* d_output_surface, void* dy_upper,
                                                   void* dy_surface, int32_t lat_rows, int32_t tok_rows, void* stream) {
  if (!d_output || !d_output_surface || !dy_upper || !dy_surface) { set_error("patch_recover_gather_backward: null pointer"); return PANGU_ERR_BAD_ARG; }
  if (lat_rows <= 0 || tok_rows <= 0 || 4 * tok_rows < lat_rows) { set_error("patch_recover_gather_backward: inconsistent rows"); return PANGU_ERR_BAD_ARG; }
  cudaStream_t st = as_stream(stream);
  dim3 gu((kTokW + kTileTok - 1) / kTileTok, tok_rows, 7), gs((kTokW + kTileTok - 1) / kTileTok, tok_rows, 1);
  patch_recover_bwd_kernel<160, 40, true><<<gu, 256, 0, st>>>(d_output, (__nv_bfloat16*)dy_upper, lat_rows, tok_rows);
  patch_recover_bwd_kernel<64, 16, false><<<gs, 256, 0, st>>>(d_output_surface, (__nv_bfloat16*)dy_surface, lat_rows, tok_rows);
  return check_launch("patch_recover_gather_backward");
}

static int launch_weighted_l1(const float* out, const float* target, const float* mean, const float* stdv, const float* weight,
                              const float* mask, int32_t planes, int32_t planes_per_var, int64_t plane_elems, float scale,
                              float* loss_sum, float* d_out, void* stream) {
  if (!out || !target || !weight || !loss_sum || planes <= 0 || planes_per_var <= 0 || plane_elems <= 0 || (plane_elems & 3) ||
      ((mean == nullptr) != (stdv == nullptr)) ||
      ((((uintptr_t)out | (uintptr_t)target | (uintptr_t)mask | (uintptr_t)d_out) & 15) != 0)) {
    set_error("weighted_l1_loss: bad argument (pointers must be 16-byte aligned, plane_elems a multiple of 4)");
    return PANGU_ERR_BAD_ARG;
  }
  const long long total4 = (long long)planes * (plane_elems / 4);
  long long want = (total4 + 255) / 256;
  const unsigned grid = (unsigned)(want < 148LL * 16 ? want : 148LL * 16);
  weighted_l1_loss_kernel<<<grid, 256, 0, as_stream(stream)>>>((const float4*)out, (const float4*)target, mean, stdv, weight,
                                                              (const float4*)mask, planes_per_var, plane_elems / 4, total4, scale,
                                                              loss_sum, (float4*)d_out);
  return check_launch("weighted_l1_loss");
}

extern "C" int pangu_weighted_l1_loss(const float* out, const float* target, const float* mean, const float* stdv,
                                      const float* weight, int32_t planes, int32_t planes_per_var, int64_t plane_elems,
                                      float scale, float* loss_sum, float* d_out, void* stream) {
  return launch_weighted_l1(out, target, mean, stdv, weight, nullptr, planes, planes_per_var, plane_elems, scale, loss_sum, d_out, stream);
}

extern "C" int pangu_weighted_l1_loss_masked(const float* out, const float* target, const float* mean, const float* stdv,
                                             const float* weight, const float* mask, int32_t planes, int32_t planes_per_var,
                                             int64_t plane_elems, float scale, float* loss_sum, float* d_out, void* stream) {
  return launch_weighted_l1(out, target, mean, stdv, weight, mask, planes, planes_per_var, plane_elems, scale, loss_sum, d_out, stream);
}

extern "C" int pangu_wind_speed_l1_loss(const float* out_u, const float* out_v, const float* tgt_u, const float* tgt_v,
                                        const float* mean_u, const float* std_u, const float* mean_v, const float* std_v,
                                        const float* mask, int32_t planes, int64_t plane_elems, float scale, float* loss_sum,
                                        float* d_u, float* d_v, void* stream) {
  const bool stats = mean_u != nullptr;
  if (!out_u || !out_v || !tgt_u || !tgt_v || !loss_sum || planes <= 0 || plane_elems <= 0 || (plane_elems & 3) ||
      (stats != (std_u != nullptr)) || (stats != (mean_v != nullptr)) || (stats != (std_v != nullptr)) ||
      ((d_u == nullptr) != (d_v == nullptr)) ||
      ((((uintptr_t)out_u | (uintptr_t)out_v | (uintptr_t)tgt_u | (uintptr_t)tgt_v | (uintptr_t)mask | (uintptr_t)d_u |
         (uintptr_t)d_v) & 15) != 0)) {
    set_error("wind_speed_l1_loss: bad argument (pointers must be 16-byte aligned, plane_elems a multiple of 4)");
    return PANGU_ERR_BAD_ARG;
  }
  const long long total4 = (long long)planes * (plane_elems / 4);
  long long want = (total4 + 255) / 256;
  const unsigned grid = (unsigned)(want < 148LL * 16 ? want : 148LL * 16);
  wind_speed_l1_loss_kernel<<<grid, 256, 0, as_stream(stream)>>>((const float4*)out_u, (const float4*)out_v, (const float4*)tgt_u,
                                                                (const float4*)tgt_v, mean_u, std_u, mean_v, std_v,
                                                                (const float4*)mask, plane_elems / 4, total4, scale, loss_sum,
                                                                (float4*)d_u, (float4*)d_v);
  return check_launch("wind_speed_l1_loss");
}

extern "C" int pangu_lat_weighted_score_sums(const float* pred, const float* target, const float* mask, const float* clim,
                                             const float* lat_weight, int32_t planes, int32_t H, int32_t W, double* sums,
                                             void* stream) {
  if (!pred || !target || !lat_weight || !sums || planes <= 0 || H <= 0 || W <= 0) { set_error("lat_weighted_score_sums: bad argument"); return PANGU_ERR_BAD_ARG; }
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 5 * planes, as_stream(stream));
  if (e != cudaSuccess) { set_error("lat_weighted_score_sums: cudaMemsetAsync: %s", cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
  int chunks = (148 * 8 + planes - 1) / planes;                 // ~8 CTAs per SM in total
  if (chunks > H) chunks = H;
  const dim3 grid((unsigned)planes, (unsigned)chunks);
  const bool vec = W % 4 == 0 && (((uintptr_t)pred | (uintptr_t)target | (uintptr_t)mask) & 15) == 0;
  if (vec) lat_weighted_score_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(pred, target, mask, clim, lat_weight, H, W, sums);
  else     lat_weighted_score_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(pred, target, mask, clim, lat_weight, H, W, sums);
  return check_launch("lat_weighted_score_sums");
}
