// Bandwidth kernels: patch-embed gather, patch-recover scatter, down/up-sample reshuffles fused with
// their LayerNorm, casts.  All are HBM-bound: coalesced 128-bit global accesses, smem transposes
// where the patch layout requires one, warp-shuffle LayerNorm statistics.
#include "common.cuh"

namespace pangu {

constexpr int kLat = 721, kLon = 1440, kLev = 13;
constexpr int kLatPad = 724;                     // models/layers.py:37,49 (+3 rows)
constexpr int kTokH = 181, kTokW = 360;          // 724/4, 1440/4
// The kernels take the latitude extent at run time (lat = valid pixel rows of the arrays they are given,
// tokH = token rows, mapRows = rows of the constant-map array) so that a latitude band of the grid can be
// embedded / recovered on its own; the full grid is lat = 721, tokH = 181, mapRows = 724.
constexpr int kTileTok = 32;                     // tokens (along w') per CTA

template <typename T>
__device__ __forceinline__ void store_tile_row_major(T* __restrict__ dst, const float* tile, int pitch,
                                                     int ntok, int F, int tid, int nthreads);
template <>
__device__ __forceinline__ void store_tile_row_major<float>(float* __restrict__ dst, const float* tile,
                                                            int pitch, int ntok, int F, int tid, int nthreads) {
  for (int i = tid; i < ntok * F; i += nthreads) dst[i] = tile[(i / F) * pitch + (i % F)];
}
template <>
__device__ __forceinline__ void store_tile_row_major<__nv_bfloat16>(__nv_bfloat16* __restrict__ dst, const float* tile,
                                                                    int pitch, int ntok, int F, int tid, int nthreads) {
  __nv_bfloat162* d2 = reinterpret_cast<__nv_bfloat162*>(dst);
  const int F2 = F / 2;
  for (int i = tid; i < ntok * F2; i += nthreads) {
    const int r = i / F2, c = (i - r * F2) * 2;
    d2[i] = __floats2bfloat162_rn(tile[r * pitch + c], tile[r * pitch + c + 1]);
  }
}

// Upper air: grid (12, 181, 7).  Feature f = c*32 + pz*16 + ph*4 + pw (layers.py:107-112).
template <typename T>
__global__ void __launch_bounds__(256)
patch_embed_upper_kernel(const float* __restrict__ input, const float* __restrict__ const_h,
                         const float* __restrict__ upper_mean, const float* __restrict__ upper_std,
                         T* __restrict__ patches, int lat, int tokH) {
  constexpr int F = 192, PITCH = F + 1;
  __shared__ float tile[kTileTok * PITCH];
  const int w0 = blockIdx.x * kTileTok, hp = blockIdx.y, zp = blockIdx.z;
  const int ntok = min(kTileTok, kTokW - w0);
  for (int i = threadIdx.x; i < 48 * kTileTok; i += blockDim.x) {
    const int tok = i & (kTileTok - 1);
    const int r = i / kTileTok;                     // r = c*8 + pz*4 + ph
    const int c = r >> 3, pz = (r >> 2) & 1, ph = r & 3;
    const int lev = 2 * zp + pz, y = 4 * hp + ph;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tok < ntok && lev < kLev && y < lat) {
      const long long off = ((long long)lev * lat + y) * kLon + 4 * (w0 + tok);
      if (c < 5) {
        v = __ldg(reinterpret_cast<const float4*>(input + (long long)c * kLev * lat * kLon + off));
        // statistics are stored in flipped level order (layers.py:95-99): index 12 - lev
        const float m = __ldg(upper_mean + (kLev - 1 - lev) * 5 + c);
        const float s = __ldg(upper_std + (kLev - 1 - lev) * 5 + c);
        v.x = (v.x - m) / s; v.y = (v.y - m) / s; v.z = (v.z - m) / s; v.w = (v.w - m) / s;
      } else {
        v = __ldg(reinterpret_cast<const float4*>(const_h + off));
      }
    }
    float* d = tile + tok * PITCH + c * 32 + pz * 16 + ph * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const long long tok0 = ((long long)zp * tokH + hp) * kTokW + w0;
  store_tile_row_major<T>(patches + tok0 * F, tile, PITCH, ntok, F, threadIdx.x, blockDim.x);
}

// Surface: grid (12, 181).  Feature f = c*16 + ph*4 + pw, c = 4 variables then 3 constant maps
// (layers.py:75-87).  maps already has 724 rows; the variables are zero-padded AFTER normalisation.
template <typename T>
__global__ void __launch_bounds__(256)
patch_embed_surface_kernel(const float* __restrict__ input_surface, const float* __restrict__ maps,
                           const float* __restrict__ surface_mean, const float* __restrict__ surface_std,
                           T* __restrict__ patches, int lat, int mapRows) {
  constexpr int F = 112, PITCH = F + 1;
  __shared__ float tile[kTileTok * PITCH];
  const int w0 = blockIdx.x * kTileTok, hp = blockIdx.y;
  const int ntok = min(kTileTok, kTokW - w0);
  for (int i = threadIdx.x; i < 28 * kTileTok; i += blockDim.x) {
    const int tok = i & (kTileTok - 1);
    const int r = i / kTileTok;                     // r = c*4 + ph
    const int c = r >> 2, ph = r & 3;
    const int y = 4 * hp + ph;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tok < ntok) {
      if (c < 4) {
        if (y < lat) {
          v = __ldg(reinterpret_cast<const float4*>(input_surface + ((long long)c * lat + y) * kLon + 4 * (w0 + tok)));
          const float m = __ldg(surface_mean + c), s = __ldg(surface_std + c);
          v.x = (v.x - m) / s; v.y = (v.y - m) / s; v.z = (v.z - m) / s; v.w = (v.w - m) / s;
        }
      } else {
        if (y < mapRows)
          v = __ldg(reinterpret_cast<const float4*>(maps + ((long long)(c - 4) * mapRows + y) * kLon + 4 * (w0 + tok)));
      }
    }
    float* d = tile + tok * PITCH + c * 16 + ph * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const long long tok0 = (long long)hp * kTokW + w0;
  store_tile_row_major<T>(patches + tok0 * F, tile, PITCH, ntok, F, threadIdx.x, blockDim.x);
}

// Un-patchify (layers.py:593-603 upper, :609-619 surface).  One CTA = 32 tokens along w'.
template <int kF, int kRows, bool kUpper>
__global__ void __launch_bounds__(256)
patch_recover_kernel(const float* __restrict__ y, float* __restrict__ out, int lat, int tokH,
                     const float* __restrict__ scale, const float* __restrict__ shift) {
  constexpr int PITCH = kF + 1;
  __shared__ float tile[kTileTok * PITCH];
  const int w0 = blockIdx.x * kTileTok, hp = blockIdx.y, zp = blockIdx.z;
  const int ntok = min(kTileTok, kTokW - w0);
  const long long tok0 = ((long long)zp * tokH + hp) * kTokW + w0;
  const float* src = y + tok0 * kF;
  for (int i = threadIdx.x; i < ntok * kF; i += blockDim.x) tile[(i / kF) * PITCH + (i % kF)] = __ldg(src + i);
  __syncthreads();
  for (int i = threadIdx.x; i < kRows * kTileTok; i += blockDim.x) {
    const int tok = i & (kTileTok - 1);
    const int r = i / kTileTok;
    if (tok >= ntok) continue;
    int v, lev, yy;
    if (kUpper) {                                   // r = v*8 + pz*4 + ph ; channel = v*32+pz*16+ph*4+pw
      v = r >> 3; lev = 2 * zp + ((r >> 2) & 1); yy = 4 * hp + (r & 3);
      if (lev >= kLev || yy >= lat) continue;
    } else {                                        // r = v*4 + ph ; channel = v*16+ph*4+pw
      v = r >> 2; lev = 0; yy = 4 * hp + (r & 3);
      if (yy >= lat) continue;
    }
    const float* s = tile + tok * PITCH + r * 4;
    const long long plane = kUpper ? ((long long)v * kLev + lev) : (long long)v;
    float4 o = make_float4(s[0], s[1], s[2], s[3]);
    if (scale != nullptr) {                           // physical units: x * std + mean (era5_data/utils_data.py:540-546)
      const float a = __ldg(scale + plane), b = __ldg(shift + plane);
      o.x = fmaf(o.x, a, b); o.y = fmaf(o.y, a, b); o.z = fmaf(o.z, a, b); o.w = fmaf(o.w, a, b);
    }
    *reinterpret_cast<float4*>(out + (plane * lat + yy) * kLon + 4 * (w0 + tok)) = o;
  }
}

// DownSample: 2x2 merge + LN(4C) (layers.py:501-519); one warp per output row, a lane owns groups of 4 consecutive
// features (C % 4 == 0, so a group never straddles two source tokens): 16-byte loads, 8- / 16-byte stores.
template <typename TO, int kPerLane>
__global__ void __launch_bounds__(256)
downsample_merge_ln_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                           const float* __restrict__ beta, TO* __restrict__ out, int Z, int H, int W,
                           int C, float eps) {
  constexpr int F = kPerLane * 32;                 // 4C
  constexpr int G = kPerLane / 4;                  // groups of 4 per lane
  static_assert(kPerLane % 4 == 0, "downsample_merge_ln: 4-element groups");
  const int H2 = (H + 1) / 2, W2 = W / 2;
  const long long rows = (long long)Z * H2 * W2;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int w2 = (int)(row % W2);
  const long long zh = row / W2;
  const int h2 = (int)(zh % H2), z = (int)(zh / H2);
  float4 v[G];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const int f = (i * 32 + lane) * 4;
    const int q = f / C, c = f - q * C;            // f = dh*2C + dw*C + c
    const int h = 2 * h2 + (q >> 1), w = 2 * w2 + (q & 1);
    v[i] = (h < H) ? __ldg(reinterpret_cast<const float4*>(x + (((long long)z * H + h) * W + w) * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / F);
  float qv = 0.f;
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
    qv = fmaf(d0, d0, qv); qv = fmaf(d1, d1, qv); qv = fmaf(d2, d2, qv); qv = fmaf(d3, d3, qv);
  }
  const float rstd = rsqrtf(warp_sum(qv) * (1.0f / F) + eps);
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const int f = (i * 32 + lane) * 4;
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + f)), be = __ldg(reinterpret_cast<const float4*>(beta + f));
    st4(out + row * F + f, (v[i].x - mean) * rstd * ga.x + be.x, (v[i].y - mean) * rstd * ga.y + be.y,
        (v[i].z - mean) * rstd * ga.z + be.z, (v[i].w - mean) * rstd * ga.w + be.w);
  }
}

// UpSample: pixel shuffle + crop + LN(C') (layers.py:546-563); one warp per output row, a lane owns pairs of features.
template <typename TI, typename TO, int kPerLane>
__global__ void __launch_bounds__(256)
upsample_shuffle_ln_kernel(const TI* __restrict__ y, const float* __restrict__ gamma,
                           const float* __restrict__ beta, TO* __restrict__ out, int Z, int H2, int W2,
                           int H, float eps) {
  constexpr int Co = kPerLane * 32;
  constexpr int G = kPerLane / 2;                  // pairs per lane
  static_assert(kPerLane % 2 == 0, "upsample_shuffle_ln: 2-element groups");
  const int W = 2 * W2;
  const long long rows = (long long)Z * H * W;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int w = (int)(row % W);
  const long long zh = row / W;
  const int h = (int)(zh % H), z = (int)(zh / H);
  const TI* src = y + (((long long)z * H2 + (h >> 1)) * W2 + (w >> 1)) * (4 * Co) + ((h & 1) * 2 + (w & 1)) * Co;
  float2 v[G];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < G; ++i) { v[i] = ld2(src + (i * 32 + lane) * 2); s += v[i].x + v[i].y; }
  const float mean = warp_sum(s) * (1.0f / Co);
  float qv = 0.f;
#pragma unroll
  for (int i = 0; i < G; ++i) { const float d0 = v[i].x - mean, d1 = v[i].y - mean; qv = fmaf(d0, d0, qv); qv = fmaf(d1, d1, qv); }
  const float rstd = rsqrtf(warp_sum(qv) * (1.0f / Co) + eps);
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const int c = (i * 32 + lane) * 2;
    const float2 ga = __ldg(reinterpret_cast<const float2*>(gamma + c)), be = __ldg(reinterpret_cast<const float2*>(beta + c));
    st2(out + row * Co + c, (v[i].x - mean) * rstd * ga.x + be.x, (v[i].y - mean) * rstd * ga.y + be.y);
  }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(in + i);
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  out[i] = o;
}

__global__ void __launch_bounds__(256)
concat_cast_bf16_kernel(const float* __restrict__ a, const float* __restrict__ b, uint2* __restrict__ out,
                        long long n, int C1, int C2) {
  const int q = (C1 + C2) / 4;                     // float4 groups per output row
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * q) return;
  const long long row = i / q;
  const int c = (int)(i - row * q) * 4;
  const float4 v = c < C1 ? __ldg(reinterpret_cast<const float4*>(a + row * C1 + c))
                          : __ldg(reinterpret_cast<const float4*>(b + row * C2 + (c - C1)));
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&lo);
  o.y = *reinterpret_cast<uint32_t*>(&hi);
  out[i] = o;
}

}  // namespace pangu

using namespace pangu;

extern "C" int pangu_patch_embed_gather_rows(const float* input, const float* input_surface,
                                             const float* surface_mean, const float* surface_std,
                                             const float* upper_mean, const float* upper_std,
                                             const float* maps, const float* const_h, void* patches_surface,
                                             void* patches_upper, int out_dtype, int32_t lat_rows,
                                             int32_t tok_rows, int32_t map_rows, void* stream) {
  if (!input || !input_surface || !surface_mean || !surface_std || !upper_mean || !upper_std || !maps ||
      !const_h || !patches_surface || !patches_upper) { set_error("patch_embed_gather: null pointer"); return PANGU_ERR_BAD_ARG; }
  if (lat_rows <= 0 || tok_rows <= 0 || 4 * tok_rows < lat_rows || 4 * (tok_rows - 1) >= lat_rows || map_rows < lat_rows) {
    set_error("patch_embed_gather: inconsistent rows (lat %d, tokens %d, maps %d)", lat_rows, tok_rows, map_rows);
    return PANGU_ERR_BAD_ARG;
  }
  cudaStream_t st = as_stream(stream);
  dim3 gu((kTokW + kTileTok - 1) / kTileTok, tok_rows, 7), gs((kTokW + kTileTok - 1) / kTileTok, tok_rows);
  if (out_dtype == PANGU_F32) {
    patch_embed_upper_kernel<float><<<gu, 256, 0, st>>>(input, const_h, upper_mean, upper_std, (float*)patches_upper, lat_rows, tok_rows);
    patch_embed_surface_kernel<float><<<gs, 256, 0, st>>>(input_surface, maps, surface_mean, surface_std, (float*)patches_surface, lat_rows, map_rows);
  } else {
    patch_embed_upper_kernel<__nv_bfloat16><<<gu, 256, 0, st>>>(input, const_h, upper_mean, upper_std, (__nv_bfloat16*)patches_upper, lat_rows, tok_rows);
    patch_embed_surface_kernel<__nv_bfloat16><<<gs, 256, 0, st>>>(input_surface, maps, surface_mean, surface_std, (__nv_bfloat16*)patches_surface, lat_rows, map_rows);
  }
  return check_launch("patch_embed_gather");
}

extern "C" int pangu_patch_embed_gather(const float* input, const float* input_surface,
                                        const float* surface_mean, const float* surface_std,
                                        const float* upper_mean, const float* upper_std,
                                        const float* maps, const float* const_h, void* patches_surface,
                                        void* patches_upper, int out_dtype, void* stream) {
  return pangu_patch_embed_gather_rows(input, input_surface, surface_mean, surface_std, upper_mean, upper_std, maps,
                                       const_h, patches_surface, patches_upper, out_dtype, kLat, kTokH, kLatPad, stream);
}

extern "C" int pangu_patch_recover_scatter_denorm(const float* y_upper, const float* y_surface, float* output,
                                                  float* output_surface, int32_t lat_rows, int32_t tok_rows,
                                                  const float* upper_std, const float* upper_mean,
                                                  const float* surface_std, const float* surface_mean,
                                                  void* stream) {
  if (!y_upper || !y_surface || !output || !output_surface) { set_error("patch_recover_scatter: null pointer"); return PANGU_ERR_BAD_ARG; }
  if (lat_rows <= 0 || tok_rows <= 0 || 4 * tok_rows < lat_rows) { set_error("patch_recover_scatter: inconsistent rows"); return PANGU_ERR_BAD_ARG; }
  cudaStream_t st = as_stream(stream);
  dim3 gu((kTokW + kTileTok - 1) / kTileTok, tok_rows, 7), gs((kTokW + kTileTok - 1) / kTileTok, tok_rows, 1);
  if ((upper_std == nullptr) != (upper_mean == nullptr) || (surface_std == nullptr) != (surface_mean == nullptr) ||
      (upper_std == nullptr) != (surface_std == nullptr)) { set_error("patch_recover_scatter: give all four statistics or none"); return PANGU_ERR_BAD_ARG; }
  patch_recover_kernel<160, 40, true><<<gu, 256, 0, st>>>(y_upper, output, lat_rows, tok_rows, upper_std, upper_mean);
  patch_recover_kernel<64, 16, false><<<gs, 256, 0, st>>>(y_surface, output_surface, lat_rows, tok_rows, surface_std, surface_mean);
  return check_launch("patch_recover_scatter");
}

extern "C" int pangu_patch_recover_scatter_rows(const float* y_upper, const float* y_surface, float* output,
                                                float* output_surface, int32_t lat_rows, int32_t tok_rows,
                                                void* stream) {
  return pangu_patch_recover_scatter_denorm(y_upper, y_surface, output, output_surface, lat_rows, tok_rows, nullptr,
                                            nullptr, nullptr, nullptr, stream);
}

extern "C" int pangu_patch_recover_scatter(const float* y_upper, const float* y_surface, float* output,
                                           float* output_surface, void* stream) {
  return pangu_patch_recover_scatter_rows(y_upper, y_surface, output, output_surface, kLat, kTokH, stream);
}

extern "C" int pangu_downsample_merge_ln(const float* x, const float* gamma, const float* beta, void* out,
                                         int out_dtype, int32_t Z, int32_t H, int32_t W, int32_t C,
                                         float eps, void* stream) {
  if (!x || !gamma || !beta || !out || Z <= 0 || H <= 0 || W <= 0 || (W & 1)) { set_error("downsample_merge_ln: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (C != 192) { set_error("downsample_merge_ln: C=%d unsupported (192)", C); return PANGU_ERR_UNSUPPORTED; }
  const long long rows = (long long)Z * ((H + 1) / 2) * (W / 2);
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  cudaStream_t st = as_stream(stream);
  if (out_dtype == PANGU_F32)
    downsample_merge_ln_kernel<float, 24><<<blocks, 256, 0, st>>>(x, gamma, beta, (float*)out, Z, H, W, C, eps);
  else
    downsample_merge_ln_kernel<__nv_bfloat16, 24><<<blocks, 256, 0, st>>>(x, gamma, beta, (__nv_bfloat16*)out, Z, H, W, C, eps);
  return check_launch("downsample_merge_ln");
}

extern "C" int pangu_upsample_shuffle_ln(const void* y, int y_dtype, const float* gamma, const float* beta,
                                         void* out, int out_dtype, int32_t Z, int32_t H2, int32_t W2,
                                         int32_t H, int32_t Cout, float eps, void* stream) {
  if (!y || !gamma || !beta || !out || Z <= 0 || H2 <= 0 || W2 <= 0 || H > 2 * H2) { set_error("upsample_shuffle_ln: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (Cout != 192) { set_error("upsample_shuffle_ln: Cout=%d unsupported (192)", Cout); return PANGU_ERR_UNSUPPORTED; }
  const long long rows = (long long)Z * H * 2 * W2;
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  cudaStream_t st = as_stream(stream);
  if (y_dtype == PANGU_F32 && out_dtype == PANGU_F32)
    upsample_shuffle_ln_kernel<float, float, 6><<<blocks, 256, 0, st>>>((const float*)y, gamma, beta, (float*)out, Z, H2, W2, H, eps);
  else if (y_dtype == PANGU_BF16 && out_dtype == PANGU_BF16)
    upsample_shuffle_ln_kernel<__nv_bfloat16, __nv_bfloat16, 6><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)y, gamma, beta, (__nv_bfloat16*)out, Z, H2, W2, H, eps);
  else if (y_dtype == PANGU_F32 && out_dtype == PANGU_BF16)
    upsample_shuffle_ln_kernel<float, __nv_bfloat16, 6><<<blocks, 256, 0, st>>>((const float*)y, gamma, beta, (__nv_bfloat16*)out, Z, H2, W2, H, eps);
  else { set_error("upsample_shuffle_ln: unsupported dtype combination"); return PANGU_ERR_UNSUPPORTED; }
  return check_launch("upsample_shuffle_ln");
}

extern "C" int pangu_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream) {
  if (!in || !out || n < 0 || (n & 3)) { set_error("cast_f32_bf16: n must be a multiple of 4"); return PANGU_ERR_BAD_ARG; }
  if (n == 0) return PANGU_OK;
  const long long n4 = n / 4;
  cast_f32_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, as_stream(stream)>>>((const float4*)in, (uint2*)out, n4);
  return check_launch("cast_f32_bf16");
}

extern "C" int pangu_concat_cast_bf16(const float* a, const float* b, void* out, int64_t n, int32_t C1,
                                      int32_t C2, void* stream) {
  if (!a || !b || !out || n < 0 || (C1 & 3) || (C2 & 3)) { set_error("concat_cast_bf16: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (n == 0) return PANGU_OK;
  const long long total = n * ((C1 + C2) / 4);
  concat_cast_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(a, b, (uint2*)out, n, C1, C2);
  return check_launch("concat_cast_bf16");
}
