// Shared helpers for libpangu_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pangu_b200.h"

namespace pangu {

constexpr int kWinZ = 2, kWinH = 6, kWinW = 12;
constexpr int kWinTokens = 144;
constexpr int kPadH = 5;
constexpr int kHeadDim = 32;
constexpr float kMaskValue = -100.0f;

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// Window-grid geometry of one stage (models/layers.py:228,253-262).
struct WinGeom {
  int Z, H, W, C, heads;
  int Hp, nZ, nH, nLon, T;
};

// Latitude band of the window grid handled by one launch (see tc_attention.cu).
struct BandGeom {
  int h0, hrows;     // qkv/out hold global rows [h0, h0 + hrows)
  int hw0, nhw;      // h-windows [hw0, hw0 + nhw) ; with wrap the last one is the global window nH-1
  int wrap;
  int halo;          // rows available in the halo buffers right after the own rows
  int halo_lo;       // rows available in the northern halo buffer right before the own rows
  int halo_kv = 0;   // 1: the halo buffers hold only the K and V columns, [rows, 2C] (PANGU_ATTN_HALO_KV; tcgen05 kernel only)
};

inline bool make_geom(const pangu_geom* g, WinGeom& o) {
  if (!g) return false;
  o.Z = g->Z; o.H = g->H; o.W = g->W; o.C = g->C; o.heads = g->heads;
  o.Hp = g->H + kPadH;
  if (o.Z <= 0 || o.H <= 0 || o.W <= 0 || o.C <= 0) return false;
  if (o.Z % kWinZ || o.Hp % kWinH || o.W % kWinW) return false;
  o.nZ = o.Z / kWinZ; o.nH = o.Hp / kWinH; o.nLon = o.W / kWinW; o.T = o.nZ * o.nH;
  return true;
}

// Source token of window element (l, t, k), or -1 for a zero pad row.
// pad (layers.py:228) + roll by (-1,-3,-6) on the padded grid (:237) + partition (:253-262).
__host__ __device__ __forceinline__ long long window_source(const WinGeom& g, int l, int t, int k,
                                                            int roll) {
  if (roll == 2) return ((long long)l * g.T + t) * kWinTokens + k;   // pre-partitioned windows: identity
  const int zw = t / g.nH, hw = t - zw * g.nH;
  const int dz = k / 72, r = k - dz * 72;
  const int dh = r / 12, dw = r - dh * 12;
  int z = 2 * zw + dz, h = 6 * hw + dh, w = 12 * l + dw;
  if (roll == 1) {
    z += 1; if (z >= g.Z) z -= g.Z;
    h += 3; if (h >= g.Hp) h -= g.Hp;
    w += 6; if (w >= g.W) w -= g.W;
  }
  return h < g.H ? ((long long)z * g.H + h) * g.W + w : -1LL;
}

// Compact shift-mask group id of element k in window type t (attention-equivalent to gen_mask,
// layers.py:187-216): differs <=> masked.
__host__ __device__ __forceinline__ int shift_group(const WinGeom& g, int t, int k) {
  const int zw = t / g.nH, hw = t - zw * g.nH;
  const int dz = k / 72, dh = (k - dz * 72) / 12;
  return (zw == g.nZ - 1 ? 2 * dz : 0) + ((hw == g.nH - 1 && dh >= 3) ? 1 : 0);
}

// The reference's own region id (slice painting order of gen_mask) -- used by pangu_shift_mask so
// that the exported mask is derived the same way the reference derives it.
__host__ __device__ __forceinline__ int shift_region_reference(const WinGeom& g, int t, int k) {
  const int zw = t / g.nH, hw = t - zw * g.nH;
  const int dz = k / 72, dh = (k - dz * 72) / 12;
  const int z = 2 * zw + dz, h = 6 * hw + dh;
  // z slices: [0, Z-2) -> 0, [Z-2, Z-1) -> 1, [Z-1, Z) -> 2
  const int zi = z < g.Z - 2 ? 0 : (z < g.Z - 1 ? 1 : 2);
  // h slices painted in order [0,Hp-6) -> 0, [6,Hp-3) -> 1, [Hp-3,Hp) -> 2; later wins.
  int hi = 0;
  if (h >= 6 && h < g.Hp - 3) hi = 1;
  if (h >= g.Hp - 3) hi = 2;
  return zi * 3 + hi;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 2- and 4-element vector accesses of the row kernels (layout_kernels.cu, bwd_kernels.cu) (8 / 16 bytes per lane instead of 2 / 4)
__device__ __forceinline__ float2 ld2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ld2(const __nv_bfloat16* p) {
  const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void st2(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}


inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// cudaFuncSetAttribute is per device: `mask` is a function-local static bitset over device ordinals (one process may
// drive several GPUs: the reference builds torch.device('cuda:' + str(local_rank)) without cudaSetDevice).
template <typename K>
inline cudaError_t set_max_smem_once(unsigned long long& mask, K kernel, int bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) mask |= bit;
  return e;
}

}  // namespace pangu
