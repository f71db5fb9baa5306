// Row-tile epilogue shared by the tcgen05 kernels:  x_out = residual + LayerNorm_C(Y + bias) * gamma + beta
// for a [128 x C] fp32 accumulator tile in TMEM (models/layers.py:296-297), executed by 8 epilogue warps
// (two per TMEM lane quarter; warp (q, hf) handles the 16-column units hf, hf+2, ... of rows [32q, 32q+32)).
//
//   prefetch()       issue the row-contiguous loads of the first three residual units (can run before the
//                    accumulator is complete);
//   stats(y)         pass 1: row sums of x and x^2 from TMEM, combined with the sibling warp through smem
//                    (named barrier 1 over the 256 epilogue threads) -> mean, rstd;
//   unit<S>(y, i)    pass 2 for unit i (S = i % 3, the register slot its residual was fetched into): stage the
//                    residual in a swizzled per-warp smem tile, add the normalised accumulator in the
//                    thread-per-row (TMEM) view, write fp32 + bf16 row-contiguously, prefetch unit i + 3.
//
// Splitting pass 2 into units lets a kernel interleave the HBM-heavy part with other work (tc_mlp.cu runs one
// unit after each hidden chunk of the NEXT row tile when the accumulator is double-buffered).
#pragma once
#include "tc_common.cuh"

namespace pangu {
namespace tc {

__device__ __forceinline__ int ln_stg16(int r, int cc) { return r * 64 + ((cc ^ ((r >> 1) & 3)) << 4); }

// UW = unit width in columns (16: 64-byte row segments, 2 KiB staging per warp; 32: 128-byte segments, 4 KiB --
// fewer instructions per byte), D = prefetch depth in units (register slots).
// ASYNC: the residual tiles are prefetched by ONE TMA load per warp and unit straight into (hardware-swizzled)
// staging tiles -- NBUF = D + 1 tiles and mbarriers per warp, D units in flight -- no registers held across
// other work (the interleaved GELU chunks of tc_mlp.cu) and no per-lane address arithmetic.
// TMASTORE: results leave through TMA bulk stores issued by one lane from the (hardware-swizzle-compatible) staging
// tiles -- fp32 tile in place, bf16 tile in `stg_b` -- instead of 2 x NV st.global per lane; the affine parameters
// (bias, gamma, beta) are then read from a shared-memory copy `sparams` [3][C] (broadcast LDS) rather than __ldg.
// NB16 = bf16 staging tiles per warp: unit idx only waits for the bulk stores of unit idx - NB16 to have read their
// source (cp.async.bulk.wait_group.read NB16-1), so with NB16 = 2 the store of one unit drains while the next unit is
// computed (r2: the single-tile version serialised every unit behind its predecessor's store, ~2.5 k clocks per unit).
// The fp32 tile of unit idx + D is the one unit idx + D - NBUF used, whose store must be complete: NBUF >= D + NB16.
// NPART = warps per TMEM lane quarter (2: eight epilogue warps; 4: sixteen): warp (q, hf) takes the units hf, hf + NPART, ...
template <int C, int UW = 16, int D = 3, bool ASYNC = false, bool TMASTORE = false, int NB16 = 1, int NBUF_ = 0, int NPART = 2>
struct LnTileEpilogue {
  const CUtensorMap* tm_out = nullptr;               // fp32 [M, C], box UW x 32, swizzle UW*4 bytes
  const CUtensorMap* tm_xb = nullptr;                // bf16 [M, C], box UW x 32, swizzle UW*2 bytes (or null)
  uint8_t* stg_b = nullptr;                          // per warp: NB16 tiles of 32 rows x UW bf16
  const float* sparams = nullptr;                    // smem [3][C]: bias, gamma, beta
  const CUtensorMap* tm_res = nullptr;               // ASYNC: fp32 residual [M, C], same box / swizzle as tm_out
  uint64_t* ld_bar = nullptr;                        // ASYNC: this warp's NBUF mbarriers (count 1) for the tile loads
  uint32_t ld_phase = 0;                             // bit b = phase of ld_bar[b]
  __device__ __forceinline__ static int stg_b_off(int r, int cc) {   // 16-byte chunk cc of bf16 row r
    return UW == 32 ? r * 64 + ((cc ^ ((r >> 1) & 3)) << 4) : r * 32 + ((cc ^ ((r >> 2) & 1)) << 4);
  }
  static constexpr int UNIT_BYTES = UW * 128;        // one staging tile: 32 rows x UW fp32
  static constexpr int NBUF = ASYNC ? (NBUF_ > 0 ? NBUF_ : D + NB16) : 1;   // fp32 staging tiles per warp (ASYNC: D landing + NB16 draining)
  static constexpr int STGB_TILE = UW * 64;          // one bf16 staging tile: 32 rows x UW bf16
  static_assert(!ASYNC || NBUF >= D + NB16, "a refilled fp32 tile must belong to a unit whose store has been waited for");
  static_assert(ASYNC || NB16 == 1, "register-prefetch mode stages through one tile");
  static constexpr int NU = C / (NPART * UW);        // units per warp
  static_assert(C % (NPART * UW) == 0 && (C / 32) % NPART == 0, "columns must split evenly over the warps of a lane quarter");
  static constexpr int NV = UW / 4;                  // float4 per lane per unit
  static constexpr int RPI = 128 / UW;               // rows covered by one warp-wide 16-byte access (8 or 4)
  static_assert(UW == 16 || UW == 32, "unit width");
  static_assert(ASYNC || NU % D == 0, "register prefetch: units are processed in groups of the prefetch depth");
  __device__ __forceinline__ static int stg_off(int r, int cc) {
    return UW == 16 ? r * 64 + ((cc ^ ((r >> 1) & 3)) << 4) : r * 128 + ((cc ^ (r & 7)) << 4);
  }
  const float* bias;
  const float* gamma;
  const float* beta;
  const float* residual;
  float* x_out;
  __nv_bfloat16* xb;
  long long M, m_base;                               // m_base: global row of this warp's first row
  long long m_res = -1;                              // row of the residual tensor map to read instead of m_base (>= 0), e.g. a scratch tile
  uint8_t* umma_x = nullptr;                         // when set: the bf16 result does not leave through tm_xb but is written into this
                                                     // [C/64][128 rows x 128 B] SW128 K-major operand tile (the x tile of a following GEMM)
  int x_row0 = 0;                                    // row of this warp's first row inside that operand tile
  float eps, mean, rstd;
  uint8_t* stg;                                      // per warp: UNIT_BYTES (x2 with ASYNC)
  float2* ln_part;                                   // [2 parity][NPART][128]
  int q, hf, lane;
  uint32_t tile_par;
  uint64_t out_hint = 0, res_hint = 0;               // L2 cache-hint policies of the fp32 tile stores / residual tile loads (0 = none)
  long long* dbg = nullptr;                          // bring-up: 6 clock64 stamps per unit when non-null
  float4 rg[ASYNC ? 1 : D][ASYNC ? 1 : NV];

  template <int S>
  __device__ __forceinline__ void load_residual(int idx) {
    const int u = hf + NPART * idx, rcc = lane & (NV - 1), rr0 = lane / NV;
    if constexpr (ASYNC) {                             // one TMA tile load per warp (rows >= M are zero-filled)
      const int b = idx % NBUF;
      if (lane == 0) {
        mbar_expect_tx(&ld_bar[b], UNIT_BYTES);
        if (res_hint) tma_load_2d_hint(stg + b * UNIT_BYTES, tm_res, &ld_bar[b], u * UW, (int)(m_res >= 0 ? m_res : m_base), res_hint);
        else tma_load_2d(stg + b * UNIT_BYTES, tm_res, &ld_bar[b], u * UW, (int)(m_res >= 0 ? m_res : m_base));
      }
      __syncwarp();
      return;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const long long m = m_base + rr0 + RPI * i;
      rg[S][i] = (residual != nullptr && m < M) ? __ldg(reinterpret_cast<const float4*>(residual + m * C + u * UW + rcc * 4))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void prefetch() {
    if constexpr (ASYNC) {
      if constexpr (TMASTORE) {                        // the previous row tile's bulk stores still read the staging tiles
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
#pragma unroll
      for (int i = 0; i < D; ++i)
        if (i < NU) load_residual<0>(i);
      return;
    }
    load_residual<0>(0);
    if constexpr (D > 1) load_residual<(D > 1 ? 1 : 0)>(1);
    if constexpr (D > 2) load_residual<(D > 2 ? 2 : 0)>(2);
  }
  // y = TMEM address of this warp's lanes, column 0 of the accumulator tile
  __device__ __forceinline__ void stats(uint32_t y) {
    uint32_t v[32];
    float s = 0.f, ss = 0.f;
#pragma unroll 1
    for (int c = hf; c < C / 32; c += NPART) {
      tmem_ld_32x32(y + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c * 32 + i));
        const float x0 = __uint_as_float(v[i]) + b.x, x1 = __uint_as_float(v[i + 1]) + b.y;
        const float x2 = __uint_as_float(v[i + 2]) + b.z, x3 = __uint_as_float(v[i + 3]) + b.w;
        s += (x0 + x1) + (x2 + x3);
        ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss); ss = fmaf(x2, x2, ss); ss = fmaf(x3, x3, ss);
      }
    }
    float2* part = ln_part + tile_par * (NPART * 128);
    part[hf * 128 + q * 32 + lane] = make_float2(s, ss);
    asm volatile("bar.sync 1, %0;" ::"n"(NPART * 128) : "memory");
    float sx = 0.f, sxx = 0.f;
#pragma unroll
    for (int k = 0; k < NPART; ++k) {
      const float2 pk = part[k * 128 + q * 32 + lane];
      sx += pk.x; sxx += pk.y;
    }
    mean = sx * (1.0f / C);
    const float var = fmaxf(sxx * (1.0f / C) - mean * mean, 0.f);
    rstd = rsqrtf(var + eps);
    tile_par ^= 1;
  }
  template <int S>
  __device__ __forceinline__ void unit(uint32_t y, int idx, bool store = true) {
    const int u = hf + NPART * idx, rcc = lane & (NV - 1), rr0 = lane / NV;
    uint8_t* stg_u = stg;
    if (dbg) dbg[0] = clock64();
    uint8_t* stg_bu = stg_b + (NB16 > 1 ? (idx % NB16) * STGB_TILE : 0);
    if constexpr (TMASTORE) {                          // the bulk stores of unit idx - NB16 must have drained their tiles
      if (lane == 0) tma_store_wait_read_n<NB16 - 1>();
      __syncwarp();
    }
    if constexpr (ASYNC) {
      // the tile that unit idx-1 consumed (and stored from) is free again: refill it with unit idx+D
      if (idx + D < NU) load_residual<0>(idx + D);     // lands in tile (idx - 1) mod NBUF
      const int b = idx % NBUF;
      stg_u = stg + b * UNIT_BYTES;
      mbar_wait(&ld_bar[b], (ld_phase >> b) & 1);
      ld_phase ^= 1u << b;
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(stg + stg_off(rr0 + RPI * i, rcc)) = rg[S][i];
    }
    __syncwarp();
    if (dbg) dbg[1] = clock64();
    if constexpr (!ASYNC) {
      if (idx + D < NU) load_residual<S>(idx + D);
    }
    if (dbg) dbg[2] = clock64();
    uint32_t w[UW];
    if constexpr (UW == 16) tmem_ld_32x16(y + u * 16, reinterpret_cast<uint32_t(&)[16]>(w));
    else tmem_ld_32x32(y + u * 32, reinterpret_cast<uint32_t(&)[32]>(w));
    tmem_ld_wait();
    if (dbg) dbg[3] = clock64();
    float4 rr[NV];                                     // this thread's row of the residual tile
#pragma unroll
    for (int cc = 0; cc < NV; ++cc) rr[cc] = *reinterpret_cast<const float4*>(stg_u + stg_off(lane, cc));
#pragma unroll
    for (int cc = 0; cc < NV; ++cc) {
      const int i = cc * 4;
      float4 b, g, be;
      if constexpr (TMASTORE) {
        b = *reinterpret_cast<const float4*>(sparams + u * UW + i);
        g = *reinterpret_cast<const float4*>(sparams + C + u * UW + i);
        be = *reinterpret_cast<const float4*>(sparams + 2 * C + u * UW + i);
      } else {
        b = __ldg(reinterpret_cast<const float4*>(bias + u * UW + i));
        g = __ldg(reinterpret_cast<const float4*>(gamma + u * UW + i));
        be = __ldg(reinterpret_cast<const float4*>(beta + u * UW + i));
      }
      rr[cc].x += fmaf((__uint_as_float(w[i]) + b.x - mean) * rstd, g.x, be.x);
      rr[cc].y += fmaf((__uint_as_float(w[i + 1]) + b.y - mean) * rstd, g.y, be.y);
      rr[cc].z += fmaf((__uint_as_float(w[i + 2]) + b.z - mean) * rstd, g.z, be.z);
      rr[cc].w += fmaf((__uint_as_float(w[i + 3]) + b.w - mean) * rstd, g.w, be.w);
    }
    if (dbg) dbg[6] = clock64();
#pragma unroll
    for (int cc = 0; cc < NV; ++cc) {
      *reinterpret_cast<float4*>(stg_u + stg_off(lane, cc)) = rr[cc];
      if constexpr (TMASTORE) {
        if (tm_xb != nullptr) {                        // 4 bf16 = 8 bytes: half of 16-byte chunk cc >> 1
          uint8_t* pb = stg_bu + stg_b_off(lane, cc >> 1) + (cc & 1) * 8;
          *reinterpret_cast<uint2*>(pb) = make_uint2(pack_bf16(rr[cc].x, rr[cc].y), pack_bf16(rr[cc].z, rr[cc].w));
        } else if (umma_x != nullptr) {                // straight into the UMMA operand tile: k-block col / 64, 16-byte chunk
          const int col = u * UW + 4 * cc, r = x_row0 + lane;           // (col % 64) / 8 XOR-swizzled by the row, 8 bytes per float4
          uint8_t* pb = umma_x + (col >> 6) * 16384 + r * 128 + ((((col & 63) >> 3) ^ (r & 7)) << 4) + ((col & 7) >> 2) * 8;
          *reinterpret_cast<uint2*>(pb) = make_uint2(pack_bf16(rr[cc].x, rr[cc].y), pack_bf16(rr[cc].z, rr[cc].w));
        }
      }
    }
    if (dbg) dbg[7] = clock64();
    if constexpr (TMASTORE) {
      fence_async_smem();                              // generic-proxy writes -> visible to the bulk-copy engine
      __syncwarp();
      if (dbg) dbg[4] = clock64();
      if (lane == 0 && store) {
        if (out_hint) tma_store_2d_hint(tm_out, stg_u, u * UW, (int)m_base, out_hint);
        else tma_store_2d(tm_out, stg_u, u * UW, (int)m_base);
        if (tm_xb != nullptr) tma_store_2d(tm_xb, stg_bu, u * UW, (int)m_base);
        tma_store_commit();
      }
      __syncwarp();
      if (dbg) dbg[5] = clock64();
      return;
    }
    __syncwarp();
    if (dbg) dbg[4] = clock64();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int rr = rr0 + RPI * i;
      const long long m = m_base + rr;
      const float4 val = *reinterpret_cast<const float4*>(stg_u + stg_off(rr, rcc));
      if (m < M && store) {
        const long long off = m * C + u * UW + rcc * 4;
        *reinterpret_cast<float4*>(x_out + off) = val;
        if (xb != nullptr) *reinterpret_cast<uint2*>(xb + off) = make_uint2(pack_bf16(val.x, val.y), pack_bf16(val.z, val.w));
      }
    }
    __syncwarp();
    if (dbg) dbg[5] = clock64();
  }
  // before the kernel exits (or the tiles are reused for something else): every bulk store has read its source
  __device__ __forceinline__ void drain_stores() {
    if constexpr (TMASTORE) {
      if (lane == 0) tma_store_wait_all();
      __syncwarp();
    }
  }
  // all of pass 2 back to back
  __device__ __forceinline__ void all_units(uint32_t y, bool store = true) {
    if constexpr (ASYNC) {
#pragma unroll 1
      for (int idx = 0; idx < NU; ++idx) {
        unit<0>(y, idx, store);
        if (dbg) dbg += 8;                             // bring-up: 8 stamps per unit
      }
      return;
    }
#pragma unroll 1
    for (int base = 0; base < NU; base += D) {
      unit<0>(y, base, store);
      if constexpr (D > 1) unit<(D > 1 ? 1 : 0)>(y, base + 1, store);
      if constexpr (D > 2) unit<(D > 2 ? 2 : 0)>(y, base + 2, store);
    }
  }
};

}  // namespace tc
}  // namespace pangu
