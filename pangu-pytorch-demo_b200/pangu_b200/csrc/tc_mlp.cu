// Fused Mlp block on the 5th-gen tensor cores (models/layers.py:311-317 + :297):
//
//     x_out = residual + LayerNorm_C( GELU(x . W1^T + b1) . W2^T + b2 ) * gamma + beta
//
// The 4C-wide hidden activation never leaves the SM: it is produced 64 columns at a time into TMEM
// (GEMM1), passed through bias+GELU in registers, written back to TMEM as the bf16 A operand of GEMM2,
// and accumulated into the [128 x C] fp32 output tile that stays in TMEM for the whole row tile.
//
// CTA pair (cluster 2 x 1 x 1, tcgen05 cta_group::2): one MMA covers 256 tokens (128 per CTA); every
// weight tile is split between the two CTAs' shared memories, so each SM streams only HALF of W1/W2 per
// 128 tokens -- that halves the L2->SM operand traffic (the measured limiter of the un-fused GEMMs) and
// the shared-memory footprint (the 96 KiB x tile + a 4-deep weight ring fit at C = 384).
//
// Warp roles (per CTA, 384 threads):
//   warp 0      TMA producer (own x rows, own halves of the W1/W2 chunks; signals the LEADER's barriers)
//   warp 1      MMA issuer   (leader CTA only; one thread)
//   warp 2      TMEM allocator (cta_group::2, both CTAs)
//   warps 4..11 epilogue: per hidden chunk  H(TMEM) -> +b1, GELU -> bf16 P(TMEM);
//               per row tile  Y(TMEM) -> +b2, LayerNorm, +residual -> fp32 + bf16, row-contiguous stores
//
// TMEM columns: Y [0,C) | HP0 [C,C+64) | HP1 [C+64,C+128)   (C = 384 uses all 512).  HP_b holds the fp32
// hidden chunk H_j (j & 1 == b); each epilogue warp overwrites the first half of ITS OWN 32 columns with
// the packed bf16 P_j, which GEMM2 then reads as its A operand.  Because the tensor pipe executes MMAs in
// issue order, G1(j+2) (which overwrites HP_b) needs no barrier against G2(j) (which reads it).
// Tensor-pipe order: G1(0) G1(1) | G2(0) G1(2) | G2(1) G1(3) | ...  -- GELU of chunk j overlaps G1(j+1).
#include <cstdlib>

#include "tc_common.cuh"

namespace pangu {
namespace tc {

constexpr int kMlpThreads = 384;
constexpr int kMlpEpiWarps = 8;
constexpr int NH = 64;                           // hidden columns per chunk (per CTA pair)

struct MlpArgs {
  long long M;
  int pair_tiles;
  const float* b1;
  const float* b2;
  const float* gamma;
  const float* beta;
  const float* residual;
  float* x_out;
  __nv_bfloat16* x_out_bf16;
  float eps;
  int dbg;   // bring-up knobs ($PANGU_MLP_DBG): 2 no LN pass 2, 32 no LN stores, 64 no residual loads, 4 G1 issues 1 of 4 k-steps, 8 G2 1 of 4
};

// Bring-up timeline: with dbg bit 16 set, CTA 0 records clock64() at pipeline events of its first tile.
// Layout: [64 chunks][8 events]; events 0-3 = MMA thread (p_full passed, G2 issued, G1 issued, -),
// 4-7 = epilogue warp 4 (h_full passed, H loaded, GELU packed, P stored).  Read with pangu_debug_trace().
__device__ long long g_mlp_trace[64 * 8];

template <int C>
struct MlpCfg {
  static constexpr int KB1 = C / 64;             // 64-wide k-blocks of GEMM1 (K = C)
  static constexpr int NCH = 4 * C / NH;         // hidden chunks per row tile
  static constexpr int NSPLIT = C / 192;         // GEMM2: N = C issued as NSPLIT MMAs of N = 192
  static constexpr int X_BYTES = 128 * C * 2;    // this CTA's rows of the x tile
  static constexpr int SLOT_BYTES = C * 64;      // half W1 chunk [32 x C] == half W2 chunk [C/2 x 64] (bf16)
  static constexpr int EPI_BYTES = kMlpEpiWarps * 2048 + 2 * 2 * 128 * 8;
  static constexpr int BAR_BYTES = 512;
  static constexpr int AVAIL = 227 * 1024 - 1024 - BAR_BYTES - X_BYTES - EPI_BYTES;
  static constexpr int NSLOT = AVAIL / SLOT_BYTES > 8 ? 8 : AVAIL / SLOT_BYTES;
  static constexpr int SMEM_BYTES = 1024 + X_BYTES + NSLOT * SLOT_BYTES + EPI_BYTES + BAR_BYTES;
  static constexpr int COL_HP = C;               // two 64-column H/P buffers
  static_assert(C % 192 == 0 && C + 128 <= 512, "C must be 192 or 384");
  static_assert(NSLOT >= 3, "weight ring too shallow");
};

__device__ __forceinline__ int stg16_f32(int r, int cc) { return r * 64 + ((cc ^ ((r >> 1) & 3)) << 4); }
__device__ __forceinline__ void mlp_epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kMlpEpiWarps * 32) : "memory"); }

template <int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const MlpArgs a) {
  using Cfg = MlpCfg<C>;
  constexpr int NSLOT = Cfg::NSLOT, NCH = Cfg::NCH, KB1 = Cfg::KB1, NSPLIT = Cfg::NSPLIT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                                         // [KB1][128 rows x 128 B]   (SW128 K-major)
  uint8_t* sW = smem + Cfg::X_BYTES;                          // [NSLOT][SLOT_BYTES]
  uint8_t* epi_smem = sW + NSLOT * Cfg::SLOT_BYTES;           // 8 x 2 KiB staging + LN partial sums
  float2* ln_part = reinterpret_cast<float2*>(epi_smem + kMlpEpiWarps * 2048);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
  uint64_t* x_full = bars + 0;
  uint64_t* x_empty = bars + 1;
  uint64_t* h_full = bars + 2;       // [2]  MMA -> epilogue: H_j complete in HP[j&1]   (both CTAs)
  uint64_t* p_full = bars + 4;       // [2]  epilogue -> MMA: P_j written to HP[j&1]    (leader's copy)
  uint64_t* y_full = bars + 8;
  uint64_t* y_empty = bars + 9;
  uint64_t* w_full = bars + 10;      // [NSLOT]
  uint64_t* w_empty = bars + 10 + NSLOT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10 + 2 * NSLOT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                    // 0 = leader of the pair
  const int pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(x_full, 1); mbar_init(x_empty, 1);
    mbar_init(&h_full[0], 1); mbar_init(&h_full[1], 1);
    mbar_init(&p_full[0], 2 * kMlpEpiWarps); mbar_init(&p_full[1], 2 * kMlpEpiWarps);
    mbar_init(y_full, 1); mbar_init(y_empty, 2 * kMlpEpiWarps);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    fence_barrier_init();
  }
  cluster_sync_all();                                         // barrier inits visible to the peer CTA
  if (warp == 2) tmem_alloc_cg2(tmem_slot, 512);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();                                         // both CTAs own their TMEM before any MMA

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    {
      const uint32_t x_full_L = mapa_u32(smem_u32(x_full), 0);
      int slot = 0;
      uint32_t wphase = 0, xphase = 0;
      auto acquire_slot = [&]() -> uint32_t {
        mbar_wait(&w_empty[slot], wphase ^ 1);
        if (rank == 0 && elect_one()) mbar_expect_tx(&w_full[slot], 2 * Cfg::SLOT_BYTES);
        return mapa_u32(smem_u32(&w_full[slot]), 0);
      };
      auto advance = [&]() { if (++slot == NSLOT) { slot = 0; wphase ^= 1; } };
      auto load_w1 = [&](int j) {                             // rows [64j + 32 rank, +32) of W1, all K
        const uint32_t bar = acquire_slot();
        uint8_t* dst = sW + slot * Cfg::SLOT_BYTES;
#pragma unroll
        for (int kb = 0; kb < KB1; ++kb)
          if (elect_one()) tma_load_2d_cg2(dst + kb * 4096, &tmW1, bar, kb * 64, j * NH + (int)rank * 32);
        advance();
      };
      auto load_w2 = [&](int j) {                             // rows [192h + 96 rank, +96) of W2, k-cols [64j, +64)
        const uint32_t bar = acquire_slot();
        uint8_t* dst = sW + slot * Cfg::SLOT_BYTES;
#pragma unroll
        for (int h = 0; h < NSPLIT; ++h)
          if (elect_one()) tma_load_2d_cg2(dst + h * 12288, &tmW2, bar, j * NH, h * 192 + (int)rank * 96);
        advance();
      };
      for (int pt = pair0; pt < a.pair_tiles; pt += npairs) {
        const int m0 = pt * 256 + (int)rank * 128;
        mbar_wait(x_empty, xphase ^ 1);
        xphase ^= 1;
        if (rank == 0 && elect_one()) mbar_expect_tx(x_full, 2 * Cfg::X_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB1; ++kb)
          if (elect_one()) tma_load_2d_cg2(sX + kb * 16384, &tmX, x_full_L, kb * 64, m0);
        load_w1(0);
        load_w1(1);
        for (int j = 0; j < NCH; ++j) {                       // same order as the MMA issuer consumes
          load_w2(j);
          if (j + 2 < NCH) load_w1(j + 2);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform
    // control flow, one elected lane issues)
    if (rank == 0) {
      constexpr uint32_t idesc1 = make_idesc_bf16(256, NH, 0, 0);
      constexpr uint32_t idesc2 = make_idesc_f16(256, 192);       // P (TMEM) and W2 (smem) are fp16
      const uint32_t tY = tmem_base, tHP = tmem_base + Cfg::COL_HP;
      int slot = 0;
      uint32_t wphase = 0, xphase = 0, yphase = 0, pphase = 0;   // pphase: bit b = phase of p_full[b]
      auto advance = [&]() { if (++slot == NSLOT) { slot = 0; wphase ^= 1; } };
      auto issue_g1 = [&](int b) {                            // HP_b = X . W1chunk^T   (K = C)
        mbar_wait(&w_full[slot], wphase);
        tcgen05_after_sync();
        const uint32_t sw = smem_u32(sW + slot * Cfg::SLOT_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB1; ++kb) {
          const uint64_t da = make_desc_k_sw128(smem_u32(sX) + kb * 16384);
          const uint64_t db = make_desc_k_sw128(sw + kb * 4096);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if ((a.dbg & 4) && k) break;
            if (elect_one()) umma2_bf16(tHP + b * 64, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0);
          }
        }
        if (elect_one()) { umma2_commit_mc(&w_empty[slot]); umma2_commit_mc(&h_full[b]); }
        advance();
      };
      for (int pt = pair0; pt < a.pair_tiles; pt += npairs) {
        mbar_wait(x_full, xphase);
        xphase ^= 1;
        tcgen05_after_sync();
        issue_g1(0);
        issue_g1(1);
        for (int j = 0; j < NCH; ++j) {
          const int b = j & 1;
          mbar_wait(&p_full[b], (pphase >> b) & 1);           // GELU(H_j) written to TMEM by both CTAs
          pphase ^= 1u << b;
          const bool tr = (a.dbg & 16) && blockIdx.x == 0 && pt == pair0 && j < 64 && lane == 0;
          if (tr) g_mlp_trace[j * 8 + 0] = clock64();
          if (j == 0) { mbar_wait(y_empty, yphase ^ 1); yphase ^= 1; }   // previous tile's Y drained
          mbar_wait(&w_full[slot], wphase);
          tcgen05_after_sync();
          const uint32_t sw = smem_u32(sW + slot * Cfg::SLOT_BYTES);
#pragma unroll
          for (int h = 0; h < NSPLIT; ++h) {
            const uint64_t db = make_desc_k_sw128(sw + h * 12288);
#pragma unroll
            for (int k = 0; k < 4; ++k)                        // Y += P_j . W2chunk^T ; K-step k lives at columns (k>>1)*32 + (k&1)*8
              if (!((a.dbg & 8) && k) && elect_one()) umma2_bf16_ts(tY + h * 192, tHP + b * 64 + (k >> 1) * 32 + (k & 1) * 8, db + 2 * k, idesc2, (j | k) != 0);
          }
          if (elect_one()) umma2_commit_mc(&w_empty[slot]);
          advance();
          if (tr) g_mlp_trace[j * 8 + 1] = clock64();
          if (j + 2 < NCH) {
            issue_g1(b);
            if (tr) g_mlp_trace[j * 8 + 2] = clock64();                                      // H_{j+2} overwrites HP_b after G2(j) (pipe order)
            if (j + 2 == NCH - 1 && elect_one()) umma2_commit_mc(x_empty);   // last read of the x tile
          }
        }
        if (elect_one()) umma2_commit_mc(y_full);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp & 3, hf = (warp - 4) >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t y_empty_L = mapa_u32(smem_u32(y_empty), 0);
    const uint32_t p_full_L0 = mapa_u32(smem_u32(&p_full[0]), 0), p_full_L1 = mapa_u32(smem_u32(&p_full[1]), 0);
    uint8_t* stg = epi_smem + (warp - 4) * 2048;
    uint32_t hphase = 0, yphase = 0, tile_par = 0;            // hphase: bit b = phase of h_full[b]
    uint32_t v[32];
    for (int pt = pair0; pt < a.pair_tiles; pt += npairs) {
      const long long m_base = (long long)pt * 256 + rank * 128 + q * 32;
#pragma unroll 1
      for (int j = 0; j < NCH; ++j) {
        const int b = j & 1;
        float4 bb[8];                                          // b1 of this warp's 32 hidden units: fetched before the wait
        {
          const float4* b1 = reinterpret_cast<const float4*>(a.b1 + j * NH + hf * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) bb[i] = __ldg(b1 + i);
        }
        mbar_wait(&h_full[b], (hphase >> b) & 1);
        hphase ^= 1u << b;
        tcgen05_after_sync();
        const bool tr = (a.dbg & 16) && blockIdx.x == 0 && warp == 4 && lane == 0 && pt == pair0 && j < 64;
        if (tr) g_mlp_trace[j * 8 + 4] = clock64();
        const uint32_t t_own = lane_base + Cfg::COL_HP + b * 64 + hf * 32;   // this warp's 32 columns of H_j
        tmem_ld_32x32(t_own, v);
        tmem_ld_wait();
        if (tr) g_mlp_trace[j * 8 + 5] = clock64();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pk[2 * i] = gelu_fast_h2(__uint_as_float(v[4 * i]) + bb[i].x, __uint_as_float(v[4 * i + 1]) + bb[i].y);
          pk[2 * i + 1] = gelu_fast_h2(__uint_as_float(v[4 * i + 2]) + bb[i].z, __uint_as_float(v[4 * i + 3]) + bb[i].w);
        }
        if (tr) g_mlp_trace[j * 8 + 6] = clock64();
        tmem_st_32x16(t_own, pk);                             // P_j over the first 16 of the warp's own columns
        tmem_st_wait();
        if (tr) g_mlp_trace[j * 8 + 7] = clock64();
        tcgen05_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(b ? p_full_L1 : p_full_L0);
      }

      // ---- row-tile epilogue: +b2, LayerNorm, +residual, stores.  The fp32 residual is fetched
      // row-contiguously three 16-column units ahead (first fetches are issued before the accumulator
      // is even complete), so that enough bytes are in flight to use the HBM share of this SM.
      constexpr int NU = C / 32;                              // 16-column units handled by this warp
      const int rcc = lane & 3, rr0 = lane >> 2;              // row-contiguous view: 16-byte chunk rcc of rows rr0 + 8i
      float4 rg[3][4];
      auto load_residual = [&](int idx, float4 (&dst)[4]) {
        const int u = hf + 2 * idx;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long long m = m_base + rr0 + 8 * i;
          dst[i] = (a.residual != nullptr && m < a.M && !(a.dbg & 64))
                       ? __ldg(reinterpret_cast<const float4*>(a.residual + m * C + u * 16 + rcc * 4))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_residual(0, rg[0]);
      load_residual(1, rg[1]);
      load_residual(2, rg[2]);

      mbar_wait(y_full, yphase);
      yphase ^= 1;
      tcgen05_after_sync();
      float s = 0.f, ss = 0.f;
#pragma unroll 1
      for (int c = hf; c < C / 32; c += 2) {
        tmem_ld_32x32(lane_base + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(a.b2 + c * 32 + i));
          const float x0 = __uint_as_float(v[i]) + b.x, x1 = __uint_as_float(v[i + 1]) + b.y;
          const float x2 = __uint_as_float(v[i + 2]) + b.z, x3 = __uint_as_float(v[i + 3]) + b.w;
          s += (x0 + x1) + (x2 + x3);
          ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss); ss = fmaf(x2, x2, ss); ss = fmaf(x3, x3, ss);
        }
      }
      float2* part = ln_part + tile_par * 256;
      part[hf * 128 + q * 32 + lane] = make_float2(s, ss);
      mlp_epi_bar_sync();
      const float2 p0 = part[q * 32 + lane], p1 = part[128 + q * 32 + lane];
      const float mean = (p0.x + p1.x) * (1.0f / C);
      const float var = fmaxf((p0.y + p1.y) * (1.0f / C) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + a.eps);
      tile_par ^= 1;

      auto ln_unit = [&](int idx, float4 (&cur)[4]) {
        const int u = hf + 2 * idx;
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(stg + stg16_f32(rr0 + 8 * i, rcc)) = cur[i];
        __syncwarp();
        if (idx + 3 < NU) load_residual(idx + 3, cur);
        uint32_t w[16];
        tmem_ld_32x16(lane_base + u * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int i = cc * 4;
          float4* p = reinterpret_cast<float4*>(stg + stg16_f32(lane, cc));
          float4 r = *p;
          const float4 b = __ldg(reinterpret_cast<const float4*>(a.b2 + u * 16 + i));
          const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma + u * 16 + i));
          const float4 be = __ldg(reinterpret_cast<const float4*>(a.beta + u * 16 + i));
          r.x += fmaf((__uint_as_float(w[i]) + b.x - mean) * rstd, g.x, be.x);
          r.y += fmaf((__uint_as_float(w[i + 1]) + b.y - mean) * rstd, g.y, be.y);
          r.z += fmaf((__uint_as_float(w[i + 2]) + b.z - mean) * rstd, g.z, be.z);
          r.w += fmaf((__uint_as_float(w[i + 3]) + b.w - mean) * rstd, g.w, be.w);
          *p = r;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = rr0 + 8 * i;
          const long long m = m_base + rr;
          const float4 val = *reinterpret_cast<const float4*>(stg + stg16_f32(rr, rcc));
          if (m < a.M && !(a.dbg & 32)) {
            const long long off = m * C + u * 16 + rcc * 4;
            *reinterpret_cast<float4*>(a.x_out + off) = val;
            if (a.x_out_bf16 != nullptr)
              *reinterpret_cast<uint2*>(a.x_out_bf16 + off) = make_uint2(pack_bf16(val.x, val.y), pack_bf16(val.z, val.w));
          }
        }
        __syncwarp();
      };
      static_assert(NU % 3 == 0, "unit loop is unrolled by the prefetch depth");
      if (!(a.dbg & 2)) {
#pragma unroll 1
        for (int base = 0; base < NU; base += 3) {
          ln_unit(base, rg[0]);
          ln_unit(base + 1, rg[1]);
          ln_unit(base + 2, rg[2]);
        }
      }
      tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(y_empty_L);
    }
  }

  tcgen05_before_sync();
  cluster_sync_all();                                         // the peer may still be read by the leader's MMAs
  tcgen05_after_sync();
  if (warp == 2) tmem_dealloc_cg2(tmem_base, 512);
}

template <int C>
static int launch_mlp_t(const void* x, const void* w1, const void* w2, MlpArgs& a, cudaStream_t st) {
  using Cfg = MlpCfg<C>;
  CUtensorMap tmX, tmW1, tmW2;
  if (!encode_tmap_2d_bf16(&tmX, x, C, (uint64_t)a.M, (uint64_t)C * 2, 64, 128)) return PANGU_ERR_CUDA;
  if (!encode_tmap_2d_bf16(&tmW1, w1, C, 4 * C, (uint64_t)C * 2, 64, 32)) return PANGU_ERR_CUDA;
  if (!encode_tmap_2d_bf16(&tmW2, w2, 4 * C, C, (uint64_t)4 * C * 2, 64, 96)) return PANGU_ERR_CUDA;
  a.pair_tiles = (int)((a.M + 255) / 256);
  auto kern = mlp_fused_kernel<C>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("mlp_fused<%d>: cudaFuncSetAttribute(%d B): %s", C, Cfg::SMEM_BYTES, cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
    configured = true;
  }
  const int max_pairs = num_sms() / 2;
  const int pairs = a.pair_tiles < max_pairs ? a.pair_tiles : max_pairs;
  kern<<<2 * pairs, kMlpThreads, Cfg::SMEM_BYTES, st>>>(tmX, tmW1, tmW2, a);
  return check_launch("mlp_fused");
}

}  // namespace tc

int debug_read_mlp_trace(long long* out, int n) {
  if (n > 64 * 8) n = 64 * 8;
  cudaError_t e = cudaMemcpyFromSymbol(out, tc::g_mlp_trace, sizeof(long long) * n);
  if (e != cudaSuccess) { set_error("debug trace: %s", cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
  return PANGU_OK;
}

int launch_tc_mlp(const void* x, const void* w1, const float* b1, const void* w2, const float* b2,
                  const float* gamma, const float* beta, const float* residual, float* x_out,
                  void* x_out_bf16, long long M, int C, float eps, cudaStream_t st) {
  if (M == 0) return PANGU_OK;
  tc::MlpArgs a{};
  a.M = M; a.b1 = b1; a.b2 = b2; a.gamma = gamma; a.beta = beta; a.residual = residual;
  a.x_out = x_out; a.x_out_bf16 = reinterpret_cast<__nv_bfloat16*>(x_out_bf16); a.eps = eps;
  const char* dbg = getenv("PANGU_MLP_DBG");
  a.dbg = dbg ? atoi(dbg) : 0;
  if (C == 192) return tc::launch_mlp_t<192>(x, w1, w2, a, st);
  if (C == 384) return tc::launch_mlp_t<384>(x, w1, w2, a, st);
  set_error("mlp_ln_residual(bf16): C=%d unsupported (192/384)", C);
  return PANGU_ERR_UNSUPPORTED;
}

}  // namespace pangu
