// A-resident CTA-pair GEMM for the skinny-K linears (QKV, up-sample, embed):  out = act(A[M,K] . W[N,K]^T + b),
// K in {192, 384}, N a multiple of 192.
//
// Why a second GEMM kernel: with K this small the plain tiled GEMM (tc_gemm.cu) re-streams A and W from L2 for
// every 128 x BN tile and saturates the L2->SM path (~40 B/clk/SM measured) long before the tensor pipe, and
// cta_group::1 SS-mode MMAs fetch (128+N)*32 B of operands per K-step at ~64 B/clk.  Here a CTA PAIR
// (cluster 2x1x1, tcgen05 cta_group::2) owns 256 rows: each CTA keeps its 128 x K slice of A resident in shared
// memory for ALL N tiles, and every W tile is split between the two CTAs (each streams half of it through a
// ring of [96 x 64] k-blocks).  Per 128 rows a CTA therefore loads A once and W/2 once: at K = 384, N = 1152
// that is 528 KB per 15.7 k MMA cycles (34 B/clk) instead of 1.44 MB.
//
// Warp roles as in tc_gemm.cu / tc_mlp.cu: warp 0 TMA producer (both CTAs, crediting the leader's barriers),
// warp 1 MMA issuer (leader), warp 2 TMEM allocator, warps 4..11 epilogue (two accumulator tiles in TMEM so
// the epilogue of N tile i overlaps the MMAs of N tile i+1; staged row-contiguous stores).
#include "tc_common.cuh"

namespace pangu {
namespace tc {

constexpr int kG2Threads = 384;
constexpr int kG2EpiWarps = 8;
constexpr int G2_BN = 192;                        // N tile (per CTA pair); each CTA holds 96 rows of W
constexpr int G2_SLOT = (G2_BN / 2) * 128;        // one [96 x 64] bf16 k-block of this CTA's W half: 12 KiB
constexpr int G2_MAX_N = 1536;                    // widest linear of the model (Mlp.linear1 at C = 384)

struct Gemm2Args {
  long long M;
  int N, n_tiles, pair_tiles;
  const float* bias;
  void* out;
  long long ldo;
  int out_dtype, act;
  __nv_bfloat16* shadow;      // optional bf16 copy of an fp32 output
  // Fine-tune fusions on a bf16 output (aux bf16 [M, N], row pitch ldo):
  //   G2_AUX_PRE_OUT : aux <- A.W^T + bias (the pre-activation the backward needs), out <- act(...) in the same pass;
  //   G2_AUX_GELU_BWD: out <- (A.W^T + bias) * GELU'(aux), colsum[n] += sum_m out[m, n] (bias gradient) -- the dgrad of
  //                    Mlp.linear2 fused with the GELU backward (models/layers.py:313).
  __nv_bfloat16* aux;
  float* colsum;
  int aux_mode;
  int tma_out;                // bf16 output without an aux tensor: the staged [32 x 32] chunks leave through TMA bulk stores
};
enum { G2_AUX_NONE = 0, G2_AUX_PRE_OUT = 1, G2_AUX_GELU_BWD = 2 };

// d/dx of gelu_fast for two values at once in packed fp16 (as gelu_fast_h2, tc_common.cuh): with t = tanh(u),
// u = x (k1 + k3 x^2 + k5 x^4):  0.5 (1 + t) + 0.5 x (1 - t^2) (k1 + 3 k3 x^2 + 5 k5 x^4).  ~13 half2 ops + one MUFU per
// PAIR; |error| ~1e-3, below the bf16 rounding of the product it scales.
__device__ __forceinline__ float2 gelu_fast_grad_h2(float xa, float xb) {
  const __half2 lim = __float2half2_rn(8.0f), hf = __float2half2_rn(0.5f);
  const __half2 xc = __hmin2(__hmax2(__floats2half2_rn(xa, xb), __hneg2(lim)), lim);
  const __half2 x2 = __hmul2(xc, xc);
  __half2 p = __hfma2(x2, __float2half2_rn(-3.51519787e-4f), __float2half2_rn(3.70056658e-2f));
  p = __hfma2(x2, p, __float2half2_rn(7.97507861e-1f));
  __half2 du = __hfma2(x2, __float2half2_rn(5.0f * -3.51519787e-4f), __float2half2_rn(3.0f * 3.70056658e-2f));
  du = __hfma2(x2, du, __float2half2_rn(7.97507861e-1f));
  const __half2 u = __hmul2(xc, p);
  uint32_t ui = *reinterpret_cast<const uint32_t*>(&u), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(ui));
  const __half2 t = *reinterpret_cast<const __half2*>(&ti);
  const __half2 s = __hfma2(__hneg2(t), t, __float2half2_rn(1.0f));
  const __half2 a = __hmul2(__hmul2(xc, hf), du);
  return __half22float2(__hfma2(a, s, __hfma2(t, hf, hf)));
}

template <int KB>                                 // K / 64
struct Gemm2Cfg {
  static constexpr int A_BYTES = KB * 16384;      // this CTA's 128 rows x K
  static constexpr int EPI_BYTES = kG2EpiWarps * 4096;
  static constexpr int BAR_BYTES = 512;
  static constexpr int BIAS_BYTES = G2_MAX_N * 4;   // the bias vector, read by the epilogue as broadcast LDS (an LDG per chunk
                                                    // was the largest epilogue stall in the ncu source page)
  static constexpr int AVAIL = 227 * 1024 - 1024 - BAR_BYTES - BIAS_BYTES - A_BYTES - EPI_BYTES;
  static constexpr int NSLOT = AVAIL / G2_SLOT > 8 ? 8 : AVAIL / G2_SLOT;
  static constexpr int SMEM_BYTES = 1024 + A_BYTES + NSLOT * G2_SLOT + EPI_BYTES + BAR_BYTES + BIAS_BYTES;
  static_assert(NSLOT >= 4, "W ring too shallow");
};

__device__ __forceinline__ int g2_stg_f32(int r, int cc) { return r * 128 + ((cc ^ (r & 7)) << 4); }
__device__ __forceinline__ int g2_stg_b16(int r, int cc) { return r * 64 + ((cc ^ ((r >> 1) & 3)) << 4); }

template <int KB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kG2Threads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmOut, const Gemm2Args a) {
  using Cfg = Gemm2Cfg<KB>;
  constexpr int NSLOT = Cfg::NSLOT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                         // [KB][128 rows x 128 B]
  uint8_t* sW = smem + Cfg::A_BYTES;                          // [NSLOT][96 rows x 128 B]
  uint8_t* epi_smem = sW + NSLOT * G2_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* t_full = bars + 2;       // [2]  accumulator complete (both CTAs)
  uint64_t* t_empty = bars + 4;      // [2]  accumulator drained  (leader's copy, 16 arrivals)
  uint64_t* w_full = bars + 6;       // [NSLOT]
  uint64_t* w_empty = bars + 6 + NSLOT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * NSLOT);
  float* s_bias = reinterpret_cast<float*>(epi_smem + Cfg::EPI_BYTES + Cfg::BAR_BYTES);
  for (int i = threadIdx.x; i < a.N; i += kG2Threads) s_bias[i] = a.bias != nullptr ? __ldg(a.bias + i) : 0.f;   // visible after the __syncthreads below

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    if (a.tma_out) tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 2 * kG2EpiWarps); }
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_cg2(tmem_slot, 512);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();
  griddep_wait();                                // PDL (tc_common.cuh): everything above overlapped the predecessor's tail
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    const uint32_t a_full_L = mapa_u32(smem_u32(a_full), 0);
    int slot = 0;
    uint32_t wphase = 0, aphase = 0;
    for (int pt = pair0; pt < a.pair_tiles; pt += npairs) {
      const int m0 = pt * 256 + (int)rank * 128;
      mbar_wait(a_empty, aphase ^ 1);
      aphase ^= 1;
      if (rank == 0 && elect_one()) mbar_expect_tx(a_full, 2 * Cfg::A_BYTES);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb)
        if (elect_one()) tma_load_2d_cg2(sA + kb * 16384, &tmA, a_full_L, kb * 64, m0);
      for (int nt = 0; nt < a.n_tiles; ++nt) {
        const int n0 = nt * G2_BN + (int)rank * (G2_BN / 2);
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&w_empty[slot], wphase ^ 1);
          if (rank == 0 && elect_one()) mbar_expect_tx(&w_full[slot], 2 * G2_SLOT);
          const uint32_t bar = mapa_u32(smem_u32(&w_full[slot]), 0);
          if (elect_one()) tma_load_2d_cg2(sW + slot * G2_SLOT, &tmW, bar, kb * 64, n0);
          if (++slot == NSLOT) { slot = 0; wphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA, warp-uniform)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, G2_BN, 0, 0);
      int slot = 0, acc = 0;
      uint32_t wphase = 0, aphase = 0, tphase = 0;           // tphase: bit acc = phase of t_empty[acc]
      for (int pt = pair0; pt < a.pair_tiles; pt += npairs) {
        mbar_wait(a_full, aphase);
        aphase ^= 1;
        for (int nt = 0; nt < a.n_tiles; ++nt) {
          mbar_wait(&t_empty[acc], ((tphase >> acc) & 1) ^ 1);
          tphase ^= 1u << acc;
          tcgen05_after_sync();
          const uint32_t d_tmem = tmem_base + acc * G2_BN;
#pragma unroll 1
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&w_full[slot], wphase);
            tcgen05_after_sync();
            const uint64_t da = make_desc_k_sw128(smem_u32(sA) + kb * 16384);
            const uint64_t db = make_desc_k_sw128(smem_u32(sW) + slot * G2_SLOT);
            if (elect_one()) {                     // one elected lane issues the whole k-block (uniform-register descriptors, see tc_mlp.cu)
#pragma unroll
              for (int k = 0; k < 4; ++k) umma2_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
              umma2_commit_mc(&w_empty[slot]);
            }
            __syncwarp();
            if (++slot == NSLOT) { slot = 0; wphase ^= 1; }
          }
          if (elect_one()) {
            umma2_commit_mc(&t_full[acc]);
            if (nt == a.n_tiles - 1) umma2_commit_mc(a_empty);   // last read of the resident A tile
          }
          __syncwarp();
          acc ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp & 3, hf = (warp - 4) >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_empty_L0 = mapa_u32(smem_u32(&t_empty[0]), 0), t_empty_L1 = mapa_u32(smem_u32(&t_empty[1]), 0);
    const uint32_t stg = smem_u32(epi_smem) + (warp - 4) * 4096;
    const uint32_t s_bias_u32 = smem_u32(s_bias);
    int acc = 0;
    int tma_half = 0;
    uint32_t tphase = 0;
    uint32_t v[32];
    // GELU_BWD: the aux (pre-activation) chunk of this warp is fetched ONE CHUNK AHEAD into registers, in the
    // row-contiguous pattern of the write-out loop (lane -> row (lane>>2)+8i, 16-byte column group lane&3).
    const bool gelu_bwd = a.aux_mode == G2_AUX_GELU_BWD;
    uint4 pre_cur[4], pre_nxt[4];
    auto load_aux = [&](int pt_, int nt_, int c_, uint4 (&dst)[4]) {
      const long long mb = (long long)pt_ * 256 + rank * 128 + q * 32;
      const int n_ = nt_ * G2_BN + c_ * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long m = mb + (lane >> 2) + 8 * i;
        dst[i] = (pt_ < a.pair_tiles && m < a.M)
                     ? __ldg(reinterpret_cast<const uint4*>(a.aux + m * a.ldo + n_ + (lane & 3) * 8)) : make_uint4(0, 0, 0, 0);
      }
    };
    // ... and the CTA's whole [128 x 192] aux tile is pulled into L2 one N tile ahead (lane -> row, warp half -> 96
    // columns = 192 B): the register prefetch then pays an L2 hit, not a DRAM round trip (16 KB in flight per SM was
    // the limit: 2.6 TB/s).
    auto prefetch_aux_tile = [&](int pt_, int nt_) {
      const long long m = (long long)pt_ * 256 + rank * 128 + q * 32 + lane;
      if (pt_ < a.pair_tiles && m < a.M) prefetch_l2_bulk(a.aux + m * a.ldo + nt_ * G2_BN + hf * (G2_BN / 2), G2_BN);
    };
    if (gelu_bwd) { prefetch_aux_tile(pair0, 0); load_aux(pair0, 0, hf, pre_cur); }
    for (int pt = pair0; pt < a.pair_tiles; pt += npairs) {
      const long long m_base = (long long)pt * 256 + rank * 128 + q * 32;
      for (int nt = 0; nt < a.n_tiles; ++nt) {
        if (gelu_bwd) {
          if (nt + 1 < a.n_tiles) prefetch_aux_tile(pt, nt + 1); else prefetch_aux_tile(pt + npairs, 0);
        }
        mbar_wait(&t_full[acc], (tphase >> acc) & 1);
        tphase ^= 1u << acc;
        tcgen05_after_sync();
        const uint32_t taddr = lane_base + acc * G2_BN;
#pragma unroll 1
        for (int c = hf; c < G2_BN / 32; c += 2) {
          if (gelu_bwd) {
            int c2 = c + 2, nt2 = nt, pt2 = pt;
            if (c2 >= G2_BN / 32) { c2 = hf; if (++nt2 == a.n_tiles) { nt2 = 0; pt2 += npairs; } }
            load_aux(pt2, nt2, c2, pre_nxt);
          }
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          const int n = nt * G2_BN + c * 32;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = lds128f(s_bias_u32 + (n + j) * 4);
            f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
          }
          if (a.aux_mode == G2_AUX_PRE_OUT) {     // the pre-activation tile, staged in the second half of the warp's 4 KiB
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
              sts128(stg + 2048 + g2_stg_b16(lane, cc), pack_bf16(f[8 * cc], f[8 * cc + 1]), pack_bf16(f[8 * cc + 2], f[8 * cc + 3]),
                     pack_bf16(f[8 * cc + 4], f[8 * cc + 5]), pack_bf16(f[8 * cc + 6], f[8 * cc + 7]));
          }
          if (a.act == PANGU_ACT_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
          }
          if (a.out_dtype == PANGU_BF16) {
            const uint32_t stg_o = stg + (a.tma_out ? tma_half * 2048 : 0);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
              sts128(stg_o + g2_stg_b16(lane, cc), pack_bf16(f[8 * cc], f[8 * cc + 1]), pack_bf16(f[8 * cc + 2], f[8 * cc + 3]),
                     pack_bf16(f[8 * cc + 4], f[8 * cc + 5]), pack_bf16(f[8 * cc + 6], f[8 * cc + 7]));
            if (a.tma_out) {
              // The staging tile IS the SWIZZLE_64B image of a [32 rows x 32 bf16] box (g2_stg_b16): one elected lane hands it
              // to the bulk-copy engine (full 64-byte row segments, rows past M clipped by the tensor map); two 2 KiB staging
              // halves alternate, so only the store issued two chunks ago has to have been read out before a half is rewritten.
              fence_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmOut, reinterpret_cast<const void*>(epi_smem + (warp - 4) * 4096 + tma_half * 2048), n, (int)m_base);
                tma_store_commit();
                tma_store_wait_read1();
              }
              tma_half ^= 1;
              __syncwarp();
              continue;
            }
            __syncwarp();
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
            if (!gelu_bwd) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rr = (lane >> 2) + 8 * i, cc = lane & 3;
                const uint4 val = lds128(stg + g2_stg_b16(rr, cc));
                const long long m = m_base + rr;
                if (m < a.M) {
                  *reinterpret_cast<uint4*>(out + m * a.ldo + n + cc * 8) = val;
                  if (a.aux_mode == G2_AUX_PRE_OUT)
                    *reinterpret_cast<uint4*>(a.aux + m * a.ldo + n + cc * 8) = lds128(stg + 2048 + g2_stg_b16(rr, cc));
                }
              }
            } else {
              float cs[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) cs[j] = 0.f;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rr = (lane >> 2) + 8 * i, cc = lane & 3;
                uint4 val = lds128(stg + g2_stg_b16(rr, cc));
                __nv_bfloat162* gp = reinterpret_cast<__nv_bfloat162*>(&val);
                const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&pre_cur[i]);
                const long long m = m_base + rr;
                if (m < a.M) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 gf = __bfloat1622float2(gp[j]), xf = __bfloat1622float2(xp[j]);
                    const float2 d = gelu_fast_grad_h2(xf.x, xf.y);
                    const float ox = gf.x * d.x, oy = gf.y * d.y;
                    gp[j] = __floats2bfloat162_rn(ox, oy);
                    cs[2 * j] += ox; cs[2 * j + 1] += oy;
                  }
                  *reinterpret_cast<uint4*>(out + m * a.ldo + n + cc * 8) = val;
                }
              }
              if (a.colsum != nullptr) {                          // 8 lanes share a column group: reduce, then 2 x red.v4
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 4);
                  cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 8);
                  cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
                }
                if (lane < 4) {
                  float* dst = a.colsum + n + lane * 8;
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(cs[0]), "f"(cs[1]), "f"(cs[2]), "f"(cs[3]) : "memory");
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(cs[4]), "f"(cs[5]), "f"(cs[6]), "f"(cs[7]) : "memory");
                }
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) pre_cur[i] = pre_nxt[i];
            }
          } else {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc)
              sts128(stg + g2_stg_f32(lane, cc), __float_as_uint(f[4 * cc]), __float_as_uint(f[4 * cc + 1]), __float_as_uint(f[4 * cc + 2]), __float_as_uint(f[4 * cc + 3]));
            __syncwarp();
            float* out = reinterpret_cast<float*>(a.out);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = (lane >> 3) + 4 * i, cc = lane & 7;
              const float4 val = lds128f(stg + g2_stg_f32(rr, cc));
              const long long m = m_base + rr;
              if (m < a.M) {
                *reinterpret_cast<float4*>(out + m * a.ldo + n + cc * 4) = val;
                if (a.shadow != nullptr)
                  *reinterpret_cast<uint2*>(a.shadow + m * a.ldo + n + cc * 4) = make_uint2(pack_bf16(val.x, val.y), pack_bf16(val.z, val.w));
              }
            }
          }
          __syncwarp();
        }
        tcgen05_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc ? t_empty_L1 : t_empty_L0);
        acc ^= 1;
      }
    }
    if (a.tma_out && lane == 0) tma_store_wait_all();          // the staging tiles must outlive the bulk stores that read them
  }

  tcgen05_before_sync();
  cluster_sync_all();
  tcgen05_after_sync();
  if (warp == 2) tmem_dealloc_cg2(tmem_base, 512);
}

template <int KB>
static int launch_gemm2_t(const void* A, long long lda, const void* W, Gemm2Args& a, cudaStream_t st) {
  using Cfg = Gemm2Cfg<KB>;
  constexpr int K = KB * 64;
  CUtensorMap tmA, tmW;
  if (!encode_tmap_2d_bf16(&tmA, A, K, (uint64_t)a.M, (uint64_t)lda * 2, 64, 128)) return PANGU_ERR_CUDA;
  if (!encode_tmap_2d_bf16(&tmW, W, K, (uint64_t)a.N, (uint64_t)K * 2, 64, G2_BN / 2)) return PANGU_ERR_CUDA;
  a.n_tiles = a.N / G2_BN;
  a.pair_tiles = (int)((a.M + 255) / 256);
  static const bool tma_out_on = []() { const char* e = getenv("PANGU_B200_GEMM2_TMA_OUT"); return e == nullptr || atoi(e) != 0; }();
  a.tma_out = (tma_out_on && a.out_dtype == PANGU_BF16 && a.aux_mode == G2_AUX_NONE && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0) ? 1 : 0;
  CUtensorMap tmOut = tmA;                                    // placeholder when the store path is off (never dereferenced)
  if (a.tma_out && !encode_tmap_2d(&tmOut, 1, a.out, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldo * 2, 32, 32, 64)) return PANGU_ERR_CUDA;
  auto kern = gemm2_bf16_kernel<KB>;
  static unsigned long long configured = 0;
  {
    cudaError_t e = pangu::set_max_smem_once(configured, kern, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("gemm2<%d>: cudaFuncSetAttribute(%d B): %s", K, Cfg::SMEM_BYTES, cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
  }
  const int max_pairs = num_sms() / 2;
  const int pairs = a.pair_tiles < max_pairs ? a.pair_tiles : max_pairs;
  cudaError_t le = launch_pdl(kern, dim3(2 * pairs), dim3(kG2Threads), Cfg::SMEM_BYTES, st, tmA, tmW, tmOut, a);
  if (le != cudaSuccess) { set_error("gemm2_bf16: launch: %s", cudaGetErrorString(le)); return PANGU_ERR_CUDA; }
  return check_launch("gemm2_bf16");
}

}  // namespace tc

// Returns PANGU_ERR_UNSUPPORTED (without touching the error string) when the shape is not one this kernel
// covers; the caller then uses the generic tiled GEMM.
int launch_tc_linear_pair(const void* A, long long lda, const void* W, const float* bias, void* out,
                          long long ldo, long long M, int K, int N, int act, int out_dtype, void* shadow, cudaStream_t st,
                          void* aux, int aux_mode, float* colsum) {
  if ((K != 192 && K != 384) || N % tc::G2_BN != 0 || N > tc::G2_MAX_N || M < 2048 || lda % 8 || ldo % 8) return PANGU_ERR_UNSUPPORTED;
  if (aux_mode != tc::G2_AUX_NONE && (aux == nullptr || out_dtype != PANGU_BF16)) { set_error("linear(bf16): aux modes need an aux tensor and a bf16 output"); return PANGU_ERR_BAD_ARG; }
  tc::Gemm2Args a{};
  a.M = M; a.N = N; a.bias = bias; a.out = out; a.ldo = ldo; a.out_dtype = out_dtype; a.act = act;
  a.shadow = reinterpret_cast<__nv_bfloat16*>(shadow);
  a.aux = reinterpret_cast<__nv_bfloat16*>(aux); a.aux_mode = aux_mode; a.colsum = colsum;
  return K == 192 ? tc::launch_gemm2_t<3>(A, lda, W, a, st) : tc::launch_gemm2_t<6>(A, lda, W, a, st);
}

}  // namespace pangu
