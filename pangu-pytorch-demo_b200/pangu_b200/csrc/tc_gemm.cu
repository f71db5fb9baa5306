// bf16 linear layers on the 5th-gen tensor cores: out = epilogue(A[M,K] . W[N,K]^T).
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor 128B-swizzled A/W tiles -> smem ring)
//   warp 1      MMA issuer     (one thread; tcgen05.mma kind::f16, fp32 accumulators in TMEM)
//   warp 2      TMEM allocator
//   warps 4..11 epilogue       (tcgen05.ld -> registers -> bias / GELU / LayerNorm+residual -> swizzled smem
//                               staging -> row-contiguous global stores; two warps per TMEM lane quarter)
// TMEM holds two accumulator tiles when 2*BN <= 512 columns, so the epilogue of tile i overlaps the
// mainloop of tile i+1.  Tiles are enumerated m-major, so the N tiles that share an A tile run on
// neighbouring CTAs at the same time and A is fetched from HBM once (the rest hits L2).
#include <cstdlib>

#include "tc_common.cuh"
#include "tc_ln_epilogue.cuh"

namespace pangu {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;                       // one 128-byte swizzle span of bf16
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB

struct GemmArgs {
  long long M;
  int K, N;
  int m_tiles, n_tiles, k_blocks;
  const float* bias;
  void* out;
  long long ldo;
  int out_dtype;
  int act;
  __nv_bfloat16* shadow;        // optional bf16 copy of an fp32 output (operand of the next GEMM), row pitch ldo
  int k_blocks_a1;              // k-blocks taken from the first A tensor; the rest come from the second (channel concat)
  const float* addend;          // optional fp32 tensor added to an fp32 output (same row pitch ldo): out = A.W^T + bias + addend
  // LayerNorm + residual epilogue
  const float* gamma;
  const float* beta;
  const float* residual;
  float* x_out;
  __nv_bfloat16* x_out_bf16;
  float eps;
  int ln_prefetch;              // $PANGU_GEMM_LN_PREFETCH=1 (experiment): pull the next row tile's residual rows into L2
};

constexpr int kEpiWarps = 8;                 // two warps per TMEM lane quarter, interleaved 32-column chunks
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int kStageEpiBytes = 4096;         // per epilogue warp: 32 rows x 128 B staging tile

// LayerNorm epilogue configuration (tc_ln_epilogue.cuh): 32-column units with 2 residual tiles in flight at
// C = 192, 16-column units with 3 in flight at C = 384 (shared memory is tighter there).  (r2 experiment: 2 bf16 staging
// tiles + fewer tiles in flight + the next row tile's residual prefetched into L2 was SLOWER, 1.54 -> 1.77 ms per step at
// C = 384: this kernel is HBM-bound and wants bytes in flight, not shorter store waits.)
template <int BN> struct LnCfg {
  static constexpr int UW = BN == 192 ? 32 : 16;
  static constexpr int D = BN == 192 ? 2 : 3;                 // residual tiles in flight per warp
  static constexpr int NB16 = 1;
  static constexpr int NBUF = D + NB16;
  static constexpr int STG = UW * 128 * NBUF;                 // fp32 staging tiles per warp
  static constexpr int STGB = UW * 64 * NB16;                 // bf16 staging tiles per warp
  static constexpr int BYTES = kEpiWarps * (STG + STGB) + 2 * 2 * 128 * 8 + 3 * BN * 4;
  static_assert(NBUF <= 4, "four load barriers per epilogue warp");
};

template <int BN, bool LN = false>
struct GemmCfg {
  static constexpr int UMMA_N = BN <= 256 ? BN : BN / 2;      // 384 -> 2 x 192
  static constexpr int N_SPLIT = BN / UMMA_N;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int EPI_BYTES = LN ? LnCfg<BN>::BYTES : kEpiWarps * kStageEpiBytes;
  static constexpr int AVAIL = 227 * 1024 - 1024 /*align slack*/ - 512 /*barriers*/ - EPI_BYTES;
  static constexpr int STAGES = AVAIL / STAGE_BYTES > 6 ? 6 : AVAIL / STAGE_BYTES;
  static constexpr int NACC = 2 * BN <= 512 ? 2 : 1;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 512;
  static_assert(UMMA_N % 16 == 0 && UMMA_N <= 256, "invalid UMMA N");
  static_assert(BN % 32 == 0, "epilogue works in 32-column chunks");
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
};

// Staging tile addressing (per warp, 4 KiB).  fp32: 32 rows x 128 B, 16-byte chunk cc of row r lives at
// chunk position cc ^ (r & 7); bf16: 32 rows x 64 B, chunk position cc ^ ((r >> 1) & 3).  Both the
// thread-per-row accesses (TMEM layout) and the row-contiguous accesses (coalesced global I/O) are
// bank-conflict free.
__device__ __forceinline__ int stg_f32(int r, int cc) { return r * 128 + ((cc ^ (r & 7)) << 4); }
__device__ __forceinline__ int stg_b16(int r, int cc) { return r * 64 + ((cc ^ ((r >> 1) & 3)) << 4); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

template <int BN, bool LN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmXb, const __grid_constant__ CUtensorMap tmRes, const GemmArgs a) {
  using Cfg = GemmCfg<BN, LN>;
  constexpr int STAGES = Cfg::STAGES, NACC = Cfg::NACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
  using LC = LnCfg<LN ? BN : 192>;
  uint8_t* stgb_smem = epi_smem + kEpiWarps * LC::STG;                                    // LN only
  float2* ln_part = reinterpret_cast<float2*>(stgb_smem + kEpiWarps * LC::STGB);         // [2 parity][2 half][128]
  float* sparams = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ln_part) + 2 * 2 * 128 * 8);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [NACC]
  uint64_t* tempty_bar = tfull_bar + NACC;      // [NACC]
  uint64_t* ln_bar = tempty_bar + NACC;         // [8 warps][4] residual tile loads of the LN epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ln_bar + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = a.m_tiles * a.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], kEpiWarps); }
    if (LN) for (int i = 0; i < 32; ++i) mbar_init(&ln_bar[i], 1);
    fence_barrier_init();
  }
  if constexpr (LN) {
    for (int i = threadIdx.x; i < 3 * BN; i += kThreads)      // affine parameters of the LN epilogue -> smem
      sparams[i] = i < BN ? a.bias[i] : (i < 2 * BN ? a.gamma[i - BN] : a.beta[i - 2 * BN]);
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmOut); tma_prefetch_desc(&tmRes); }
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                                // PDL: the set-up above overlapped the predecessor's tail; its results from here on
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (warp-uniform loop, one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / a.n_tiles) * BM, n0 = (tile % a.n_tiles) * BN;
      for (int kb = 0; kb < a.k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (kb < a.k_blocks_a1) tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
          else tma_load_2d(sa, &tmA2, &full_bar[stage], (kb - a.k_blocks_a1) * BK, m0);   // second half of cat(A1, A2)
#pragma unroll
          for (int h = 0; h < Cfg::N_SPLIT; ++h)
            tma_load_2d(sb + h * Cfg::UMMA_N * BK * 2, &tmB, &full_bar[stage], kb * BK, n0 + h * Cfg::UMMA_N);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues)
    constexpr uint32_t idesc = make_idesc_bf16(BM, Cfg::UMMA_N, 0, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tcgen05_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < a.k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_after_sync();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + A_STAGE_BYTES;
        const uint64_t da = make_desc_k_sw128(sa);
        if (elect_one()) {                       // one elected lane issues the whole k-block (uniform-register descriptors, see tc_mlp.cu)
#pragma unroll
          for (int h = 0; h < Cfg::N_SPLIT; ++h) {
            const uint64_t db = make_desc_k_sw128(sb + h * Cfg::UMMA_N * BK * 2);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)   // +32 bytes per UMMA_K=16 step inside the swizzle span
              umma_bf16(d_tmem + h * Cfg::UMMA_N, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);        // frees the smem slot once these MMAs have read it
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tfull_bar[acc]);           // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: 8 warps.
    // Warp (q, hf): TMEM lanes [32q, 32q+32) = tile rows, 32-column chunks hf, hf+2, ...  Results go
    // through a per-warp swizzled staging tile so that every global access is row-contiguous.
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int hf = (warp - 4) >> 2;
    uint8_t* stg = epi_smem + (warp - 4) * kStageEpiBytes;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t v[32];
    LnTileEpilogue<LN ? BN : 192, LC::UW, LC::D, true, true, LC::NB16, LC::NBUF> ln;
    if constexpr (LN) {
      ln.bias = a.bias; ln.gamma = a.gamma; ln.beta = a.beta; ln.residual = a.residual;
      ln.x_out = a.x_out; ln.xb = a.x_out_bf16; ln.M = a.M; ln.eps = a.eps;
      ln.stg = epi_smem + (warp - 4) * LC::STG; ln.stg_b = stgb_smem + (warp - 4) * LC::STGB;
      ln.ln_part = ln_part; ln.sparams = sparams; ln.q = q; ln.hf = hf; ln.lane = lane; ln.tile_par = 0;
      ln.tm_out = &tmOut; ln.tm_xb = a.x_out_bf16 != nullptr ? &tmXb : nullptr; ln.tm_res = &tmRes;
      ln.ld_bar = &ln_bar[(warp - 4) * 4];
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long m_base = (long long)(tile / a.n_tiles) * BM + q * 32;
      const int n0 = (tile % a.n_tiles) * BN;
      if constexpr (LN) {
        ln.m_base = m_base; ln.prefetch();                           // residual tiles fly while the MMAs finish
        // the residual rows of this CTA's NEXT row tile -> L2 (16 rows per warp), so that its tile loads are L2 hits
        const int nxt = tile + gridDim.x;
        if (a.ln_prefetch && nxt < num_tiles && lane == 0) {
          const long long r0 = (long long)(nxt / a.n_tiles) * BM + q * 32 + hf * 16;
          long long nrows = a.M - r0;
          if (nrows > 16) nrows = 16;
          if (nrows > 0) prefetch_l2_bulk(a.residual + r0 * BN, (uint32_t)(nrows * BN * 4));
        }
      }
      if (!LN && a.addend != nullptr && a.out_dtype != PANGU_BF16) {
        // dgrad + residual-gradient add: this tile's addend rows go to L2 while its (long-K) main loop runs, so the
        // loads of the write-out loop below are L2 hits (lane -> row, warp half -> half of the BN columns)
        const long long m = m_base + lane;
        if (m < a.M) prefetch_l2_bulk(a.addend + m * a.ldo + n0 + hf * (BN / 2), (BN / 2) * 4);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      if (!LN) {
#pragma unroll 1
        for (int c = hf; c < BN / 32; c += 2) {
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          const int n = n0 + c * 32;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (a.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n + j));
              f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
            }
          }
          if (a.act == PANGU_ACT_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
          }
          if (a.out_dtype == PANGU_BF16) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
              *reinterpret_cast<uint4*>(stg + stg_b16(lane, cc)) =
                  make_uint4(pack_bf16(f[8 * cc], f[8 * cc + 1]), pack_bf16(f[8 * cc + 2], f[8 * cc + 3]),
                             pack_bf16(f[8 * cc + 4], f[8 * cc + 5]), pack_bf16(f[8 * cc + 6], f[8 * cc + 7]));
            __syncwarp();
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = (lane >> 2) + 8 * i, cc = lane & 3;
              const uint4 val = *reinterpret_cast<const uint4*>(stg + stg_b16(rr, cc));
              const long long m = m_base + rr;
              if (m < a.M) *reinterpret_cast<uint4*>(out + m * a.ldo + n + cc * 8) = val;
            }
          } else {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc)
              *reinterpret_cast<float4*>(stg + stg_f32(lane, cc)) =
                  make_float4(f[4 * cc], f[4 * cc + 1], f[4 * cc + 2], f[4 * cc + 3]);
            __syncwarp();
            float* out = reinterpret_cast<float*>(a.out);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = (lane >> 3) + 4 * i, cc = lane & 7;
              float4 val = *reinterpret_cast<const float4*>(stg + stg_f32(rr, cc));
              const long long m = m_base + rr;
              if (m < a.M) {
                if (a.addend != nullptr) {
                  const float4 ad = __ldg(reinterpret_cast<const float4*>(a.addend + m * a.ldo + n + cc * 4));
                  val.x += ad.x; val.y += ad.y; val.z += ad.z; val.w += ad.w;
                }
                *reinterpret_cast<float4*>(out + m * a.ldo + n + cc * 4) = val;
                if (a.shadow != nullptr)
                  *reinterpret_cast<uint2*>(a.shadow + m * a.ldo + n + cc * 4) = make_uint2(pack_bf16(val.x, val.y), pack_bf16(val.z, val.w));
              }
            }
          }
          __syncwarp();
        }
      } else {
        // LayerNorm over the full row (BN == C) + residual (tc_ln_epilogue.cuh): residual tiles arrive by TMA
        // (issued before the accumulator was complete), results leave through TMA bulk stores.
        if constexpr (LN) {
          ln.stats(taddr);
          ln.all_units(taddr);
        }
      }
      // all of this warp's TMEM reads are done -> hand the accumulator back to the MMA warp
      tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (LN) ln.drain_stores();
  }

  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  return fn;
}

bool encode_tmap_2d_bf16(CUtensorMap* map, const void* gptr, uint64_t inner, uint64_t rows,
                         uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_rows) {
  return encode_tmap_2d(map, 1, gptr, inner, rows, row_pitch_bytes, box_inner, box_rows, 128);
}

bool encode_tmap_2d(CUtensorMap* map, int is_bf16, const void* gptr, uint64_t inner, uint64_t rows,
                    uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_rows, int swizzle_bytes) {
  PFN_tmapEncodeTiled fn = get_encode_fn();
  if (!fn) return false;
  if ((reinterpret_cast<uintptr_t>(gptr) & 15) || (row_pitch_bytes & 15)) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (ptr=%p pitch=%llu)", gptr,
              (unsigned long long)row_pitch_bytes);
    return false;
  }
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstride[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                               : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(gptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu rows=%llu pitch=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_pitch_bytes, box_inner, box_rows);
    return false;
  }
  return true;
}

// 3-D bf16 tiled map: dims {d0 (contiguous), d1, d2}, strides in bytes for d1 / d2.
bool encode_tmap_3d_bf16(CUtensorMap* map, const void* gptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                         uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle_bytes) {
  PFN_tmapEncodeTiled fn = get_encode_fn();
  if (!fn) return false;
  if ((reinterpret_cast<uintptr_t>(gptr) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15)) {
    set_error("TMA operand must be 16-byte aligned with 16-byte-multiple strides");
    return false;
  }
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                               : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(gptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed with CUresult %d (dims %llu x %llu x %llu, box %u x %u x %u)", (int)r,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1, b2);
    return false;
  }
  return true;
}

int num_sms() {
  static int per_dev[64] = {0};                      // SM count (after the optional cap) by device ordinal
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = per_dev[dev & 63];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    // $PANGU_B200_SMS=<even count>: grid cap of the persistent kernels -- leaves SMs to NCCL all-reduce CTAs that run next to
    // a DDP fine-tune step; unset = every SM
    const char* e = getenv("PANGU_B200_SMS");
    if (e != nullptr) {
      const int cap = atoi(e) & ~1;
      if (cap >= 2 && cap < n) n = cap;
    }
  }
  return n;
}

template <int BN, bool LN>
static int launch_gemm_t(const void* A, long long lda, const void* W, GemmArgs& a, cudaStream_t st,
                         const void* A2 = nullptr, long long lda2 = 0, int K1 = 0) {
  using Cfg = GemmCfg<BN, LN>;
  CUtensorMap tmA, tmA2, tmB;
  const int Ka = A2 ? K1 : a.K;                      // columns of the first A tensor
  if (!encode_tmap_2d_bf16(&tmA, A, (uint64_t)Ka, (uint64_t)a.M, (uint64_t)lda * 2, BK, BM)) return PANGU_ERR_CUDA;
  if (A2) {
    if (!encode_tmap_2d_bf16(&tmA2, A2, (uint64_t)(a.K - K1), (uint64_t)a.M, (uint64_t)lda2 * 2, BK, BM)) return PANGU_ERR_CUDA;
    a.k_blocks_a1 = K1 / BK;
  } else {
    tmA2 = tmA;
    a.k_blocks_a1 = 1 << 30;
  }
  if (!encode_tmap_2d_bf16(&tmB, W, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.K * 2, BK, Cfg::UMMA_N)) return PANGU_ERR_CUDA;
  CUtensorMap tmOut = tmA, tmXb = tmA, tmRes = tmA;
  if constexpr (LN) {
    constexpr int UW = LnCfg<BN>::UW;
    if (!a.residual) { set_error("linear_ln(bf16): a residual tensor is required"); return PANGU_ERR_BAD_ARG; }
    if (!encode_tmap_2d(&tmOut, 0, a.x_out, BN, (uint64_t)a.M, (uint64_t)BN * 4, UW, 32, UW * 4)) return PANGU_ERR_CUDA;
    if (!encode_tmap_2d(&tmRes, 0, a.residual, BN, (uint64_t)a.M, (uint64_t)BN * 4, UW, 32, UW * 4)) return PANGU_ERR_CUDA;
    if (a.x_out_bf16 && !encode_tmap_2d(&tmXb, 1, a.x_out_bf16, BN, (uint64_t)a.M, (uint64_t)BN * 2, UW, 32, UW * 2)) return PANGU_ERR_CUDA;
  }
  a.m_tiles = (int)((a.M + BM - 1) / BM);
  a.n_tiles = a.N / BN;
  a.k_blocks = (a.K + BK - 1) / BK;
  auto kern = gemm_bf16_kernel<BN, LN>;
  static unsigned long long configured = 0;
  {
    cudaError_t e = pangu::set_max_smem_once(configured, kern, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("gemm_bf16<%d>: cudaFuncSetAttribute(%d B): %s", BN, Cfg::SMEM_BYTES, cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
  }
  const int tiles = a.m_tiles * a.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  cudaError_t le = launch_pdl(kern, dim3(grid), dim3(kThreads), Cfg::SMEM_BYTES, st, tmA, tmA2, tmB, tmOut, tmXb, tmRes, a);
  if (le != cudaSuccess) { set_error("gemm_bf16: launch: %s", cudaGetErrorString(le)); return PANGU_ERR_CUDA; }
  return check_launch("gemm_bf16");
}

}  // namespace tc

int launch_tc_linear_pair(const void* A, long long lda, const void* W, const float* bias, void* out,
                          long long ldo, long long M, int K, int N, int act, int out_dtype, void* shadow, cudaStream_t st,
                          void* aux = nullptr, int aux_mode = 0, float* colsum = nullptr);

// A2 != nullptr: the A operand is the channel concat cat(A[M,K1], A2[M,K-K1]) (models/pangu_model.py:98), read
// from the two tensors directly.  shadow != nullptr (fp32 output only): also write a bf16 copy of the output.
int launch_tc_linear(const void* A, long long lda, const void* W, const float* bias, void* out,
                     long long ldo, long long M, int K, int N, int act, int out_dtype, cudaStream_t st,
                     void* shadow, const void* A2, long long lda2, int K1, const float* addend) {
  if (M == 0) return PANGU_OK;
  if (shadow && out_dtype != PANGU_F32) { set_error("linear(bf16): a bf16 shadow needs an fp32 output"); return PANGU_ERR_BAD_ARG; }
  if (A2 && (K1 <= 0 || K1 >= K || K1 % tc::BK || (K - K1) % tc::BK || lda2 % 8)) { set_error("linear(bf16): concat split K1=%d of K=%d must be a multiple of 64", K1, K); return PANGU_ERR_BAD_ARG; }
  if (K % 8 || lda % 8 || ldo % 8) { set_error("linear(bf16): K, lda, ldo must be multiples of 8 (K=%d lda=%lld ldo=%lld)", K, lda, ldo); return PANGU_ERR_BAD_ARG; }
  if (out_dtype != PANGU_BF16 && out_dtype != PANGU_F32) { set_error("linear(bf16): bad out_dtype"); return PANGU_ERR_BAD_ARG; }
  {   // skinny-K shapes: A-resident CTA-pair kernel (tc_gemm2.cu); $PANGU_B200_GEMM2=0 keeps the tiled kernel
    static const bool use_pair = []() { const char* e = getenv("PANGU_B200_GEMM2"); return e == nullptr || atoi(e) != 0; }();
    if (use_pair && !A2 && !addend) {
      const int rc = launch_tc_linear_pair(A, lda, W, bias, out, ldo, M, K, N, act, out_dtype, shadow, st);
      if (rc != PANGU_ERR_UNSUPPORTED) return rc;
    }
  }
  tc::GemmArgs a{};
  a.M = M; a.K = K; a.N = N; a.bias = bias; a.out = out; a.ldo = ldo; a.out_dtype = out_dtype; a.act = act;
  a.shadow = reinterpret_cast<__nv_bfloat16*>(shadow);
  a.addend = addend;
  if (addend && out_dtype != PANGU_F32) { set_error("linear(bf16): an addend needs an fp32 output"); return PANGU_ERR_BAD_ARG; }
  if (N % 256 == 0) return tc::launch_gemm_t<256, false>(A, lda, W, a, st, A2, lda2, K1);
  if (N % 192 == 0) return tc::launch_gemm_t<192, false>(A, lda, W, a, st, A2, lda2, K1);
  if (N % 160 == 0) return tc::launch_gemm_t<160, false>(A, lda, W, a, st, A2, lda2, K1);
  if (N % 64 == 0) return tc::launch_gemm_t<64, false>(A, lda, W, a, st, A2, lda2, K1);
  set_error("linear(bf16): N=%d is not a multiple of 256/192/160/64", N);
  return PANGU_ERR_UNSUPPORTED;
}

int launch_tc_linear_ln(const void* A, long long lda, const void* W, const float* bias,
                        const float* gamma, const float* beta, const float* residual, float* x_out,
                        void* x_out_bf16, long long M, int K, int C, float eps, cudaStream_t st) {
  if (M == 0) return PANGU_OK;
  if (K % 8 || lda % 8) { set_error("linear_ln(bf16): K and lda must be multiples of 8"); return PANGU_ERR_BAD_ARG; }
  if (!bias) { set_error("linear_ln(bf16): bias is required"); return PANGU_ERR_BAD_ARG; }
  tc::GemmArgs a{};
  a.M = M; a.K = K; a.N = C; a.bias = bias; a.gamma = gamma; a.beta = beta; a.residual = residual;
  a.x_out = x_out; a.x_out_bf16 = reinterpret_cast<__nv_bfloat16*>(x_out_bf16); a.eps = eps;
  static const int ln_pf = []() { const char* e = getenv("PANGU_GEMM_LN_PREFETCH"); return e ? atoi(e) : 0; }();
  a.ln_prefetch = ln_pf;
  if (C == 192) return tc::launch_gemm_t<192, true>(A, lda, W, a, st);
  if (C == 384) return tc::launch_gemm_t<384, true>(A, lda, W, a, st);
  set_error("linear_ln(bf16): C=%d unsupported (192/384)", C);
  return PANGU_ERR_UNSUPPORTED;
}

}  // namespace pangu
