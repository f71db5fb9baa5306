// bf16 linear layers on the 5th-gen tensor cores: out = epilogue(A[M,K] . W[N,K]^T).
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor 128B-swizzled A/W tiles -> smem ring)
//   warp 1      MMA issuer     (one thread; tcgen05.mma kind::f16, fp32 accumulators in TMEM)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue       (tcgen05.ld -> registers -> bias / GELU / LayerNorm+residual -> global)
// TMEM holds two accumulator tiles when 2*BN <= 512 columns, so the epilogue of tile i overlaps the
// mainloop of tile i+1.  Tiles are enumerated m-major, so the N tiles that share an A tile run on
// neighbouring CTAs at the same time and A is fetched from HBM once (the rest hits L2).
#include "tc_common.cuh"

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
  // LayerNorm + residual epilogue
  const float* gamma;
  const float* beta;
  const float* residual;
  float* x_out;
  __nv_bfloat16* x_out_bf16;
  float eps;
};

template <int BN>
struct GemmCfg {
  static constexpr int UMMA_N = BN <= 256 ? BN : BN / 2;      // 384 -> 2 x 192
  static constexpr int N_SPLIT = BN / UMMA_N;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  static constexpr int NACC = 2 * BN <= 512 ? 2 : 1;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(UMMA_N % 16 == 0 && UMMA_N <= 256, "invalid UMMA N");
  static_assert(BN % 32 == 0, "epilogue works in 32-column chunks");
};

template <int BN, bool LN>
__global__ void __launch_bounds__(256, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmArgs a) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES, NACC = Cfg::NACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [NACC]
  uint64_t* tempty_bar = tfull_bar + NACC;      // [NACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + NACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = a.m_tiles * a.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / a.n_tiles) * BM, n0 = (tile % a.n_tiles) * BN;
        for (int kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
#pragma unroll
          for (int h = 0; h < Cfg::N_SPLIT; ++h)
            tma_load_2d(sb + h * Cfg::UMMA_N * BK * 2, &tmB, &full_bar[stage], kb * BK, n0 + h * Cfg::UMMA_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
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
#pragma unroll
          for (int h = 0; h < Cfg::N_SPLIT; ++h) {
            const uint64_t db = make_desc_k_sw128(sb + h * Cfg::UMMA_N * BK * 2);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)   // +32 bytes per UMMA_K=16 step inside the swizzle span
              umma_bf16(d_tmem + h * Cfg::UMMA_N, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);       // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);           // accumulator complete -> epilogue
        if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (128 threads, thread = row)
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t v[32];
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long m = (long long)(tile / a.n_tiles) * BM + q * 32 + lane;
      const int n0 = (tile % a.n_tiles) * BN;
      const bool valid = m < a.M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      if (!LN) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
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
          if (valid) {
            if (a.out_dtype == PANGU_BF16) {
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + m * a.ldo + n);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                dst[j] = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                    pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
            } else {
              float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + m * a.ldo + n);
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
          }
        }
      } else {
        // LayerNorm over the full row (BN == C): pass 1 mean, pass 2 centred variance, pass 3 write.
        float s = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + c * 32 + j));
            s += (__uint_as_float(v[j]) + b.x) + (__uint_as_float(v[j + 1]) + b.y) +
                 (__uint_as_float(v[j + 2]) + b.z) + (__uint_as_float(v[j + 3]) + b.w);
          }
        }
        const float mean = s * (1.0f / BN);
        float qs = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + c * 32 + j));
            const float d0 = __uint_as_float(v[j]) + b.x - mean, d1 = __uint_as_float(v[j + 1]) + b.y - mean;
            const float d2 = __uint_as_float(v[j + 2]) + b.z - mean, d3 = __uint_as_float(v[j + 3]) + b.w - mean;
            qs = fmaf(d0, d0, qs); qs = fmaf(d1, d1, qs); qs = fmaf(d2, d2, qs); qs = fmaf(d3, d3, qs);
          }
        }
        const float rstd = rsqrtf(qs * (1.0f / BN) + a.eps);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + c * 32 + j));
            const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma + c * 32 + j));
            const float4 be = __ldg(reinterpret_cast<const float4*>(a.beta + c * 32 + j));
            f[j] = fmaf((__uint_as_float(v[j]) + b.x - mean) * rstd, g.x, be.x);
            f[j + 1] = fmaf((__uint_as_float(v[j + 1]) + b.y - mean) * rstd, g.y, be.y);
            f[j + 2] = fmaf((__uint_as_float(v[j + 2]) + b.z - mean) * rstd, g.z, be.z);
            f[j + 3] = fmaf((__uint_as_float(v[j + 3]) + b.w - mean) * rstd, g.w, be.w);
          }
          if (valid) {
            const long long off = m * BN + c * 32;
            if (a.residual != nullptr) {
              const float4* r = reinterpret_cast<const float4*>(a.residual + off);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 rv = r[j];
                f[4 * j] += rv.x; f[4 * j + 1] += rv.y; f[4 * j + 2] += rv.z; f[4 * j + 3] += rv.w;
              }
            }
            float4* dst = reinterpret_cast<float4*>(a.x_out + off);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            if (a.x_out_bf16 != nullptr) {
              uint4* db = reinterpret_cast<uint4*>(a.x_out_bf16 + off);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                db[j] = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                   pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
            }
          }
        }
      }
      // all of this warp's TMEM reads are done -> hand the accumulator back to the MMA warp
      tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
    }
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
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu rows=%llu pitch=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_pitch_bytes, box_inner, box_rows);
    return false;
  }
  return true;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, bool LN>
static int launch_gemm_t(const void* A, long long lda, const void* W, GemmArgs& a, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  if (!encode_tmap_2d_bf16(&tmA, A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)lda * 2, BK, BM)) return PANGU_ERR_CUDA;
  if (!encode_tmap_2d_bf16(&tmB, W, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.K * 2, BK, Cfg::UMMA_N)) return PANGU_ERR_CUDA;
  a.m_tiles = (int)((a.M + BM - 1) / BM);
  a.n_tiles = a.N / BN;
  a.k_blocks = (a.K + BK - 1) / BK;
  auto kern = gemm_bf16_kernel<BN, LN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("gemm_bf16<%d>: cudaFuncSetAttribute(%d B): %s", BN, Cfg::SMEM_BYTES, cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
    configured = true;
  }
  const int tiles = a.m_tiles * a.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, 256, Cfg::SMEM_BYTES, st>>>(tmA, tmB, a);
  return check_launch("gemm_bf16");
}

}  // namespace tc

int launch_tc_linear(const void* A, long long lda, const void* W, const float* bias, void* out,
                     long long ldo, long long M, int K, int N, int act, int out_dtype, cudaStream_t st) {
  if (M == 0) return PANGU_OK;
  if (K % 8 || lda % 8 || ldo % 8) { set_error("linear(bf16): K, lda, ldo must be multiples of 8 (K=%d lda=%lld ldo=%lld)", K, lda, ldo); return PANGU_ERR_BAD_ARG; }
  if (out_dtype != PANGU_BF16 && out_dtype != PANGU_F32) { set_error("linear(bf16): bad out_dtype"); return PANGU_ERR_BAD_ARG; }
  tc::GemmArgs a{};
  a.M = M; a.K = K; a.N = N; a.bias = bias; a.out = out; a.ldo = ldo; a.out_dtype = out_dtype; a.act = act;
  if (N % 256 == 0) return tc::launch_gemm_t<256, false>(A, lda, W, a, st);
  if (N % 192 == 0) return tc::launch_gemm_t<192, false>(A, lda, W, a, st);
  if (N % 160 == 0) return tc::launch_gemm_t<160, false>(A, lda, W, a, st);
  if (N % 64 == 0) return tc::launch_gemm_t<64, false>(A, lda, W, a, st);
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
  if (C == 192) return tc::launch_gemm_t<192, true>(A, lda, W, a, st);
  if (C == 384) return tc::launch_gemm_t<384, true>(A, lda, W, a, st);
  set_error("linear_ln(bf16): C=%d unsupported (192/384)", C);
  return PANGU_ERR_UNSUPPORTED;
}

}  // namespace pangu
