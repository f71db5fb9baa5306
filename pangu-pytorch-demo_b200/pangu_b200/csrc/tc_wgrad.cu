// Weight gradients of the nn.Linear / Conv1d(k=1) layers on the 5th-gen tensor cores:
//   dW[Nout, Kin] += dY[M, Nout]^T . X[M, Kin]          (autograd of models/layers.py:88,113,312,315,419,481,522,542,566,591,608)
// The contraction runs over the TOKEN axis (M = 131 040 ... 521 280) while the output is tiny, so the kernel is a
// split-K GEMM: CTA (tile, split) owns one 128 x BN output tile and a contiguous token range, accumulates it in TMEM
// and adds it to dW with vectorised fp32 reductions (red.global.add.v4.f32) at the end.
// Both operands are read exactly as the forward/backward passes leave them -- row-major [tokens, channels] -- i.e.
// MN-major for the tensor core: TMA boxes of 64 channels x 64 tokens (128-byte swizzle) land as the canonical
// MN-major SW128 layout (8 token rows x 128 B atoms; SBO = 1 KiB between 8-token groups, LBO = 8 KiB between
// 64-channel chunks) and tcgen05.mma is issued with a_major = b_major = MN.  No transposed copies are ever made.
//   warp 0 TMA producer | warp 1 MMA issuer | warp 2 TMEM allocator | warps 4..7 epilogue (one TMEM lane quarter each)
#include "tc_common.cuh"

namespace pangu {
namespace tc {

constexpr int WG_BM = 128;                 // output rows per tile (channels of dY)
constexpr int WG_BK = 64;                  // tokens per pipeline stage
constexpr int WG_CHUNK_BYTES = 64 * WG_BK * 2;   // one 64-channel x 64-token box: 8 KiB
constexpr int WG_THREADS = 256;

template <int NB> struct WgradCfg {
  static constexpr int BN = 64 * NB;
  static constexpr int A_BYTES = 2 * WG_CHUNK_BYTES, B_BYTES = NB * WG_CHUNK_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = NB <= 2 ? 6 : 4;
  static constexpr int TMEM_COLS = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

struct WgradArgs {
  float* dw;
  long long ldw;
  int n_out, k_in;
  int n_tiles;                 // tiles along k_in
  int k_blocks;                // 64-token blocks in total
  int kb_per_split;
};

template <int NB>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_bf16_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const WgradArgs a) {
  using Cfg = WgradCfg<NB>;
  constexpr int STAGES = Cfg::STAGES, BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int m0 = (tile / a.n_tiles) * WG_BM, n0 = (tile % a.n_tiles) * BN;
  const int kb0 = split * a.kb_per_split;
  const int kb1 = min(a.k_blocks, kb0 + a.kb_per_split);
  const int nkb = kb1 - kb0;                       // > 0 by construction of the grid

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmDy); tma_prefetch_desc(&tmX); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
      uint8_t* sb = sa + Cfg::A_BYTES;
      if (elect_one()) {
        const int tok = (kb0 + i) * WG_BK;
        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        tma_load_2d(sa, &tmDy, &full_bar[stage], m0, tok);
        tma_load_2d(sa + WG_CHUNK_BYTES, &tmDy, &full_bar[stage], m0 + 64, tok);
#pragma unroll
        for (int c = 0; c < NB; ++c) tma_load_2d(sb + c * WG_CHUNK_BYTES, &tmX, &full_bar[stage], n0 + 64 * c, tok);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(WG_BM, BN, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_after_sync();
      const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
      const uint32_t sb = sa + Cfg::A_BYTES;
      if (elect_one()) {                              // one elected lane issues the whole k-block (see tc_mlp.cu)
#pragma unroll
        for (int k = 0; k < WG_BK / 16; ++k) {        // 16 tokens = two 8-token groups = 2 KiB per MMA
          const uint64_t da = make_desc_mn_sw128(sa + k * 2048, WG_CHUNK_BYTES, 1024);
          const uint64_t db = make_desc_mn_sw128(sb + k * 2048, WG_CHUNK_BYTES, 1024);
          umma_bf16(tmem_base, da, db, idesc, (i | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(tfull_bar);
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    mbar_wait(tfull_bar, 0);
    tcgen05_after_sync();
    const int row = m0 + q * 32 + lane;
    float* dst = a.dw + (long long)row * a.ldw + n0;
    uint32_t v[32];
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      if (row < a.n_out) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + c * 32 + j;
          if (n + 3 < a.k_in) {
            atomicAdd(reinterpret_cast<float4*>(dst + c * 32 + j),
                      make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n + e < a.k_in) atomicAdd(dst + c * 32 + j + e, __uint_as_float(v[j + e]));
          }
        }
      }
    }
  }
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int NB>
static int launch_wgrad_t(const CUtensorMap& tmDy, const CUtensorMap& tmX, WgradArgs& a, long long M, cudaStream_t st) {
  using Cfg = WgradCfg<NB>;
  auto kern = wgrad_bf16_kernel<NB>;
  static unsigned long long configured = 0;
  {
    cudaError_t e = pangu::set_max_smem_once(configured, kern, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("wgrad<%d>: cudaFuncSetAttribute(%d B): %s", NB, Cfg::SMEM_BYTES, cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
  }
  const int m_tiles = (a.n_out + WG_BM - 1) / WG_BM;
  a.n_tiles = (a.k_in + Cfg::BN - 1) / Cfg::BN;
  const int tiles = m_tiles * a.n_tiles;
  a.k_blocks = (int)((M + WG_BK - 1) / WG_BK);
  // enough token splits for ~2 waves of CTAs, but at least 32 k-blocks each so the reductions stay negligible
  int splits = (2 * num_sms() + tiles - 1) / tiles;
  const int max_splits = (a.k_blocks + 31) / 32;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.kb_per_split = (a.k_blocks + splits - 1) / splits;
  splits = (a.k_blocks + a.kb_per_split - 1) / a.kb_per_split;      // no empty split
  dim3 grid((unsigned)tiles, (unsigned)splits);
  kern<<<grid, WG_THREADS, Cfg::SMEM_BYTES, st>>>(tmDy, tmX, a);
  return check_launch("wgrad_bf16");
}

}  // namespace tc

// dw[n_out, k_in] (fp32, row pitch ldw) += dy[M, n_out]^T . x[M, k_in]; dy / x bf16 with row pitches ldy / ldx (elements).
int launch_tc_wgrad(const void* dy, long long ldy, const void* x, long long ldx, float* dw, long long ldw, long long M,
                    int n_out, int k_in, cudaStream_t st) {
  if (M == 0) return PANGU_OK;
  if ((n_out & 7) || (k_in & 7) || (ldy & 7) || (ldx & 7) || (ldw & 3) || (reinterpret_cast<uintptr_t>(dw) & 15)) {
    set_error("wgrad(bf16): n_out, k_in, ldy, ldx must be multiples of 8 and dw 16-byte aligned with ldw %% 4 == 0");
    return PANGU_ERR_BAD_ARG;
  }
  CUtensorMap tmDy, tmX;
  if (!tc::encode_tmap_2d_bf16(&tmDy, dy, (uint64_t)n_out, (uint64_t)M, (uint64_t)ldy * 2, 64, tc::WG_BK)) return PANGU_ERR_CUDA;
  if (!tc::encode_tmap_2d_bf16(&tmX, x, (uint64_t)k_in, (uint64_t)M, (uint64_t)ldx * 2, 64, tc::WG_BK)) return PANGU_ERR_CUDA;
  tc::WgradArgs a{};
  a.dw = dw; a.ldw = ldw; a.n_out = n_out; a.k_in = k_in;
  if (k_in <= 64) return tc::launch_wgrad_t<1>(tmDy, tmX, a, M, st);
  if (k_in <= 128) return tc::launch_wgrad_t<2>(tmDy, tmX, a, M, st);
  if (k_in % 256 == 0) return tc::launch_wgrad_t<4>(tmDy, tmX, a, M, st);
  return tc::launch_wgrad_t<3>(tmDy, tmX, a, M, st);
}

}  // namespace pangu

using namespace pangu;

extern "C" int pangu_linear_wgrad_bf16(const void* dy, int64_t ldy, const void* x, int64_t ldx, float* dw, int64_t ldw,
                                       int64_t M, int32_t n_out, int32_t k_in, void* stream) {
  if (!dy || !x || !dw || M < 0 || n_out <= 0 || k_in <= 0) { set_error("linear_wgrad: bad argument"); return PANGU_ERR_BAD_ARG; }
  return launch_tc_wgrad(dy, ldy, x, ldx, dw, ldw, M, n_out, k_in, as_stream(stream));
}
