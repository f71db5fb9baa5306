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
// Warp roles (per CTA, 640 threads; r2: the LayerNorm epilogue has warps of its own):
//   warp 0        TMA producer (own x rows, own halves of the W1/W2 chunks; signals the LEADER's barriers); also pulls the
//                 row tile's fp32 residual rows into L2 while the MMAs run (cp.async.bulk.prefetch.L2)
//   warp 1        MMA issuer   (leader CTA only; one thread)
//   warp 2        TMEM allocator (cta_group::2, both CTAs)
//   warps 4..11   GELU: per hidden chunk  H(TMEM) -> +b1, GELU -> fp16 P(TMEM)
//   warps 12..19  LayerNorm: per row tile  Y(TMEM) -> +b2, LayerNorm, +residual -> fp32 + bf16 through TMA bulk stores
// setmaxnreg moves registers from the control warpgroup (56) to the GELU warpgroups (112); the LayerNorm groups keep 96.
//
// TMEM columns: Y [0,384) (one accumulator at C = 384, TWO at C = 192) | HP0 [384,448) | HP1 [448,512).  With two
// Y accumulators the LayerNorm warps work on row tile i while the MMA / GELU warps are on row tile i+1, so the
// epilogue's HBM traffic overlaps the MMAs.  At C = 384 TMEM is full (Y 384 + H/P 128): the epilogue is a phase of its own,
// and it is made short instead: the residual rows are already in L2, the units are staged through 4 fp32 + 2 bf16 tiles
// per warp so that loads, arithmetic and bulk stores of neighbouring units overlap (tc_ln_epilogue.cuh), and that
// staging lives in the x tile's shared memory, which is dead once the last GEMM1 of the row tile has completed.
// HP_b holds the fp32
// hidden chunk H_j (j & 1 == b); each epilogue warp overwrites the first half of ITS OWN 32 columns with
// the packed bf16 P_j, which GEMM2 then reads as its A operand.  Because the tensor pipe executes MMAs in
// issue order, G1(j+2) (which overwrites HP_b) needs no barrier against G2(j) (which reads it).
// Tensor-pipe order: G1(0) G1(1) | G2(0) G1(2) | G2(1) G1(3) | ...  -- GELU of chunk j overlaps G1(j+1).
#include <cstdlib>

#include "tc_common.cuh"
#include "tc_ln_epilogue.cuh"

#ifndef PANGU_MLP_X2
#define PANGU_MLP_X2 1                            // C = 192: double-buffered x tile (0 = the single-buffer kernel, for A/B runs)
#endif

namespace pangu {
namespace tc {

constexpr int kMlpThreads = 640;
constexpr int kMlpEpiWarps = 8;                  // GELU warps 4..11; and as many LayerNorm warps, 12..19
constexpr int kMlpLnWarp0 = 12;
#ifndef PANGU_MLP_LN_JOIN
#define PANGU_MLP_LN_JOIN 1          // 1: at C = 384 the 8 GELU warps join the 8 LayerNorm warps for the LayerNorm phase (0.373 -> 0.360 ms per launch;
                                     // 8 warps with 32-column units: 0.373; the phase stays ~15 k clocks either way: half of the SMs burst 490 KB each
                                     // at the same time, i.e. it is bound by HBM, profiles/r2_mlp_trace.md)
#endif
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
  const __nv_bfloat16* x_bf16;  // the kernel's bf16 A operand [M, C] (x, or the attention output o with PROJ): L2 prefetch of the next tile
  // PROJ variant (attention.linear2 + norm1 + shortcut fused in front, models/layers.py:481,296): the x tile is produced in
  // shared memory by a LayerNorm phase of its own instead of being loaded
  const float* bp;              // attention.linear2.bias [C]
  const float* gamma1;          // norm1 weight / bias (DropPath factor folded in by the caller)
  const float* beta1;
  const float* x_in;            // fp32 block input x [M, C]: the residual of norm1
  float eps1;
  int stagger;   // clocks by which the odd CTA pairs start late: de-phases the HBM-bound LayerNorm epilogues of the pairs
  int dbg;   // bring-up knobs ($PANGU_MLP_DBG): 2 no LN pass 2 (C=384), 32 no LN stores, 4 G1 issues 1 of 4 k-steps, 8 G2 1 of 4, 128 no L2 evict_last hint on the x1 scratch (PROJ)
};

// Bring-up timeline: with dbg bit 16 set, CTA 0 records clock64() at pipeline events of its SECOND row tile (steady state).
// Layout: [64 chunks][8 events]; events 0-3 = MMA thread (p_full passed, G2 issued, G1 issued, -),
// 4-7 = epilogue warp 4 (h_full passed, H loaded, GELU packed, P stored).  Read with pangu_debug_trace().
__device__ long long g_mlp_trace[64 * 8];

template <int C, bool PROJ = false>
struct MlpCfg {
  static constexpr int KB1 = C / 64;             // 64-wide k-blocks of GEMM1 (K = C)
  static constexpr int NCH = 4 * C / NH;         // hidden chunks per row tile
  static constexpr int NSPLIT = C / 192;         // GEMM2: N = C issued as NSPLIT MMAs of N = 192
  static constexpr int X_BYTES = 128 * C * 2;    // this CTA's rows of the x tile
  static constexpr int SLOT_BYTES = C * 64;      // half W1 chunk [32 x C] == half W2 chunk [C/2 x 64] (bf16)
  // C = 192 (no PROJ): the x tile is DOUBLE-BUFFERED -- the next row tile's operand rows land while this tile's MMAs run (the
  // single buffer could only be re-loaded after the tile's last GEMM1: 5.2 k clocks from load to landing with the tensor pipe
  // idle for ~3.7 k of the 24 k per tile, profiles/r2_mlp_trace.md).  The 48 KB come from the LayerNorm staging (16-column
  // units: 5 instead of 10 KB per warp) and one ring slot.
  static constexpr int NX = (C == 192 && !PROJ && PANGU_MLP_X2) ? 2 : 1;
  static constexpr int LN_UW = (C == 192 && NX == 1) ? 32 : 16;   // LayerNorm unit width (columns)
  static constexpr int LN_D = C == 192 ? 1 : 2;      // residual tiles in flight per LayerNorm warp (TMA loads)
  static constexpr int LN_NB16 = C == 192 ? 1 : 2;   // bf16 staging tiles: 2 = the store of unit i drains while unit i+1 is computed
  static constexpr int LN_NBUF = LN_D + LN_NB16;      // fp32 staging tiles: LN_D landing + LN_NB16 draining
  // C = 384: TMEM is full, the LayerNorm is a phase of its own during which the x tile AND the weight ring are dead, and
  // the per-unit chain (tile wait, TMEM load, arithmetic, staging, proxy fence, bulk stores) is latency-bound (~1.5 k clocks
  // per 16-column unit and warp, profiles/r2_mlp_trace.md): the 8 GELU warps JOIN the 8 LayerNorm warps (4 warps per TMEM
  // lane quarter), and all staging tiles alias x + ring (the producer waits for them before it loads the next row tile).
  // C = 192: the LayerNorm of tile i runs next to the MMAs of tile i+1 (two Y accumulators) on its own 8 warps with
  // dedicated staging.
  static constexpr bool LN_JOIN = C == 384 && PANGU_MLP_LN_JOIN;   // GELU warps join the LayerNorm phase
  static constexpr bool LN_ALIAS = C == 384;          // staging tiles alias x tile + weight ring
  static constexpr int LN_NPART = LN_JOIN ? 4 : 2;    // LayerNorm warps per TMEM lane quarter
  static constexpr int LN_WARPS = 4 * LN_NPART;
  static constexpr int STG_BYTES = LN_UW * 128 * LN_NBUF;     // per LayerNorm warp: fp32 staging tiles
  static constexpr int STGB_BYTES = LN_UW * 64 * LN_NB16;     // per LayerNorm warp: bf16 staging tiles of the TMA stores
  static constexpr int PART_BYTES = 2 * LN_NPART * 128 * 8 * (PROJ ? 2 : 1);   // LayerNorm partial sums (one set per LayerNorm)
  static constexpr int PARAM_BYTES = 3 * C * 4 * (PROJ ? 2 : 1);      // b2, gamma, beta (and bp, gamma1, beta1)
  // PROJ: the first LayerNorm stages its fp32 tiles (the x1 scratch stores) in the weight ring only -- the x tile is being
  // written with the bf16 x1 -- with one residual tile in flight and two draining
  static constexpr int LN1_D = 1, LN1_NB16 = 2, LN1_NBUF = 3;
  static constexpr int STG1_BYTES = LN_UW * 128 * LN1_NBUF;
  static constexpr int EPI_BYTES = (LN_ALIAS ? 0 : LN_WARPS * (STG_BYTES + STGB_BYTES)) + PART_BYTES + PARAM_BYTES;
  static constexpr int BAR_BYTES = 2048;
  static constexpr int AVAIL = 227 * 1024 - 1024 - BAR_BYTES - NX * X_BYTES - EPI_BYTES;
  static constexpr int NSLOT = AVAIL / SLOT_BYTES > 8 ? 8 : AVAIL / SLOT_BYTES;
  static constexpr int SMEM_BYTES = 1024 + NX * X_BYTES + NSLOT * SLOT_BYTES + EPI_BYTES + BAR_BYTES;
  static_assert(!LN_ALIAS || LN_WARPS * (STG_BYTES + STGB_BYTES) <= X_BYTES + NSLOT * SLOT_BYTES, "staging must fit in x tile + weight ring");
  static_assert(NX == 1 || !LN_ALIAS, "the double-buffered x tile is never LayerNorm staging");
  static_assert(!PROJ || (LN_JOIN && LN_WARPS * STG1_BYTES <= NSLOT * SLOT_BYTES), "PROJ: 16 LayerNorm warps, first LayerNorm staged in the weight ring");
  static constexpr int NY = C == 192 ? 2 : 1;    // output accumulators: double-buffered when TMEM has room
  static constexpr int COL_HP = 384;             // two 64-column H/P buffers behind the Y accumulator(s)
  static_assert(C % 192 == 0 && C + 128 <= 512, "C must be 192 or 384");
  static_assert(NSLOT >= 3, "weight ring too shallow");
};


// PROJ = true: tmX is the attention OUTPUT o (bf16 [M, C]); the kernel first computes o . Wp^T into the Y accumulator, a first
// LayerNorm phase turns it into x1 = x + LN1(. + bp) -- fp32 into a per-CTA scratch tile (tmRes: L2-resident, read back as the
// residual of the second LayerNorm), bf16 straight into the x tile in shared memory -- and then runs the Mlp as before:
// x1 and its bf16 shadow never reach HBM (12 of the block tail's 24 bytes per element).
template <int C, bool PROJ = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmXb, const __grid_constant__ CUtensorMap tmRes,
                 const __grid_constant__ CUtensorMap tmWp, const __grid_constant__ CUtensorMap tmRes1,
                 const MlpArgs a) {
  using Cfg = MlpCfg<C, PROJ>;
  constexpr int NSLOT = Cfg::NSLOT, NCH = Cfg::NCH, KB1 = Cfg::KB1, NSPLIT = Cfg::NSPLIT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                                         // [NX][KB1][128 rows x 128 B]   (SW128 K-major)
  uint8_t* sW = smem + Cfg::NX * Cfg::X_BYTES;                // [NSLOT][SLOT_BYTES]
  uint8_t* epi_smem = sW + NSLOT * Cfg::SLOT_BYTES;
  // LayerNorm staging: fp32 tiles of all warps, then bf16 tiles of all warps (inside the x tile at C = 384)
  uint8_t* stg_smem = Cfg::LN_ALIAS ? sX : epi_smem;
  uint8_t* stgb_smem = stg_smem + Cfg::LN_WARPS * Cfg::STG_BYTES;
  float2* ln_part = reinterpret_cast<float2*>(epi_smem + (Cfg::LN_ALIAS ? 0 : Cfg::LN_WARPS * (Cfg::STG_BYTES + Cfg::STGB_BYTES)));
  float* sparams = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ln_part) + Cfg::PART_BYTES);   // [3][C] (+ [3][C] of norm1)
  float2* ln_part1 = ln_part + 2 * Cfg::LN_NPART * 128;       // PROJ: partial sums of the first LayerNorm
  float* sparams1 = sparams + 3 * C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
  uint64_t* x_full = bars + 0;
  uint64_t* x_empty = bars + 1;      // MMA -> producer (and LayerNorm warps): every GEMM1 of the row tile has read x  (both CTAs)
  uint64_t* h_full = bars + 2;       // [2]  MMA -> GELU: H_j complete in HP[j&1]            (both CTAs)
  uint64_t* p_full = bars + 4;       // [2]  GELU -> MMA: P_j written to HP[j&1]             (leader's copy)
  uint64_t* y_full = bars + 6;       // [2]  MMA -> LayerNorm: Y accumulator complete        (both CTAs)
  uint64_t* y_empty = bars + 8;      // [2]  LayerNorm -> MMA: Y accumulator drained         (leader's copy)
  uint64_t* xs_free = bars + 10;     // LayerNorm -> producer: the staging tiles inside the x tile are free (local)
  uint64_t* x1_ready = bars + 11;    // PROJ: LayerNorm 1 -> MMA: the bf16 x1 tile is in shared memory             (leader's copy)
  uint64_t* ring_free = bars + 12;   // PROJ: LayerNorm 1 -> producer: its staging tiles inside the weight ring are free (local)
  uint64_t* w_full = bars + 13;      // [NSLOT]
  uint64_t* w_empty = bars + 13 + NSLOT;
  uint64_t* ln_bar = bars + 13 + 2 * NSLOT;      // [LN warps][8] barriers of the residual tile loads (4 per LayerNorm)
  uint64_t* x_full2 = bars + 13 + 2 * NSLOT + 8 * Cfg::LN_WARPS;   // second x buffer (NX == 2): same roles as x_full / x_empty
  uint64_t* x_empty2 = x_full2 + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15 + 2 * NSLOT + 8 * Cfg::LN_WARPS);
  static_assert((15 + 2 * 8 + 8 * Cfg::LN_WARPS) * 8 + 8 <= Cfg::BAR_BYTES, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                    // 0 = leader of the pair
  const int pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmRes);
    if (a.x_out_bf16 != nullptr) tma_prefetch_desc(&tmXb);
    if (PROJ) { tma_prefetch_desc(&tmWp); tma_prefetch_desc(&tmRes1); }
  }
  for (int i = threadIdx.x; i < 3 * C; i += kMlpThreads)      // affine parameters of the row-tile epilogue -> smem
    sparams[i] = i < C ? a.b2[i] : (i < 2 * C ? a.gamma[i - C] : a.beta[i - 2 * C]);
  if constexpr (PROJ) {
    for (int i = threadIdx.x; i < 3 * C; i += kMlpThreads)
      sparams1[i] = i < C ? a.bp[i] : (i < 2 * C ? a.gamma1[i - C] : a.beta1[i - 2 * C]);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(x_full, 1); mbar_init(x_empty, 1); mbar_init(x_full2, 1); mbar_init(x_empty2, 1);
    mbar_init(&h_full[0], 1); mbar_init(&h_full[1], 1);
    mbar_init(&p_full[0], 2 * kMlpEpiWarps); mbar_init(&p_full[1], 2 * kMlpEpiWarps);
    for (int i = 0; i < 2; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 2 * Cfg::LN_WARPS); }
    mbar_init(xs_free, Cfg::LN_WARPS);
    mbar_init(x1_ready, 2 * Cfg::LN_WARPS); mbar_init(ring_free, Cfg::LN_WARPS);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int i = 0; i < 8 * Cfg::LN_WARPS; ++i) mbar_init(&ln_bar[i], 1);
    fence_barrier_init();
  }
  cluster_sync_all();                                         // barrier inits visible to the peer CTA
  if (warp == 2) tmem_alloc_cg2(tmem_slot, 512);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();                                         // both CTAs own their TMEM before any MMA
  griddep_wait();                                             // PDL (tc_common.cuh): the set-up overlapped the predecessor's tail
  griddep_launch_dependents();
  if (a.stagger > 0 && (pair0 & 1)) {
    const long long t0 = clock64();
    while (clock64() - t0 < a.stagger) __nanosleep(200);
  }

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");      // the control warpgroup gives registers to the GELU warps
    if (warp == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs)
      const uint32_t x_full_L = mapa_u32(smem_u32(x_full), 0), x_full2_L = mapa_u32(smem_u32(x_full2), 0);
      int slot = 0;
      uint32_t wphase = 0, xphase = 0;                        // xphase: bit b = phase of x_empty (b = 0) / x_empty2 (b = 1)
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
      // x tile of row tile pt_ -> buffer xb_ (waits until every GEMM1 of the tile that used the buffer before has read it)
      auto load_x = [&](int pt_, int xb_) {
        const int m0_ = pt_ * 256 + (int)rank * 128;
        mbar_wait(xb_ ? x_empty2 : x_empty, ((xphase >> xb_) & 1) ^ 1);
        if (Cfg::LN_ALIAS) mbar_wait(xs_free, ((xphase >> xb_) & 1) ^ 1);   // ... and its LayerNorm no longer stages through x tile + ring
        xphase ^= 1u << xb_;
        if ((a.dbg & 16) && blockIdx.x == 0 && lane == 0) { const int n = (pt_ - pair0) / npairs; if (n >= 1 && n <= 2) g_mlp_trace[256 + n * 16 + 0] = clock64(); }
        if (rank == 0 && elect_one()) mbar_expect_tx(xb_ ? x_full2 : x_full, 2 * Cfg::X_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB1; ++kb)
          if (elect_one()) tma_load_2d_cg2(sX + xb_ * Cfg::X_BYTES + kb * 16384, &tmX, xb_ ? x_full2_L : x_full_L, kb * 64, m0_);
      };
      if (Cfg::NX == 2 && pair0 < a.pair_tiles) load_x(pair0, 0);
      int tn_p = 0;                                           // row tiles this pair has started
      for (int pt = pair0; pt < a.pair_tiles; pt += npairs, ++tn_p) {
        const int m0 = pt * 256 + (int)rank * 128;
        if (Cfg::NX == 1) load_x(pt, 0);
        // this row tile's residual rows -> L2, long before the LayerNorm warps ask for them (16 rows per lane 0..7)
        if (lane < 8 && !(a.dbg & 64)) {
          const long long r0 = (long long)m0 + lane * 16;
          long long nrows = a.M - r0;
          if (nrows > 16) nrows = 16;
          if (nrows > 0) prefetch_l2_bulk((PROJ ? a.x_in : a.residual) + r0 * C, (uint32_t)(nrows * C * 4));
        }
        // the NEXT row tile's operand rows -> L2 as well: its load can only be issued after this tile's LayerNorm (the x tile
        // is LayerNorm staging at C = 384), and an L2 hit shortens that exposed restart (6.8 k -> ~3 k clocks per tile)
        if (Cfg::LN_ALIAS && lane >= 8 && lane < 16 && pt + npairs < a.pair_tiles) {
          const long long r0 = (long long)(pt + npairs) * 256 + rank * 128 + (lane - 8) * 16;
          long long nrows = a.M - r0;
          if (nrows > 16) nrows = 16;
          if (nrows > 0) prefetch_l2_bulk(a.x_bf16 + r0 * C, (uint32_t)(nrows * C * 2));
        }
        if constexpr (PROJ) {
          // attention.linear2: Wp k-blocks, laid out exactly like a W2 chunk (rows [192h + 96 rank, +96), k-cols [64kb, +64))
          for (int kb = 0; kb < KB1; ++kb) {
            const uint32_t bar = acquire_slot();
            uint8_t* dst = sW + slot * Cfg::SLOT_BYTES;
#pragma unroll
            for (int h = 0; h < NSPLIT; ++h)
              if (elect_one()) tma_load_2d_cg2(dst + h * 12288, &tmWp, bar, kb * 64, h * 192 + (int)rank * 96);
            advance();
          }
          mbar_wait(ring_free, (xphase & 1) ^ 1);             // the first LayerNorm has staged through the ring: wait until it is done
        }
        load_w1(0);
        load_w1(1);
        if (Cfg::NX == 2 && pt + npairs < a.pair_tiles) load_x(pt + npairs, (tn_p + 1) & 1);   // the NEXT row tile's operand rows, a tile ahead
        for (int j = 0; j < NCH; ++j) {                       // same order as the MMA issuer consumes
          load_w2(j);
          if (j + 2 < NCH) load_w1(j + 2);
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform
      // control flow, one elected lane issues)
      if (rank == 0) {
        constexpr uint32_t idesc1 = make_idesc_bf16(256, NH, 0, 0);
        constexpr uint32_t idesc2 = make_idesc_f16(256, 192);     // P (TMEM) and W2 (smem) are fp16
        const uint32_t tHP = tmem_base + Cfg::COL_HP;
        int yb = 0;                                             // Y accumulator of the current row tile
        int slot = 0;
        uint32_t wphase = 0, xphase = 0, yphase = 0, pphase = 0;   // pphase: bit b = phase of p_full[b]
        auto advance = [&]() { if (++slot == NSLOT) { slot = 0; wphase ^= 1; } };
        uint32_t sXa = smem_u32(sX);                            // x buffer of the current row tile
        auto issue_g1 = [&](int b) {                            // HP_b = X . W1chunk^T   (K = C)
          mbar_wait(&w_full[slot], wphase);
          tcgen05_after_sync();
          const uint32_t sw = smem_u32(sW + slot * Cfg::SLOT_BYTES);
          // ONE elected lane runs the whole unrolled group behind a real branch: ptxas then keeps the descriptors in uniform
          // registers (UIADD3 + UTCHMMA back to back, ~5 instructions per MMA); an `if (elect_one())` around every single MMA
          // compiles to ~20 predicated instructions and 4-5 R2UR per MMA (r2: the issuing thread was the limiter at C = 192)
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < KB1; ++kb) {
              const uint64_t da = make_desc_k_sw128(sXa + kb * 16384);
              const uint64_t db = make_desc_k_sw128(sw + kb * 4096);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if ((a.dbg & 4) && k) break;
                umma2_bf16(tHP + b * 64, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0);
              }
            }
            umma2_commit_mc(&w_empty[slot]);
            umma2_commit_mc(&h_full[b]);
          }
          __syncwarp();
          advance();
        };
        for (int pt = pair0; pt < a.pair_tiles; pt += npairs) {
          const int tn = (pt - pair0) / npairs;
          const int xb = Cfg::NX == 2 ? (tn & 1) : 0;
          mbar_wait(xb ? x_full2 : x_full, (xphase >> xb) & 1);
          xphase ^= 1u << xb;
          sXa = smem_u32(sX) + xb * Cfg::X_BYTES;
          tcgen05_after_sync();
          const bool trt = (a.dbg & 16) && blockIdx.x == 0 && lane == 0 && tn >= 1 && tn <= 2;
          if (trt) g_mlp_trace[256 + tn * 16 + 1] = clock64();
          if constexpr (PROJ) {
            // Y = o . Wp^T (attention.linear2; K = C, A = the o tile in shared memory, SS form), then the first LayerNorm
            // replaces the o tile by the bf16 x1 tile
            constexpr uint32_t idescp = make_idesc_bf16(256, 192, 0, 0);
            mbar_wait(&y_empty[yb], ((yphase >> yb) & 1) ^ 1);  // the previous row tile's LayerNorm 2 has drained Y
            yphase ^= 1u << yb;
            tcgen05_after_sync();
#pragma unroll 1
            for (int kb = 0; kb < KB1; ++kb) {
              mbar_wait(&w_full[slot], wphase);
              tcgen05_after_sync();
              const uint32_t sw = smem_u32(sW + slot * Cfg::SLOT_BYTES);
              if (elect_one()) {
                const uint64_t da = make_desc_k_sw128(smem_u32(sX) + kb * 16384);
#pragma unroll
                for (int h = 0; h < NSPLIT; ++h) {
                  const uint64_t db = make_desc_k_sw128(sw + h * 12288);
#pragma unroll
                  for (int k = 0; k < 4; ++k) umma2_bf16(tmem_base + yb * 192 + h * 192, da + 2 * k, db + 2 * k, idescp, (kb | k) != 0);
                }
                umma2_commit_mc(&w_empty[slot]);
              }
              __syncwarp();
              advance();
            }
            if (elect_one()) umma2_commit_mc(&y_full[yb]);      // first completion of this row tile: proj done
            __syncwarp();
            if (trt) g_mlp_trace[256 + tn * 16 + 11] = clock64();
            mbar_wait(x1_ready, tn & 1);                        // bf16 x1 is in the x tile (both CTAs)
            tcgen05_after_sync();
            if (trt) g_mlp_trace[256 + tn * 16 + 12] = clock64();
          }
          issue_g1(0);
          issue_g1(1);
          for (int j = 0; j < NCH; ++j) {
            const int b = j & 1;
            mbar_wait(&p_full[b], (pphase >> b) & 1);           // GELU(H_j) written to TMEM by both CTAs
            pphase ^= 1u << b;
            const bool tr = (a.dbg & 16) && blockIdx.x == 0 && pt == pair0 + npairs && j < 64 && lane == 0;
            if (tr) g_mlp_trace[j * 8 + 0] = clock64();
            if (j == 0) { mbar_wait(&y_empty[yb], ((yphase >> yb) & 1) ^ 1); yphase ^= 1u << yb; if (trt) g_mlp_trace[256 + tn * 16 + 2] = clock64(); }   // this Y buffer drained
            mbar_wait(&w_full[slot], wphase);
            tcgen05_after_sync();
            const uint32_t sw = smem_u32(sW + slot * Cfg::SLOT_BYTES);
            if (elect_one()) {
#pragma unroll
              for (int h = 0; h < NSPLIT; ++h) {
                const uint64_t db = make_desc_k_sw128(sw + h * 12288);
#pragma unroll
                for (int k = 0; k < 4; ++k)                      // Y += P_j . W2chunk^T ; K-step k lives at columns (k>>1)*32 + (k&1)*8
                  if (!((a.dbg & 8) && k)) umma2_bf16_ts(tmem_base + yb * 192 + h * 192, tHP + b * 64 + (k >> 1) * 32 + (k & 1) * 8, db + 2 * k, idesc2, (j | k) != 0);
              }
              umma2_commit_mc(&w_empty[slot]);
            }
            __syncwarp();
            advance();
            if (tr) g_mlp_trace[j * 8 + 1] = clock64();
            if (j + 2 < NCH) {
              issue_g1(b);
              if (tr) g_mlp_trace[j * 8 + 2] = clock64();                                      // H_{j+2} overwrites HP_b after G2(j) (pipe order)
              if (j + 2 == NCH - 1 && elect_one()) umma2_commit_mc(xb ? x_empty2 : x_empty);   // last read of the x tile
            }
          }
          if (elect_one()) umma2_commit_mc(&y_full[yb]);
          if (trt) g_mlp_trace[256 + tn * 16 + 3] = clock64();
          __syncwarp();
          if (Cfg::NY == 2) yb ^= 1;
        }
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ warps 4..11 GELU, warps 12..19 LayerNorm (both CTAs);
    // at C = 384 the GELU warps join the LayerNorm phase of their row tile
    const bool is_gelu = warp < kMlpLnWarp0;
    if (is_gelu) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ghf = (warp - 4) >> 2;                          // GELU: which 32 of the chunk's 64 hidden units
    const uint32_t p_full_L0 = mapa_u32(smem_u32(&p_full[0]), 0), p_full_L1 = mapa_u32(smem_u32(&p_full[1]), 0);
    uint32_t hphase = 0;                                      // bit b = phase of h_full[b]
    // LayerNorm participant?  C = 384: all 16 warps, lw = 0..15; C = 192: warps 12..19, lw = 0..7
    const bool is_ln = Cfg::LN_JOIN || !is_gelu;
    const int lw = Cfg::LN_JOIN ? warp - 4 : warp - kMlpLnWarp0;
    const uint32_t y_empty_L0 = mapa_u32(smem_u32(&y_empty[0]), 0), y_empty_L1 = mapa_u32(smem_u32(&y_empty[1]), 0);
    uint32_t yphase = 0;                                      // bit b = phase of y_full[b]
    using Ln = LnTileEpilogue<C, Cfg::LN_UW, Cfg::LN_D, true, true, Cfg::LN_NB16, Cfg::LN_NBUF, Cfg::LN_NPART>;
    static_assert(Ln::NBUF <= 4, "four load barriers per LayerNorm warp");
    Ln ln;
    ln.bias = a.b2; ln.gamma = a.gamma; ln.beta = a.beta; ln.residual = a.residual;
    ln.x_out = a.x_out; ln.xb = a.x_out_bf16; ln.M = a.M; ln.eps = a.eps;
    ln.tm_out = &tmOut; ln.tm_xb = a.x_out_bf16 != nullptr ? &tmXb : nullptr;
    ln.stg = stg_smem + (is_ln ? lw : 0) * Cfg::STG_BYTES; ln.stg_b = stgb_smem + (is_ln ? lw : 0) * Cfg::STGB_BYTES; ln.sparams = sparams;
    ln.tm_res = &tmRes; ln.ld_bar = &ln_bar[(is_ln ? lw : 0) * 8];
    ln.ln_part = ln_part; ln.q = q; ln.hf = lw >> 2; ln.lane = lane; ln.tile_par = 0;
    // PROJ: the first LayerNorm of the row tile, x1 = x + LN1(o Wp^T + bp): residual x by TMA from global memory, fp32 result
    // through the ring-resident staging tiles into this CTA's scratch rows (tmRes), bf16 result into the x tile
    using Ln1 = LnTileEpilogue<C, Cfg::LN_UW, Cfg::LN1_D, true, true, Cfg::LN1_NB16, Cfg::LN1_NBUF, Cfg::LN_NPART>;
    Ln1 ln1;
    if constexpr (PROJ) {
      ln1.bias = a.bp; ln1.gamma = a.gamma1; ln1.beta = a.beta1; ln1.residual = a.x_in;
      ln1.x_out = nullptr; ln1.xb = nullptr; ln1.M = a.M; ln1.eps = a.eps1;
      ln1.tm_out = &tmRes; ln1.tm_xb = nullptr; ln1.umma_x = sX; ln1.x_row0 = q * 32;
      ln1.stg = sW + lw * Cfg::STG1_BYTES; ln1.stg_b = nullptr; ln1.sparams = sparams1;
      ln1.tm_res = &tmRes1; ln1.ld_bar = &ln_bar[lw * 8 + 4];
      ln1.ln_part = ln_part1; ln1.q = q; ln1.hf = lw >> 2; ln1.lane = lane; ln1.tile_par = 0;
      // the x1 scratch tile (128 rows x C fp32 per CTA, 29 MB over the grid) is rewritten every row tile: pin it in the L2
      // (r2 ncu: without the hint 178 of its 201 MB per launch were evicted to DRAM by the streaming traffic and 52 MB re-read)
      if (!(a.dbg & 128)) { ln1.out_hint = l2_policy_evict_last(); ln.res_hint = ln1.out_hint; }
    }
    const uint32_t x1_ready_L = mapa_u32(smem_u32(x1_ready), 0);
    const long long scratch_row = (long long)blockIdx.x * 128 + q * 32;   // this warp's rows of the CTA's scratch tile
    const bool ln_store = !(a.dbg & 32);
    int yb = 0;
    // one row tile of GELU work: per hidden chunk  H_j (TMEM) -> +b1, GELU (packed fp16) -> P_j (TMEM, over the warp's own columns)
    auto gelu_tile = [&](int tn) {
      uint32_t v[32];
#pragma unroll 1
      for (int j = 0; j < NCH; ++j) {
        const int b = j & 1;
        float4 bb[8];                                          // b1 of this warp's 32 hidden units: fetched before the wait
        {
          const float4* b1 = reinterpret_cast<const float4*>(a.b1 + j * NH + ghf * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) bb[i] = __ldg(b1 + i);
        }
        mbar_wait(&h_full[b], (hphase >> b) & 1);
        hphase ^= 1u << b;
        tcgen05_after_sync();
        const bool tr = (a.dbg & 16) && blockIdx.x == 0 && warp == 4 && lane == 0 && tn == 1 && j < 64;
        if ((a.dbg & 16) && blockIdx.x == 0 && warp == 4 && lane == 0 && (j == 0 || j == NCH - 1) && tn >= 1 && tn <= 2)
          g_mlp_trace[256 + tn * 16 + (j == 0 ? 9 : 10)] = clock64();
        if (tr) g_mlp_trace[j * 8 + 4] = clock64();
        const uint32_t t_own = lane_base + Cfg::COL_HP + b * 64 + ghf * 32;   // this warp's 32 columns of H_j
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
        tmem_st_32x16(t_own, pk);                               // P_j over the first 16 of the warp's own columns
        tmem_st_wait();
        if (tr) g_mlp_trace[j * 8 + 7] = clock64();
        tcgen05_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(b ? p_full_L1 : p_full_L0);
      }
    };
    // PROJ: first LayerNorm of the row tile (see ln1 above)
    auto ln1_tile = [&](int pt, int tn) {
      ln1.m_base = scratch_row;                               // fp32 x1 -> scratch rows
      ln1.m_res = (long long)pt * 256 + rank * 128 + q * 32;  // residual x <- global rows
      const bool trl = (a.dbg & 16) && blockIdx.x == 0 && warp == kMlpLnWarp0 && lane == 0 && tn >= 1 && tn <= 2;
      mbar_wait(&y_full[yb], (yphase >> yb) & 1);             // o . Wp^T complete: the o tile and the ring slots are dead
      yphase ^= 1u << yb;
      tcgen05_after_sync();
      if (trl) g_mlp_trace[256 + tn * 16 + 13] = clock64();
      ln1.prefetch();
      const uint32_t y = lane_base + yb * 192 * (Cfg::NY - 1);
      ln1.stats(y);
      ln1.all_units(y, true);
      tcgen05_before_sync();                                  // all TMEM reads of Y are done: GEMM2 may accumulate into it
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(yb ? y_empty_L1 : y_empty_L0);
        tma_store_wait_read();                                // the scratch stores have read the ring-resident staging tiles
        mbar_arrive(ring_free);
        mbar_arrive_cluster(x1_ready_L);                      // (every unit fenced its x-tile writes for the async proxy)
      }
      __syncwarp();
      if (trl) g_mlp_trace[256 + tn * 16 + 14] = clock64();
    };
    // one row tile of LayerNorm work:  Y (TMEM) -> +b2, LayerNorm, + residual -> fp32 + bf16 (tc_ln_epilogue.cuh)
    auto ln_tile = [&](int pt, int tn) {
      ln.m_base = (long long)pt * 256 + rank * 128 + q * 32;
      if constexpr (PROJ) {
        ln.m_res = scratch_row;                               // residual x1 <- this CTA's scratch rows, written by ln1_tile:
        if (lane == 0) tma_store_wait_all();                  // those bulk stores (same thread) are complete, not just read
        __syncwarp();
      }
      const bool trl = (a.dbg & 16) && blockIdx.x == 0 && warp == kMlpLnWarp0 && lane == 0 && tn >= 1 && tn <= 2;
      if (!Cfg::LN_ALIAS) ln.prefetch();                      // dedicated staging: the first residual tiles can fly already
      if (trl) g_mlp_trace[256 + tn * 16 + 4] = clock64();
      mbar_wait(&y_full[yb], (yphase >> yb) & 1);             // accumulator yb complete: every MMA of the row tile has retired,
      yphase ^= 1u << yb;                                     // so at C = 384 the x tile and the weight ring are dead
      tcgen05_after_sync();
      if (trl) g_mlp_trace[256 + tn * 16 + 5] = clock64();
      if (Cfg::LN_ALIAS) ln.prefetch();                       // staging aliases x + ring (L2 hits: rows pulled in by the producer)
      const uint32_t y = lane_base + yb * 192 * (Cfg::NY - 1);
      ln.stats(y);
      if (trl) g_mlp_trace[256 + tn * 16 + 6] = clock64();
      ln.dbg = (trl && tn == 1) ? g_mlp_trace + 400 : nullptr;            // per-unit stamps of warp 12, row tile 1
      if (!(a.dbg & 2)) ln.all_units(y, ln_store);
      ln.dbg = nullptr;
      if (trl) g_mlp_trace[256 + tn * 16 + 7] = clock64();
      tcgen05_before_sync();                                  // all TMEM reads of accumulator yb are done
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(yb ? y_empty_L1 : y_empty_L0);
      if (Cfg::LN_ALIAS) {                                    // hand x tile + ring back once the bulk stores have read the staging tiles
        if (lane == 0) { tma_store_wait_read(); mbar_arrive(xs_free); }
        if (trl) g_mlp_trace[256 + tn * 16 + 8] = clock64();
        __syncwarp();
      }
      if (Cfg::NY == 2) yb ^= 1;
    };
    if constexpr (Cfg::LN_JOIN) {
      for (int pt = pair0; pt < a.pair_tiles; pt += npairs) {
        const int tn = (pt - pair0) / npairs;
        if constexpr (PROJ) ln1_tile(pt, tn);
        if (is_gelu) gelu_tile(tn);
        ln_tile(pt, tn);
      }
    } else if (is_gelu) {
      for (int pt = pair0; pt < a.pair_tiles; pt += npairs) gelu_tile((pt - pair0) / npairs);
    } else {
      for (int pt = pair0; pt < a.pair_tiles; pt += npairs) ln_tile(pt, (pt - pair0) / npairs);
    }
    if (is_ln) ln.drain_stores();
  }

  tcgen05_before_sync();
  cluster_sync_all();                                         // the peer may still be read by the leader's MMAs
  tcgen05_after_sync();
  if (warp == 2) tmem_dealloc_cg2(tmem_base, 512);
}

// PROJ: x = the attention output o; wp = attention.linear2.weight (bf16 [C, C]); scratch = fp32 [>= 128 * grid CTAs, C]
template <int C, bool PROJ = false>
static int launch_mlp_t(const void* x, const void* w1, const void* w2, MlpArgs& a, cudaStream_t st, const void* wp = nullptr,
                        float* scratch = nullptr, long long scratch_rows = 0) {
  using Cfg = MlpCfg<C, PROJ>;
  CUtensorMap tmX, tmW1, tmW2;
  if (!encode_tmap_2d_bf16(&tmX, x, C, (uint64_t)a.M, (uint64_t)C * 2, 64, 128)) return PANGU_ERR_CUDA;
  if (!encode_tmap_2d_bf16(&tmW1, w1, C, 4 * C, (uint64_t)C * 2, 64, 32)) return PANGU_ERR_CUDA;
  if (!encode_tmap_2d_bf16(&tmW2, w2, 4 * C, C, (uint64_t)4 * C * 2, 64, 96)) return PANGU_ERR_CUDA;
  CUtensorMap tmOut, tmXb;
  constexpr int UW = Cfg::LN_UW;
  if (!encode_tmap_2d(&tmOut, 0, a.x_out, C, (uint64_t)a.M, (uint64_t)C * 4, UW, 32, UW * 4)) return PANGU_ERR_CUDA;
  if (a.x_out_bf16 != nullptr) {
    if (!encode_tmap_2d(&tmXb, 1, a.x_out_bf16, C, (uint64_t)a.M, (uint64_t)C * 2, UW, 32, UW * 2)) return PANGU_ERR_CUDA;
  } else {
    tmXb = tmOut;
  }
  CUtensorMap tmRes = tmOut;
  if (a.residual != nullptr && !encode_tmap_2d(&tmRes, 0, a.residual, C, (uint64_t)a.M, (uint64_t)C * 4, UW, 32, UW * 4)) return PANGU_ERR_CUDA;
  if (a.residual == nullptr && !PROJ) { set_error("mlp_fused<%d>: a residual tensor is required", C); return PANGU_ERR_BAD_ARG; }
  a.pair_tiles = (int)((a.M + 255) / 256);
  a.x_bf16 = reinterpret_cast<const __nv_bfloat16*>(x);
  const int max_pairs = num_sms() / 2;
  const int pairs = a.pair_tiles < max_pairs ? a.pair_tiles : max_pairs;
  CUtensorMap tmWp = tmW2, tmRes1 = tmRes;
  if constexpr (PROJ) {
    if (wp == nullptr || scratch == nullptr || a.x_in == nullptr || scratch_rows < 256LL * pairs) {
      set_error("attn_proj_mlp<%d>: missing operand or scratch smaller than %lld rows", C, 256LL * pairs);
      return PANGU_ERR_BAD_ARG;
    }
    if (!encode_tmap_2d_bf16(&tmWp, wp, C, C, (uint64_t)C * 2, 64, 96)) return PANGU_ERR_CUDA;
    if (!encode_tmap_2d(&tmRes1, 0, a.x_in, C, (uint64_t)a.M, (uint64_t)C * 4, UW, 32, UW * 4)) return PANGU_ERR_CUDA;
    // the second LayerNorm's residual is x1 in the scratch tiles (row = CTA * 128 + row of the tile)
    if (!encode_tmap_2d(&tmRes, 0, scratch, C, (uint64_t)scratch_rows, (uint64_t)C * 4, UW, 32, UW * 4)) return PANGU_ERR_CUDA;
  }
  auto kern = mlp_fused_kernel<C, PROJ>;
  static unsigned long long configured = 0;
  {
    cudaError_t e = pangu::set_max_smem_once(configured, kern, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("mlp_fused<%d>: cudaFuncSetAttribute(%d B): %s", C, Cfg::SMEM_BYTES, cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
  }
  cudaError_t le = launch_pdl(kern, dim3(2 * pairs), dim3(kMlpThreads), Cfg::SMEM_BYTES, st, tmX, tmW1, tmW2, tmOut, tmXb, tmRes, tmWp, tmRes1, a);
  if (le != cudaSuccess) { set_error("mlp_fused: launch: %s", cudaGetErrorString(le)); return PANGU_ERR_CUDA; }
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
  // C = 384: TMEM has no room for a second Y accumulator, so the HBM-bound LayerNorm epilogue of a row tile is a phase of its
  // own and all CTA pairs hit it together; starting the odd pairs ~10 us late de-phases them (0.382 -> 0.372 ms measured at
  // 131 040 tokens; sweep 0 / 10 / 25 / 50 / 80 k clocks).  Only when every pair has several row tiles to walk.
  const char* stg = getenv("PANGU_MLP_STAGGER");
  const long long tiles = (M + 255) / 256;
  a.stagger = stg ? atoi(stg) : ((C == 384 && tiles >= 4LL * (tc::num_sms() / 2)) ? 20000 : 0);
  if (C == 192) return tc::launch_mlp_t<192>(x, w1, w2, a, st);
  if (C == 384) return tc::launch_mlp_t<384>(x, w1, w2, a, st);
  set_error("mlp_ln_residual(bf16): C=%d unsupported (192/384)", C);
  return PANGU_ERR_UNSUPPORTED;
}


// attention.linear2 + norm1 + shortcut + Mlp + norm2 + shortcut in ONE kernel (the tail of an EarthSpecificBlock after the
// window attention, models/layers.py:481,296-297).  C = 384 only (stage B: 12 of the 16 blocks).
int launch_tc_attn_proj_mlp(const void* o, const void* wp, const float* bp, const float* gamma1, const float* beta1,
                            const float* x_in, const void* w1, const float* b1, const void* w2, const float* b2,
                            const float* gamma2, const float* beta2, float* scratch, long long scratch_rows, float* x_out,
                            void* x_out_bf16, long long M, int C, float eps1, float eps2, cudaStream_t st) {
  if (M == 0) return PANGU_OK;
  if (C != 384) { set_error("attn_proj_mlp(bf16): C=%d unsupported (384)", C); return PANGU_ERR_UNSUPPORTED; }
  tc::MlpArgs a{};
  a.M = M; a.b1 = b1; a.b2 = b2; a.gamma = gamma2; a.beta = beta2; a.residual = nullptr;
  a.bp = bp; a.gamma1 = gamma1; a.beta1 = beta1; a.x_in = x_in; a.eps1 = eps1;
  a.x_out = x_out; a.x_out_bf16 = reinterpret_cast<__nv_bfloat16*>(x_out_bf16); a.eps = eps2;
  const char* dbg = getenv("PANGU_MLP_DBG");
  a.dbg = dbg ? atoi(dbg) : 0;
  const char* stg = getenv("PANGU_MLP_STAGGER");
  const long long tiles = (M + 255) / 256;
  a.stagger = stg ? atoi(stg) : (tiles >= 4LL * (tc::num_sms() / 2) ? 20000 : 0);
  return tc::launch_mlp_t<384, true>(o, w1, w2, a, st, wp, scratch, scratch_rows);
}

}  // namespace pangu
