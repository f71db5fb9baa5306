// bf16 3-D window attention on the 5th-gen tensor cores (tcgen05 / TMEM / TMA) -- the upgrade of tc_attention.cu
// (same contract: models/layers.py:422-478 with the block's pad / roll / partition / shift-mask / reverse / crop,
// :224-293, folded into the addressing; pre-scaled operands).
//
// Persistent kernel, one CTA per SM.  The SMs are split into TEAMS of `heads` CTAs: the CTAs of a team walk the same
// contiguous range of (window type, longitude window) pairs at the same time, one head each, so the 64-byte head slices
// of a qkv row are fetched together (one 128/256-byte DRAM burst serves the team) and every team gets the same number
// of windows.  Within a CTA the windows flow through a pipeline:
//   gather     TMA: a window is 12 runs of 12 longitude-consecutive tokens; one box (32 channels x 12 tokens, or a 3-D
//              box of 6 latitude rows) per q / k / v lands directly in the UMMA canonical SWIZZLE_64B layout; 6-stage
//              smem ring; zero-pad rows equal linear1's bias (layers.py:228,419) and are written by the producer warp
//              only when a stage changes window type
//   S = Q K^T  tcgen05.mma M=128 N=144 K=32 (K-major operands), fp32 accumulator in TMEM
//   softmax    TWO groups of 8 warps, group g takes windows i = g (mod 2) so that the MUFU-bound exp phase of one window
//              overlaps the FMA/ALU-bound max phase of the next.  A thread owns (score row, 80 or 64 keys) and makes
//              two passes over its TMEM columns in 16-column chunks: max of S + bias (+ mask), then exp2 -> bf16 P
//              written back over its OWN first columns (tcgen05.st), so no registers hold a whole row.  The bf16 bias
//              is added with the mixed-precision add (one FHADD.BF16 per score, no unpacking).
//   O = P V    tcgen05.mma, A = P from TMEM (TS form), B = V from smem (MN-major), N=32 K=144; every softmax group owns an S,
//              a P and an O buffer (2 x (144 + 72 + 32) = 496 TMEM columns): S(i+2) is issued the moment the group has
//              READ S(i) and completes under the group's epilogue of window i-2; S and PV are issued by different warps
//   epilogue   of window i-2 inside pass 2 of window i: 16 columns per thread -> one 32-byte sector at the un-rolled
//              token position
//   rows 128..143  (they do not fit M=128) run on four "tail" warps with mma.sync fragments on the same smem tiles,
//              one whole window per warp (online softmax over three 48-key blocks, as tc_attention.cu)
// When the window type changes, the 20 consumer warps swap the bias tile (the producer warp has already pulled it into
// L2); TMA and MMA keep running ahead meanwhile.
// Warps (768 threads, 512 TMEM columns): 0-15 softmax, 16 TMA producer, 17 S-MMA issuer + TMEM allocator, 18 PV-MMA issuer, 20-23 tails.
#include <cstdlib>

#include "attn_common.cuh"
#include "tc_common.cuh"

namespace pangu {
namespace attn2 {

using attn::ex2;
using attn::kBiasPitch;
using attn::kLog2e;
using attn::kTileBytes;
using attn::ldmatrix_x4;
using attn::ldmatrix_x4_trans;
using attn::mma_bf16;
using attn::tile_off;
using tc::pack_bf16;
using tc::smem_u32;
using tc::prefetch_l2_bulk;

constexpr int kSoftmaxWarps = 16;
constexpr int kTailWarps = 4;
constexpr int kWarpTma = 16, kWarpMma = 17, kWarpPv = 18, kWarpTail0 = 20;   // warp 19 only fills the warpgroup
constexpr int kThreads = (kWarpTail0 + kTailWarps) * 32;           // 768 = 6 warpgroups
constexpr int kConsumers = (kSoftmaxWarps + kTailWarps) * 32;      // 640 threads read the bias tile
#ifndef PANGU_ATTN_STAGES
#define PANGU_ATTN_STAGES 4
#endif
// Ring depth.  Measured (tools/gpu_exp_attn2.sh, same box): 4 stages are 5 % FASTER than 6 (0.284 / 0.322 vs 0.299 / 0.335 ms at
// C = 192, 0.168 / 0.200 vs 0.180 / 0.207 ms at C = 384): three windows are in flight between the S / softmax / PV buffers
// anyway, and a producer that runs further ahead only spreads the CTAs of a team over more windows (their 64-byte head
// slices then hit different DRAM bursts).  One stage per tail warp.
constexpr int kStages = PANGU_ATTN_STAGES;
#ifdef PANGU_ATTN_TRACE
constexpr bool kTrace = true;                      // clock64 stamps of CTA 0 (costs registers in the softmax loop)
#else
constexpr bool kTrace = false;
#endif
constexpr int kBiasBytes = 44032;                  // 144 x 152 bf16 = 43 776, padded to a multiple of 1024
constexpr int kBufBytes = 3 * kTileBytes;          // q, k, v: 27 648 = 27 x 1024
constexpr int kTmemCols = 512;
// TMEM: every softmax group g (= window parity) owns one S buffer, one P buffer and one O buffer
constexpr int kColS = 144;                         // S of group g: columns [144 g, 144 g + 144), fp32
constexpr int kColP = 288;                         // P of group g: columns [288 + 72 g, +72), bf16 pairs (keys 0..79 | 80..143)
constexpr int kColO = 432;                         // O of group g: columns [432 + 32 g, +32), fp32
constexpr int kKeys0 = 80;                         // key split between the two threads of a score row: 80 + 64
constexpr int kRunBytes = 12 * 64;                 // one run of 12 tokens x 32 channels
constexpr int kBiasTileBytes = kWinTokens * kWinTokens * 2;
// smem carve-up after the bias tile and the ring
constexpr int kOffTables = kBiasBytes + kStages * kBufBytes;
constexpr int kOffExch = kOffTables + 1024;        // [group][slot][max | sum][key half][128] floats
constexpr int kExchBytes = 2 * 2 * 2 * 2 * 128 * 4;
constexpr int kOffBars = kOffExch + kExchBytes;
constexpr int kSmemBytes = kOffBars + 256 + 1024 /*alignment slack*/;
static_assert(kSmemBytes <= 232448, "shared memory budget");

// K-major operand, 64-byte rows (32 bf16 = the whole K extent), SWIZZLE_64B: 8-row atoms of 512 B.
__device__ __forceinline__ uint64_t desc_k_sw64(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// MN-major operand (V as stored: [key][d], d contiguous = one 64-byte chunk), SWIZZLE_64B: SBO = 512 B between 8-key groups.
__device__ __forceinline__ uint64_t desc_mn_sw64(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(512 >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
// tcgen05.wait::ld that also "produces" the 16 registers of the load it completes, so that no use of them can be
// scheduled above the wait.
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :: "memory");
}
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :: "memory");
}
// s0 += lo(w), s1 += hi(w) with w = two packed bf16: mixed-precision add (FHADD.BF16 with .H0/.H1 selectors), exact.
__device__ __forceinline__ void add_bias2(uint32_t w, float& s0, float& s1) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t}"
      : "+f"(s0), "+f"(s1) : "r"(w));
}
// (s0, s1) += (b0, b1) as ONE packed fp32 instruction (sm_100 FADD2: two fp32 lanes per issue slot)
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
// bf16 pair by TRUNCATION: one full-rate PRMT instead of an F2FP (measured -3 % on the kernel together with the FADD2 adds; the
// XU-pipe share in ncu did not change, so the F2FP was not XU work -- it simply issues slower).  Against round-to-nearest the
// truncated value is low
// by u * ulp, u uniform in [0, 1): the mean, 0.5 ulp = 2^-8 * E[1 / mantissa] = 0.2818 %, is the same for every key and comes
// back as one factor on 1 / rowsum (kTruncFix); what is left is the same +-0.5 ulp noise as rounding has.
__device__ __forceinline__ uint32_t pack_bf16_trunc(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
}
constexpr float kTruncFix = 1.0028260f;
// Timing knock-outs (WRONG RESULTS; tools/ab_variant.sh ... -DPANGU_ATTN_KO=<bits>): what each part of the window loop costs.
// 1 no max pass, 2 no MUFU (exp replaced by a copy), 4 no bias reads / adds, 8 no output stores, 16 tail warps idle,
// 32 no max exchange between the two warps of a row, 64 no TMA gather (the operand tiles keep whatever they hold), 128 no
// TMEM round trip in pass 2 (no tcgen05.ld of S, no tcgen05.st of P)
#ifndef PANGU_ATTN_KO
#define PANGU_ATTN_KO 0
#endif
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct Maps {
  CUtensorMap own12, own6, hs12, hs6, hn12, hn6;           // qkv / southern halo / northern halo, boxes of 12 and 6 tokens
  CUtensorMap own3d12;                                     // qkv as [Z*rows][W][3C]: boxes of 6 rows x 12 longitudes
};

// window type of tile u of this launch (band-local enumeration: u = zw * nhw + local h-window)
__device__ __forceinline__ int tile_type(const WinGeom& g, const BandGeom& bd, int u) {
  const int zw = u / bd.nhw, hwl = u - zw * bd.nhw;
  return zw * g.nH + ((bd.wrap && hwl == bd.nhw - 1) ? g.nH - 1 : bd.hw0 + hwl);
}
// run r (= dz * 6 + dh) of window type t: token-row base inside the buffer of its class (0 pad, 1 own, 2 southern halo,
// 3 northern halo); pad (layers.py:228) + roll (:237) + partition (:253-262) in closed form
__device__ __forceinline__ void run_info(const WinGeom& g, const BandGeom& bd, int roll, int t, int r, int& rb, int& cls) {
  const int zw = t / g.nH, hw = t - zw * g.nH;
  const int dz = r / 6, dh = r - dz * 6;
  int z = 2 * zw + dz, h = 6 * hw + dh;
  if (roll == 1) { z += 1; if (z >= g.Z) z -= g.Z; h += 3; if (h >= g.Hp) h -= g.Hp; }
  rb = -1; cls = 0;
  if (roll == 2) { rb = t * kWinTokens + 12 * r; cls = 1; }
  else if (h < g.H) {
    const int hl = h - bd.h0;
    if (hl >= 0 && hl < bd.hrows) { rb = (z * bd.hrows + hl) * g.W; cls = 1; }
    else if (hl >= bd.hrows && hl < bd.hrows + bd.halo) { rb = (z * bd.halo + (hl - bd.hrows)) * g.W; cls = 2; }
    else if (hl < 0 && hl >= -bd.halo_lo) { rb = (z * bd.halo_lo + (hl + bd.halo_lo)) * g.W; cls = 3; }
  }
}
__device__ __forceinline__ long long run_token(const WinGeom& g, int roll, int l, int rb, int dw) {
  if (roll == 2) return (long long)l * g.T * kWinTokens + rb + dw;
  int w = 12 * l + (roll == 1 ? 6 : 0) + dw;
  if (w >= g.W) w -= g.W;
  return (long long)rb + w;
}

// Everything the bias-tile swap needs, kept in shared memory so that the (rare, out-of-line) swap does not hold
// registers of the softmax / tail loops.
struct SwapCtx {
  const __nv_bfloat16* earth_bias;
  __nv_bfloat16* s_bias;
  int* s_crun;
  int* s_ccls;
  WinGeom g;
  BandGeom bd;
  int roll, head;
};

// Bias tile + tables of tile u, cooperatively by the consumer threads (ctid = index among them).  A thread FETCHES its
// 16-byte chunks into registers before the barrier that retires the old tile (so the L2 / HBM latency overlaps the wait
// for the slowest warp) and STORES them afterwards.  The shift mask (gen_mask, layers.py:187-216: -100 where the region
// ids of query and key differ) of a masked window type is folded into the staged tile on the way, so the softmax loops
// never see it.  (bias - 144.27 rounds to about -144 in bf16: the probability underflows to 0 either way, as
// exp(-100 + s) does in the reference's fp32.)
__device__ __noinline__ void swap_tile_fn(const SwapCtx* ctx, int u, int ctid, int first) {
  constexpr int kChunksPerThread = (kWinTokens * 18 + kConsumers - 1) / kConsumers;     // 5
  const WinGeom& g = ctx->g;
  const int t = tile_type(g, ctx->bd, u);
  const int roll = ctx->roll;
  const uint4* src = reinterpret_cast<const uint4*>(ctx->earth_bias + ((long long)t * g.heads + ctx->head) * kWinTokens * kWinTokens);
  uint4 r[kChunksPerThread];
#pragma unroll
  for (int k = 0; k < kChunksPerThread; ++k) {
    const int i = ctid + k * kConsumers;
    if (i < kWinTokens * 18) r[k] = __ldg(src + i);           // the tile is 144 x 18 chunks, contiguous
  }
  if (!first) named_bar(9, kConsumers);                       // everybody has finished with the old tile
  const int zw = t / g.nH, hw = t - zw * g.nH;
  const bool masked = roll == 1 && (zw == g.nZ - 1 || hw == g.nH - 1);
  auto rgroup = [&](int rr) { return (zw == g.nZ - 1 ? 2 * (rr / 6) : 0) + ((hw == g.nH - 1 && (rr % 6) >= 3) ? 1 : 0); };
  const float mask_l2 = kMaskValue * kLog2e;
#pragma unroll
  for (int k = 0; k < kChunksPerThread; ++k) {
    const int i = ctid + k * kConsumers;
    if (i >= kWinTokens * 18) continue;
    const int row = i / 18, c = i - row * 18;
    uint4 v = r[k];
    if (masked) {
      const int gr = rgroup(row / 12);
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
        if (rgroup((8 * c + 2 * e) / 12) != gr) lo += mask_l2;
        if (rgroup((8 * c + 2 * e + 1) / 12) != gr) hi += mask_l2;
        w[e] = pack_bf16(lo, hi);
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(ctx->s_bias + row * kBiasPitch + c * 8) = v;
  }
  if (ctid >= 32 && ctid < 44) {
    int rb, cls;
    run_info(g, ctx->bd, roll, t, ctid - 32, rb, cls);
    ctx->s_crun[ctid - 32] = rb;
    ctx->s_ccls[ctid - 32] = cls;
  }
  named_bar(9, kConsumers);
}

// Bring-up aid: clock64 stamps of CTA 0 for its first windows (slot = window & 7), when $PANGU_ATTN_DBG is set.
// [slot][16]: 0 TMA issued, 1 S issued, 2 p_full passed, 3 PV issued (warps 16/17); 4 s_full passed, 5 max pass done,
// 6 max exchanged, 7 P stored, 8 epilogue(i-2) done (first warp of the window's softmax group); 10 start, 11 end (tail
// warp); 12 / 13 bias swap start / end (warp 0)
__device__ long long g_attn_trace[8 * 16];

// ---- softmax shift: per-thread bound of the row maximum of S + bias over its keys (hf = 0: keys [0, 80), hf = 1: [80, 144)).
// The shift mask (gen_mask, layers.py:187-216) separates the 144 keys of a window into at most four contiguous sub-blocks
// [0,36) [36,72) [72,108) [108,144) (dz x {dh < 3, dh >= 3}); inside one sub-block every key of a given row is either
// masked (-100 folded into the staged bias tile) or not.  Taking max(S) + max(bias) PER SUB-BLOCK therefore bounds the row
// maximum to within the spread of the un-masked bias values (r2: the single bound max(S) + max(bias row) was loose by up to
// 144 log2 units on masked window types and underflowed every exponent -- NaN rows -- when a masked key carried the largest
// score by more than ~126).  Same instruction count as the single bound.
template <int HF> __device__ __forceinline__ constexpr int key_sub(int idx) {      // local key index -> local sub-block
  return HF == 0 ? (idx < 36 ? 0 : (idx < 72 ? 1 : 2)) : (idx < 28 ? 0 : 1);
}
template <int HF>
__device__ __forceinline__ void bias_sub_max(uint32_t brow, float (&bm)[3]) {
  bm[0] = bm[1] = bm[2] = -INFINITY;
#pragma unroll
  for (int c = 0; c < (HF ? 8 : 10); ++c) {
    const uint4 bb = lds128(brow + 16 * c);
    const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = 8 * c + 2 * e;                          // compile-time after unrolling; pairs never straddle a boundary
      float& m = bm[key_sub<HF>(idx)];
      m = fmaxf(m, fmaxf(__uint_as_float(bw[e] << 16), __uint_as_float(bw[e] & 0xffff0000u)));
    }
  }
}
template <int HF>
__device__ __forceinline__ float score_bound(uint32_t tS, const float (&bm)[3]) {
  float mx[3][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}};
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tc::tmem_ld_32x32(tS + 32 * c, v);
    tmem_ld_wait32(v);
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      float& m = mx[key_sub<HF>(32 * c + e)][e & 1];
      m = fmaxf(m, __uint_as_float(v[e]));
    }
  }
  if constexpr (HF == 0) {
    uint32_t v[16];
    tc::tmem_ld_32x16(tS + 64, v);
    tmem_ld_wait16(v);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float& m = mx[key_sub<0>(64 + e)][e & 1];
      m = fmaxf(m, __uint_as_float(v[e]));
    }
  }
  float r = fmaxf(mx[0][0], mx[0][1]) + bm[0];
  r = fmaxf(r, fmaxf(mx[1][0], mx[1][1]) + bm[1]);
  if constexpr (HF == 0) r = fmaxf(r, fmaxf(mx[2][0], mx[2][1]) + bm[2]);
  return r;
}

__global__ void __launch_bounds__(kThreads, 1)
window_attention_tc_kernel(const __grid_constant__ Maps maps, const float* __restrict__ qkv_bias,
                           const __nv_bfloat16* __restrict__ earth_bias, __nv_bfloat16* __restrict__ out,
                           __nv_bfloat16* __restrict__ halo_out, WinGeom g, BandGeom bd, int roll,
                           float* __restrict__ lse, int dbg, int exact_all) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __nv_bfloat16* s_bias = reinterpret_cast<__nv_bfloat16*>(smem);                       // [144][152]
  uint8_t* s_buf = smem + kBiasBytes;                                                   // kStages x {q,k,v}
  // consumer-side tables of the current bias tile's window type
  int* s_crun = reinterpret_cast<int*>(smem + kOffTables);                              // [12] run base
  int* s_ccls = s_crun + 12;                                                            // [12] run class
  // producer-private tables
  int* s_prun = reinterpret_cast<int*>(smem + kOffTables + 384);                        // [12]
  int* s_pcls = s_prun + 12;                                                            // [12]
  int* s_stage_u = s_pcls + 12;                                                         // [kStages] tile whose pad rows the stage holds
  uint4* s_padvals = reinterpret_cast<uint4*>(smem + kOffTables + 512);                 // [3 tensors][4 chunks] linear1 bias, bf16
  SwapCtx* s_ctx = reinterpret_cast<SwapCtx*>(smem + kOffTables + 768);
  float* s_exch = reinterpret_cast<float*>(smem + kOffExch);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* full = bars;                 // [6] TMA -> MMA / tail warp (tx bytes)
  uint64_t* empty = bars + 6;            // [6] 1 (tcgen05.commit after PV) + 1 (tail warp)
  // per softmax group g = i & 1 (the k-th window of a group completes phase k of each of its barriers):
  uint64_t* s_full = bars + 12;          // [2] S(i) complete in S_g                                   (S issuer -> group)
  uint64_t* s_free = bars + 14;          // [2] the group's 8 warps have read S(i): S_g may take S(i + 2) (group -> S issuer)
  uint64_t* p_full = bars + 16;          // [2] P(i) written to P_g and O_g(i-2) drained (8 warps)     (group -> PV issuer)
  uint64_t* pv_done = bars + 18;         // [2] PV(i) complete: O_g holds O(i), P_g may be rewritten   (PV issuer -> group)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = g.C;
  const int head = blockIdx.x % g.heads, team = blockIdx.x / g.heads, nteams = gridDim.x / g.heads;
  const long long npairs = (long long)g.nZ * bd.nhw * g.nLon;
  const int p0 = (int)(npairs * team / nteams), p1 = (int)(npairs * (team + 1) / nteams);
  const int nwin = p1 - p0;
  if (nwin <= 0) return;
  const int u0 = p0 / g.nLon, last_ts = (p1 - 1) / g.nLon - u0;

  const bool is_consumer = warp < kSoftmaxWarps || warp >= kWarpTail0;
  const int ctid = warp < kSoftmaxWarps ? tid : tid - (kWarpTail0 - kSoftmaxWarps) * 32;
  auto swap_tile = [&](int ts) { swap_tile_fn(s_ctx, u0 + ts, ctid, 0); };   // all consumer threads, in lock step per tile

  // ---- prologue
  if (tid == 0) {
    s_ctx->earth_bias = earth_bias; s_ctx->s_bias = s_bias; s_ctx->s_crun = s_crun; s_ctx->s_ccls = s_ccls;
    s_ctx->g = g; s_ctx->bd = bd; s_ctx->roll = roll; s_ctx->head = head;
  }
  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < kStages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 2); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_free[i], 8); tc::mbar_init(&p_full[i], 8); tc::mbar_init(&pv_done[i], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == kWarpTma) {
    if (lane == 0) { tc::tma_prefetch_desc(&maps.own12); tc::tma_prefetch_desc(&maps.own3d12); }
    if (lane < 12) {                                          // pad-row values: (tensor s, 16-byte chunk c) of this head
      const int s = lane >> 2, c = lane & 3;
      const float* bsrc = qkv_bias + s * C + head * kHeadDim + c * 8;
      uint4 o;
      o.x = pack_bf16(bsrc[0], bsrc[1]); o.y = pack_bf16(bsrc[2], bsrc[3]);
      o.z = pack_bf16(bsrc[4], bsrc[5]); o.w = pack_bf16(bsrc[6], bsrc[7]);
      s_padvals[lane] = o;
    }
    if (lane < kStages) s_stage_u[lane] = -1;
  }
  if (warp == kWarpMma) tc::tmem_alloc(tmem_slot, kTmemCols);
  tc::tcgen05_before_sync();
  __syncthreads();
  tc::tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  tc::griddep_wait();                                         // PDL (tc_common.cuh): the set-up overlapped the predecessor's tail
  tc::griddep_launch_dependents();
  if (is_consumer) swap_tile_fn(s_ctx, u0, ctid, 1);          // first bias tile (the TMA warp is already fetching windows)

  const bool trc = kTrace && dbg && blockIdx.x == 0 && lane == 0;

  // Register budget per warpgroup (768 threads start with 80 each, and that total is the pool): the TMA / MMA group
  // shrinks to 48 so that the tail group -- whole mma.sync windows in registers -- can grow to 112.
  // (the instruction sits at the top of each role branch so that ptxas budgets the branch accordingly)
  if (warp >= kSoftmaxWarps && warp < kWarpTail0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");      // all four warps of the group, converged
  if (warp == kWarpTma) {
    // ==================================================================== TMA producer
    int cur_u = -1, real_runs = 0, halo_runs = 0;
    bool box3d[2] = {false, false};
    for (int i = 0; i < nwin; ++i) {
      const int p = p0 + i, u = p / g.nLon, l = p - u * g.nLon;
      const int st = i % kStages;
      if (u != cur_u) {
        cur_u = u;
        const int t = tile_type(g, bd, u);
        __syncwarp();
        if (lane < 12) {
          int rb, cls;
          run_info(g, bd, roll, t, lane, rb, cls);
          s_prun[lane] = rb;
          s_pcls[lane] = cls;
        }
        __syncwarp();
        real_runs = halo_runs = 0;
        for (int r = 0; r < 12; ++r) { real_runs += s_pcls[r] != 0; halo_runs += s_pcls[r] >= 2; }
        // a half window (fixed dz: 6 latitude rows x 12 longitudes) whose rows are consecutive own rows is ONE 3-D box
        // per tensor; otherwise (pad rows, halo rows, the latitude wrap) its runs are fetched one by one
#pragma unroll
        for (int dz = 0; dz < 2; ++dz) {
          bool ok = roll != 2;
          for (int dh = 0; dh < 6; ++dh) {
            ok = ok && s_pcls[dz * 6 + dh] == 1;
            if (dh > 0) ok = ok && s_prun[dz * 6 + dh] - s_prun[dz * 6 + dh - 1] == g.W;
          }
          box3d[dz] = ok;
        }
        if (u - u0 < last_ts && lane == 0) {                  // pull the next bias tile into L2 ahead of the swap
          const int tn = tile_type(g, bd, u + 1);
          prefetch_l2_bulk(earth_bias + ((long long)tn * g.heads + head) * kWinTokens * kWinTokens, kBiasTileBytes);
        }
      }
      tc::mbar_wait(&empty[st], ((i / kStages) & 1) ^ 1);
      if (real_runs < 12 && s_stage_u[st] != u) {             // pad rows of this window type into the stage
        for (int r = 0; r < 12; ++r) {
          if (s_pcls[r] != 0) continue;
          for (int j = lane; j < 144; j += 32) {
            const int k = r * 12 + j / 12, part = j % 12, s = part >> 2, c = part & 3;
            *reinterpret_cast<uint4*>(s_buf + st * kBufBytes + s * kTileBytes + tile_off(k, c)) = s_padvals[part];
          }
        }
        tc::fence_async_smem();                               // generic-proxy writes -> visible to the tensor core
      }
      __syncwarp();
      if (PANGU_ATTN_KO & 64) {
        if (lane == 0) { s_stage_u[st] = u; tc::mbar_arrive(&full[st]); }
        __syncwarp();
        continue;
      }
      if (lane == 0) {
        s_stage_u[st] = u;
        // (K/V-only halos: the q runs of halo rows are not fetched -- those score rows are never stored)
        tc::mbar_expect_tx(&full[st], (real_runs * 3 - (bd.halo_kv ? halo_runs : 0)) * kRunBytes);
      }
      __syncwarp();
      const int w = 12 * l + (roll == 1 ? 6 : 0);
      if (lane < 6) {                                         // (dz, tensor) pairs
        const int dz = lane / 3, s = lane - dz * 3;
        if (box3d[dz] && w + 12 <= g.W) {
          const int row = s_prun[dz * 6] / g.W;
          uint8_t* dst = s_buf + st * kBufBytes + s * kTileBytes + dz * 6 * kRunBytes;
          tc::tma_load_3d(dst, &maps.own3d12, &full[st], s * C + head * kHeadDim, w, row);
        }
      }
      for (int j = lane; j < 36; j += 32) {                   // (run, tensor) pairs of the half windows fetched run by run
        const int r = j / 3, s = j - r * 3;
        const int cls = s_pcls[r];
        if (cls == 0 || (box3d[r / 6] && w + 12 <= g.W)) continue;
        const int rb = s_prun[r];
        uint8_t* dst = s_buf + st * kBufBytes + s * kTileBytes + r * kRunBytes;
        int c0 = s * C + head * kHeadDim;
        if (bd.halo_kv && cls >= 2) {                         // halo buffers are [rows, 2C] = (k | v)
          if (s == 0) continue;
          c0 -= C;
        }
        if (roll == 2) {
          tc::tma_load_2d(dst, &maps.own12, &full[st], c0, l * g.T * kWinTokens + rb);
        } else if (w + 12 <= g.W) {
          const CUtensorMap* m = cls == 1 ? &maps.own12 : (cls == 2 ? &maps.hs12 : &maps.hn12);
          tc::tma_load_2d(dst, m, &full[st], c0, rb + w);
        } else {                                              // last window of a rolled block: longitudes W-6..W-1, then 0..5
          const CUtensorMap* m = cls == 1 ? &maps.own6 : (cls == 2 ? &maps.hs6 : &maps.hn6);
          tc::tma_load_2d(dst, m, &full[st], c0, rb + w);
          tc::tma_load_2d(dst + 6 * 64, m, &full[st], c0, rb);
        }
      }
      __syncwarp();
      if (trc) g_attn_trace[(i & 7) * 16 + 0] = clock64();
    }
  } else if (warp == kWarpMma) {
    // ==================================================================== S = Q K^T issuer
    // S(i) goes to the S buffer of softmax group g = i & 1 as soon as that group has finished READING S(i-2) (s_free: the end
    // of its pass 2) -- P lives in a buffer of its own, so S_g does not wait for PV(i-2).  The group covers the latency of
    // this MMA with the epilogue of window i-2, which it runs between its s_free and p_full arrivals; S(i) therefore never
    // depends on the OTHER group's progress (r2: with three shared S/P buffers S(i) waited for PV(i-3), a window of the other
    // group, and the groups spent 24 % of their time waiting for S whenever they drifted into phase).
    // The PV MMAs are issued by ANOTHER warp: issuing one tcgen05.mma costs the issuing warp cycles (descriptor build, ELECT),
    // and an S queued behind nine PV issues would stall a whole softmax group.  No tensor-pipe ordering between S and PV is
    // assumed; every hand-over has its mbarrier.
    constexpr uint32_t idesc_s = tc::make_idesc_bf16(128, 144, 0, 0);
    for (int i = 0; i < nwin; ++i) {
      const int st = i % kStages, g2 = i & 1;
      tc::mbar_wait(&full[st], (i / kStages) & 1);
      if (i >= 2) tc::mbar_wait(&s_free[g2], ((i >> 1) & 1) ^ 1);
      tc::tcgen05_after_sync();
      const uint32_t aq = smem_u32(s_buf + st * kBufBytes), ak = aq + kTileBytes;
      if (tc::elect_one()) {                                 // one elected lane, real branch: uniform-register descriptors
#pragma unroll
        for (int k = 0; k < 2; ++k)                          // K = 32: two 16-element steps, +32 B inside the 64 B swizzle span
          tc::umma_bf16(tmem_base + g2 * kColS, desc_k_sw64(aq) + 2 * k, desc_k_sw64(ak) + 2 * k, idesc_s, k);
        tc::umma_commit(&s_full[g2]);
      }
      __syncwarp();
      if (trc) g_attn_trace[(i & 7) * 16 + 1] = clock64();
    }
  } else if (warp == kWarpPv) {
    // ==================================================================== O = P V issuer
    // O(i) goes to the O buffer of group i & 1: its previous content, O(i-2), was read by the group's epilogue before it
    // arrived on p_full(i).
    constexpr uint32_t idesc_o = tc::make_idesc_bf16(128, 32, 0, 1);
    for (int i = 0; i < nwin; ++i) {
      const int st = i % kStages, g2 = i & 1;
      tc::mbar_wait(&p_full[g2], (i >> 1) & 1);
      tc::tcgen05_after_sync();
      if (trc) g_attn_trace[(i & 7) * 16 + 2] = clock64();
      const uint32_t av = smem_u32(s_buf + st * kBufBytes + 2 * kTileBytes);
      const uint32_t tP = tmem_base + kColP + 72 * g2, tO = tmem_base + kColO + 32 * g2;
      if (tc::elect_one()) {                                 // one elected lane, real branch: uniform-register descriptors
#pragma unroll
        for (int k = 0; k < 9; ++k)                          // K = 144 keys: 16 keys = 8 packed TMEM columns / 1 KiB of V per step
          tc::umma_bf16_ts(tO, tP + 8 * k, desc_mn_sw64(av + 1024 * k), idesc_o, k);
        tc::umma_commit(&pv_done[g2]); tc::umma_commit(&empty[st]);
      }
      __syncwarp();
      if (trc) g_attn_trace[(i & 7) * 16 + 3] = clock64();
    }
  }
  } else if (warp < kSoftmaxWarps) {
    // ==================================================================== softmax: thread = (score row, 80 or 64 keys)
    const int grp = warp >> 3, q = warp & 3, hf = (warp >> 2) & 1;
    const int row = q * 32 + lane;
    const int k0 = hf * kKeys0, nchunk = hf ? 4 : 5;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t brow = smem_u32(s_bias + row * kBiasPitch + k0);
    const uint32_t exg = smem_u32(s_exch + grp * 1024);       // [slot][max | sum][half][128] floats
    const int rrun = row / 12, rdw = row - rrun * 12;
    const bool trw = trc && (warp & 7) == 0;

    int tok_prev = -1;                                        // output token (x 2 + buffer class) of window i-2, or -1
    float m_prev = 0.f;                                       // merged row maximum of window i-2 (for its lse row)

    // O(i) -> global (each thread 16 channels of its row = one 32-byte sector)
    // (the caller has waited for pv_done of window i)
    auto epilogue = [&](int i, int tokc, float m_row) {
      const int slot = (i >> 1) & 1;
      uint32_t o[16];
      tc::tmem_ld_32x16(tmem_base + lane_addr + kColO + 32 * grp + hf * 16, o);
      const uint32_t exs = exg + (slot * 512 + 256 + row) * 4;
      const float sum = lds_f32(exs) + lds_f32(exs + 512);
      const float inv = kTruncFix / sum;                        // P was truncated to bf16 (pack_bf16_trunc)
      tmem_ld_wait16(o);
      if (tokc >= 0 && !(PANGU_ATTN_KO & 8)) {
        uint4 a, c;
        a.x = pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
        a.y = pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
        a.z = pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
        a.w = pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
        c.x = pack_bf16(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
        c.y = pack_bf16(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
        c.z = pack_bf16(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
        c.w = pack_bf16(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
        __nv_bfloat16* base = (tokc & 1) ? halo_out : out;
        uint4* d = reinterpret_cast<uint4*>(base + (long long)(tokc >> 1) * C + head * kHeadDim + hf * 16);
        d[0] = a; d[1] = c;
      }
      if (lse != nullptr && hf == 0) {                        // training only: log2-sum-exp of the row, for the backward
        const int p = p0 + i, pu = p / g.nLon, pl = p - pu * g.nLon;
        lse[((pl * g.T + tile_type(g, bd, pu)) * g.heads + head) * kWinTokens + row] = m_row + log2f(sum);   // m_row: the merged row maximum
      }
    };

    // tile by tile: the (out-of-line) bias swap stays outside the window loop, whose registers it would otherwise pin
    int i = grp;
#pragma unroll 1
    for (int ts = 0; ts <= last_ts; ++ts) {
    if (ts > 0) {
      if (trc && warp == 0) g_attn_trace[12] = clock64();
      swap_tile(ts);
      if (trc && warp == 0) g_attn_trace[13] = clock64();
    }
    const int i_end = min(nwin, (u0 + ts + 1) * g.nLon - p0);    // windows [.., i_end) lie in tile u0 + ts
    const bool masked_tile = exact_all != 0;                  // the caller asked for the exact row maximum (wide bias tables)
    float bm[3] = {0.f, 0.f, 0.f};                            // per key sub-block: maximum of my part of this tile's bias row
    if (i < i_end && !masked_tile) {
      if (hf == 0) bias_sub_max<0>(brow, bm); else bias_sub_max<1>(brow, bm);
    }
#pragma unroll 1
    for (; i < i_end; i += 2) {
      const int slot = (i >> 1) & 1;
      const int l = p0 + i - (u0 + ts) * g.nLon;
      long long* tr = g_attn_trace + (i & 7) * 16;
      const int rcls = s_ccls[rrun], rrb = s_crun[rrun];      // tables of the current tile (rebuilt by every swap)
      const int tok_cur = (rcls == 1 || (rcls == 2 && halo_out != nullptr)) ? (int)run_token(g, roll, l, rrb, rdw) * 2 + (rcls == 2) : -1;
      const uint32_t tS = tmem_base + lane_addr + grp * kColS + k0;              // my keys of S(i)
      const uint32_t tP = tmem_base + lane_addr + kColP + 72 * grp + hf * 40;    // my keys of P(i), two per column

      tc::mbar_wait(&s_full[grp], (i >> 1) & 1);
      tc::tcgen05_after_sync();
      if (trw) tr[4] = clock64();
      const uint32_t exm = exg + slot * 2048, exs = exm + 1024;
      float m;
      {
        // ---- pass 1: an UPPER BOUND of the row maximum of S + bias over my keys (score_bound above: max(S) + max(bias)
        // per shift-mask sub-block).  Softmax is invariant under the shift, and bf16 P / fp32 sums keep their relative
        // precision over the whole exponent range, so any bound within ~100 (log2 units) of the true maximum gives the same
        // result; this one is off by at most the spread of the un-masked bias values of the row.  It saves the bias reads
        // and adds of a full pass.  (exact_all: bias rows that spread wider than that take the exact maximum instead.)
        float mx0;
        if (masked_tile) {                                    // exact maximum of S + bias (wide bias tables only)
          float mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
          mx0 = -INFINITY;
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            if (c < nchunk) {
              uint32_t v[16];
              tc::tmem_ld_32x16(tS + 16 * c, v);
              const uint4 b0 = lds128(brow + 32 * c), b1 = lds128(brow + 32 * c + 16);
              const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              tmem_ld_wait16(v);
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                float s0 = __uint_as_float(v[2 * e]), s1 = __uint_as_float(v[2 * e + 1]);
                float s2 = __uint_as_float(v[2 * e + 2]), s3 = __uint_as_float(v[2 * e + 3]);
                add_bias2(bw[e], s0, s1);
                add_bias2(bw[e + 1], s2, s3);
                mx0 = fmaxf(mx0, s0); mx1 = fmaxf(mx1, s1); mx2 = fmaxf(mx2, s2); mx3 = fmaxf(mx3, s3);
              }
            }
          }
          mx0 = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        } else if (PANGU_ATTN_KO & 1) {
          mx0 = 30.0f;
        } else {
          mx0 = hf == 0 ? score_bound<0>(tS, bm) : score_bound<1>(tS, bm);
        }
        m = mx0;
        sts_f32(exm + (hf * 128 + row) * 4, m);
      }
      if (trw) tr[5] = clock64();
      {
        if (!(PANGU_ATTN_KO & 32)) named_bar(1 + grp * 4 + q, 64);                       // the two warps of this lane quarter
        m = fmaxf(m, lds_f32(exm + ((hf ^ 1) * 128 + row) * 4));
        sts_f32(exm + (hf * 128 + row) * 4, m);               // the merged maximum, for the lse of the epilogue
        if (trw) tr[6] = clock64();
        // PV(i-2), issued when this group finished window i-2, has read P_g and left O(i-2) in O_g (complete long ago)
        if (i >= 2) { tc::mbar_wait(&pv_done[grp], ((i >> 1) & 1) ^ 1); tc::tcgen05_after_sync(); }
        // ---- pass 2: P = exp2(S + bias - m) -> bf16 pairs into the group's P buffer; row sum
        float sm0 = 0.f, sm1 = 0.f;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          if (c < nchunk) {
            uint32_t v[16];
            if (PANGU_ATTN_KO & 128) {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(m) + e;
            } else {
              tc::tmem_ld_32x16(tS + 16 * c, v);
            }
            uint4 b0 = make_uint4(0, 0, 0, 0), b1 = b0;
            if (!(PANGU_ATTN_KO & 4)) { b0 = lds128(brow + 32 * c); b1 = lds128(brow + 32 * c + 16); }
            const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            tmem_ld_wait16(v);
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float s0 = __uint_as_float(v[2 * e]), s1 = __uint_as_float(v[2 * e + 1]);
              if (!(PANGU_ATTN_KO & 4)) add_bias2(bw[e], s0, s1);
              add2(s0, s1, -m, -m);
              const float e0 = (PANGU_ATTN_KO & 2) ? s0 : ex2(s0), e1 = (PANGU_ATTN_KO & 2) ? s1 : ex2(s1);
              add2(sm0, sm1, e0, e1);
              pk[e] = pack_bf16_trunc(e0, e1);
            }
            if (PANGU_ATTN_KO & 128) { sm0 += __uint_as_float(pk[0] ^ pk[3] ^ pk[5] ^ pk[7]); sm1 += __uint_as_float(pk[1] ^ pk[2] ^ pk[4] ^ pk[6]); }
            else tmem_st_32x8(tP + 8 * c, pk);
          }
        }
        sts_f32(exs + (hf * 128 + row) * 4, sm0 + sm1);
      }
      // S(i) has been read (every load's registers were consumed): S_g may take S(i+2) NOW; its MMAs run under the epilogue
      tc::tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_free[grp]);
      // the epilogue of window i-2 drains O_g before this group's arrival on p_full lets PV(i) refill it
      if (i >= 2) epilogue(i - 2, tok_prev, m_prev);
      tc::tmem_st_wait();
      tc::tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&p_full[grp]);
      if (trw) tr[7] = clock64();
      tok_prev = tok_cur;
      m_prev = m;
    }
    }
    if (i >= 2) {                                             // the last window of this group
      tc::mbar_wait(&pv_done[grp], ((i - 2) >> 1) & 1);
      tc::tcgen05_after_sync();
      epilogue(i - 2, tok_prev, m_prev);
    }
  } else {
    // ==================================================================== tail warps: rows 128..143 of window i = j (mod 4)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    constexpr int row0 = 128;
    const int jt = warp - kWarpTail0;
    const int gq = lane >> 2, tq = lane & 3, mi = lane >> 3, mr = lane & 7;
    int i = jt;
#pragma unroll 1
    for (int ts = 0; ts <= last_ts; ++ts) {
    if (ts > 0) swap_tile(ts);
    const int i_end = min(nwin, (u0 + ts + 1) * g.nLon - p0);
    const int t = tile_type(g, bd, u0 + ts);
#pragma unroll 1
    for (; i < i_end; i += kTailWarps) {
      const int l = p0 + i - (u0 + ts) * g.nLon;
      const int st = i % kStages;
      uint8_t* sq = s_buf + st * kBufBytes;
      uint8_t* sk = sq + kTileBytes;
      uint8_t* sv = sk + kTileBytes;
      tc::mbar_wait(&full[st], (i / kStages) & 1);
      if (PANGU_ATTN_KO & 16) { __syncwarp(); if (lane == 0) tc::mbar_arrive(&empty[st]); continue; }
      if (trc) g_attn_trace[(i & 7) * 16 + 10] = clock64();
      uint32_t qa[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int r = row0 + (mi & 1) * 8 + mr, c = ks * 2 + (mi >> 1);
        ldmatrix_x4(attn::smem_u32(sq + tile_off(r, c)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
      }
      // the whole 16 x 144 score block at once (no online rescaling chain: one warp per window must find its
      // instruction-level parallelism inside the window)
      const uint32_t b_lo = smem_u32(s_bias + (row0 + gq) * kBiasPitch + 2 * tq), b_hi = b_lo + 8 * kBiasPitch * 2;
      float s_acc[18][4];
#pragma unroll
      for (int nt = 0; nt < 18; ++nt) {                       // accumulate on top of the (pre-scaled, pre-masked) bias
        const uint32_t w_lo = lds32(b_lo + nt * 16), w_hi = lds32(b_hi + nt * 16);
        s_acc[nt][0] = __uint_as_float(w_lo << 16); s_acc[nt][1] = __uint_as_float(w_lo & 0xffff0000u);
        s_acc[nt][2] = __uint_as_float(w_hi << 16); s_acc[nt][3] = __uint_as_float(w_hi & 0xffff0000u);
        uint32_t k0r, k1r, k2r, k3r;
        ldmatrix_x4(attn::smem_u32(sk + tile_off(nt * 8 + mr, mi)), k0r, k1r, k2r, k3r);
        mma_bf16(s_acc[nt], qa[0], k0r, k1r);
        mma_bf16(s_acc[nt], qa[1], k2r, k3r);
      }
      float m_lo = -INFINITY, m_hi = -INFINITY, m_lo2 = -INFINITY, m_hi2 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 18; nt += 2) {
        m_lo = fmaxf(m_lo, fmaxf(s_acc[nt][0], s_acc[nt][1]));
        m_hi = fmaxf(m_hi, fmaxf(s_acc[nt][2], s_acc[nt][3]));
        m_lo2 = fmaxf(m_lo2, fmaxf(s_acc[nt + 1][0], s_acc[nt + 1][1]));
        m_hi2 = fmaxf(m_hi2, fmaxf(s_acc[nt + 1][2], s_acc[nt + 1][3]));
      }
      m_lo = fmaxf(m_lo, m_lo2); m_hi = fmaxf(m_hi, m_hi2);
      m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
      m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
      m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
      m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
      uint32_t pa[9][4];                                      // P as A fragments, 9 k-steps of 16 keys
#pragma unroll
      for (int nt = 0; nt < 18; ++nt) {
        const float e0 = ex2(s_acc[nt][0] - m_lo), e1 = ex2(s_acc[nt][1] - m_lo);
        const float e2 = ex2(s_acc[nt][2] - m_hi), e3 = ex2(s_acc[nt][3] - m_hi);
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(e0, e1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(e2, e3);
      }
      float o_acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) o_acc[a][c] = 0.f;
      float l_acc[4] = {0.f, 0.f, 0.f, 0.f};                  // row sums of the bf16 P, from a ones-column MMA
#pragma unroll
      for (int kk = 0; kk < 9; ++kk) {
        mma_bf16(l_acc, pa[kk], 0x3F803F80u, 0x3F803F80u);
#pragma unroll
        for (int dp = 0; dp < 2; ++dp) {
          uint32_t v0, v1, v2, v3;
          const int r = kk * 16 + (mi & 1) * 8 + mr, c = dp * 2 + (mi >> 1);
          ldmatrix_x4_trans(attn::smem_u32(sv + tile_off(r, c)), v0, v1, v2, v3);
          mma_bf16(o_acc[dp * 2], pa[kk], v0, v1);
          mma_bf16(o_acc[dp * 2 + 1], pa[kk], v2, v3);
        }
      }
      const float inv_lo = 1.0f / l_acc[0], inv_hi = 1.0f / l_acc[2];
      if (lse != nullptr && tq == 0) {
        float* L = lse + (((long long)l * g.T + t) * g.heads + head) * kWinTokens + row0 + gq;
        L[0] = m_lo + log2f(l_acc[0]);
        L[8] = m_hi + log2f(l_acc[2]);
      }
      // rows 128..143 of the Q tile are outside the tensor core's M = 128 and only read by this warp (done): stage O
      // there, then 64-byte coalesced row stores
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        *reinterpret_cast<uint32_t*>(sq + tile_off(row0 + gq, nt) + 4 * tq) = pack_bf16(o_acc[nt][0] * inv_lo, o_acc[nt][1] * inv_lo);
        *reinterpret_cast<uint32_t*>(sq + tile_off(row0 + gq + 8, nt) + 4 * tq) = pack_bf16(o_acc[nt][2] * inv_hi, o_acc[nt][3] * inv_hi);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int idx = lane + it * 32, r = row0 + (idx >> 2), c = idx & 3;
        const int rr = r / 12, dw = r - rr * 12;
        const int cls = s_ccls[rr], rb = s_crun[rr];
        __nv_bfloat16* dstp = cls == 1 ? out : (cls == 2 ? halo_out : nullptr);
        if (dstp != nullptr) {
          const uint4 val = *reinterpret_cast<const uint4*>(sq + tile_off(r, c));
          *reinterpret_cast<uint4*>(dstp + run_token(g, roll, l, rb, dw) * C + head * kHeadDim + c * 8) = val;
        }
      }
      __syncwarp();
      // the staged rows were written through the generic proxy; the next TMA into this slot writes through the async proxy
      tc::fence_async_smem();
      if (lane == 0) tc::mbar_arrive(&empty[st]);
      if (trc) g_attn_trace[(i & 7) * 16 + 11] = clock64();
    }
    }
  }
  tc::tcgen05_before_sync();
  __syncthreads();
  tc::tcgen05_after_sync();
  if (warp == kWarpMma) tc::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace attn2

int debug_read_attn_trace(long long* out, int n) {
  if (n > 8 * 16) n = 8 * 16;
  cudaError_t e = cudaMemcpyFromSymbol(out, attn2::g_attn_trace, sizeof(long long) * n);
  if (e != cudaSuccess) { set_error("debug_read_attn_trace: %s", cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
  return PANGU_OK;
}

namespace tc { int num_sms(); }

int launch_window_attention_tc(const void* qkv, const void* halo_qkv, const void* halo_lo_qkv, const float* qkv_bias,
                               const void* earth_bias, void* out, void* halo_out, const WinGeom& g, const BandGeom& bd,
                               int roll, cudaStream_t st, float* lse, int exact_max) {
  using namespace attn2;
  if (bd.nhw <= 0) return PANGU_OK;
  static unsigned long long configured = 0;
  {
    cudaError_t e = pangu::set_max_smem_once(configured, window_attention_tc_kernel, kSmemBytes);
    if (e != cudaSuccess) { set_error("attention_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
  }
  // token rows of the three tensors
  const uint64_t own_rows = roll == 2 ? (uint64_t)g.nLon * g.T * kWinTokens : (uint64_t)g.Z * bd.hrows * g.W;
  const uint64_t pitch = (uint64_t)3 * g.C * 2;
  Maps maps;
  auto enc = [&](CUtensorMap* m12, CUtensorMap* m6, const void* p, uint64_t rows, int ntens = 3) -> bool {
    const uint64_t cols = (uint64_t)ntens * g.C;
    return tc::encode_tmap_2d(m12, 1, p, cols, rows, cols * 2, 32, 12, 64) &&
           tc::encode_tmap_2d(m6, 1, p, cols, rows, cols * 2, 32, 6, 64);
  };
  if (!enc(&maps.own12, &maps.own6, qkv, own_rows)) return PANGU_ERR_CUDA;
  const int halo_tens = bd.halo_kv ? 2 : 3;
  if (bd.halo_kv && halo_out != nullptr) { set_error("attention_tc: K/V-only halos cannot produce halo outputs"); return PANGU_ERR_BAD_ARG; }
  maps.hs12 = maps.own12; maps.hs6 = maps.own6; maps.hn12 = maps.own12; maps.hn6 = maps.own6;
  {   // [Z * rows][W][3C] view of the own rows (not meaningful for pre-partitioned windows, where it is never used)
    const uint64_t rows3 = roll == 2 ? 1 : (uint64_t)g.Z * bd.hrows, W3 = roll == 2 ? 12 : (uint64_t)g.W;
    if (!tc::encode_tmap_3d_bf16(&maps.own3d12, qkv, (uint64_t)3 * g.C, W3, rows3, pitch, pitch * W3, 32, 12, 6, 64)) return PANGU_ERR_CUDA;
  }
  if (bd.halo > 0 && !enc(&maps.hs12, &maps.hs6, halo_qkv, (uint64_t)g.Z * bd.halo * g.W, halo_tens)) return PANGU_ERR_CUDA;
  if (bd.halo_lo > 0 && !enc(&maps.hn12, &maps.hn6, halo_lo_qkv, (uint64_t)g.Z * bd.halo_lo * g.W, halo_tens)) return PANGU_ERR_CUDA;
  // teams of `heads` CTAs share a range of (window type, longitude window) pairs
  if (g.heads > tc::num_sms()) { set_error("attention_tc: more heads (%d) than SMs", g.heads); return PANGU_ERR_BAD_ARG; }
  const long long npairs = (long long)g.nZ * bd.nhw * g.nLon;
  long long nteams = tc::num_sms() / g.heads;
  static const int forced = []() { const char* e = getenv("PANGU_ATTN_TEAMS"); return e ? atoi(e) : 0; }();
  if (forced > 0 && forced < nteams) nteams = forced;
  if (nteams > npairs) nteams = npairs;
  static const int dbg = []() { const char* e = getenv("PANGU_ATTN_DBG"); return e ? atoi(e) : 0; }();
  cudaError_t le = tc::launch_pdl(window_attention_tc_kernel, dim3((unsigned)(nteams * g.heads)), dim3(kThreads), kSmemBytes, st,
                                  maps, qkv_bias, (const __nv_bfloat16*)earth_bias, (__nv_bfloat16*)out, (__nv_bfloat16*)halo_out, g, bd,
                                  roll, lse, dbg, exact_max);
  if (le != cudaSuccess) { set_error("window_attention_tc: launch: %s", cudaGetErrorString(le)); return PANGU_ERR_CUDA; }
  return check_launch("window_attention_tc");
}

}  // namespace pangu
