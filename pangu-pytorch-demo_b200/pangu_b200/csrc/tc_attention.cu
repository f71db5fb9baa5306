// bf16 3-D window attention with the Earth-specific bias (models/layers.py:422-478) and the block's
// pad / roll / partition / shift-mask / reverse / crop (models/layers.py:224-293) folded into the
// addressing.  One CTA owns ONE (window type t, head) bias tile -- staged once in shared memory --
// and walks a chunk of longitude windows l, so the 144x144 bias tile is reused instead of re-read.
//
// Per (window, head): S = q k^T on tensor cores (bf16 mma, fp32 accumulate), online softmax in
// registers over three 48-key blocks (exp2 with scale*log2e folded into one FMA), O = P v on tensor
// cores, O staged through shared memory and written as 64-byte rows at the UN-rolled token position.
//
// Round-1 note: the two small GEMMs (7 % of the model's FLOPs) use warp-level mma.sync fragments so that
// softmax stays register-resident; the tcgen05/TMEM formulation is the planned upgrade.
#include <cstdlib>

#include "attn_common.cuh"

namespace pangu {
namespace tc { int num_sms(); }
static inline int tc_num_sms() { return tc::num_sms(); }
namespace attn {

constexpr int kWarps = 9;                         // 9 x 16 query rows = 144
constexpr int kThreads = kWarps * 32;
constexpr int kKvBlock = 48;                      // keys per online-softmax block (3 blocks)
constexpr int kBufBytes = 3 * kTileBytes;         // q, k, v
constexpr int kSmemBytes = kWinTokens * kBiasPitch * 2 + 2 * kBufBytes + kWinTokens * 4 /*rowbase*/ +
                           kWinTokens * 4 /*dw*/ + kWinTokens /*gid*/ + 16;

// PRE: the caller folded scale*log2(e) into the q projection (weights and bias) and log2(e) into the bias table,
// so q k^T + bias is already the softmax exponent in log2 units: the bias tile initialises the MMA accumulator
// and no per-score FMA is needed.
template <typename TB, bool PRE>
__global__ void __launch_bounds__(kThreads, 2)
window_attention_bf16_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ halo_qkv,
                             const __nv_bfloat16* __restrict__ halo_lo_qkv,
                             const float* __restrict__ qkv_bias, const TB* __restrict__ earth_bias,
                             __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ halo_out, WinGeom g,
                             BandGeom bd, int roll, int lon_chunk, float* __restrict__ lse) {
  extern __shared__ __align__(128) uint8_t smem[];
  __nv_bfloat16* s_bias = reinterpret_cast<__nv_bfloat16*>(smem);                       // [144][152]
  uint8_t* s_buf = smem + kWinTokens * kBiasPitch * 2;                                  // 2 x {q,k,v}
  int* s_rowbase = reinterpret_cast<int*>(s_buf + 2 * kBufBytes);                       // [144] (z*H+h)*W or -1
  int* s_dw = s_rowbase + kWinTokens;                                                   // [144] dw | row class << 8
  uint8_t* s_gid = reinterpret_cast<uint8_t*>(s_dw + kWinTokens);                       // [144]

  // Latitude band (pangu_b200/dist.py): this launch covers bd.nhw h-windows starting at global window bd.hw0
  // (the last one being the global wrap window nH-1 when bd.wrap); qkv/out hold rows [bd.h0, bd.h0+bd.hrows) of
  // the global grid, halo_qkv/halo_out the bd.halo rows that follow, halo_lo_qkv the bd.halo_lo rows that
  // precede them.  Row classes: 0 pad, 1 own, 2 southern halo, 3 northern halo.  Full grid: {0, H, 0, nH, 0, 0, 0}.
  const int head = blockIdx.x, lchunk = blockIdx.y;
  const int zw_ = blockIdx.z / bd.nhw, hwl_ = blockIdx.z - zw_ * bd.nhw;
  const int t = zw_ * g.nH + ((bd.wrap && hwl_ == bd.nhw - 1) ? g.nH - 1 : bd.hw0 + hwl_);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = g.C;
  const int l_begin = lchunk * lon_chunk;
  const int l_end = min(g.nLon, l_begin + lon_chunk);

  // ---- per-CTA tables: source row of every window element (independent of l) and mask group ids
  for (int k = tid; k < kWinTokens; k += kThreads) {
    const int zw = t / g.nH, hw = t - zw * g.nH;
    const int dz = k / 72, r = k - dz * 72, dh = r / 12, dw = r - dh * 12;
    int z = 2 * zw + dz, h = 6 * hw + dh;
    if (roll == 1) { z += 1; if (z >= g.Z) z -= g.Z; h += 3; if (h >= g.Hp) h -= g.Hp; }
    if (roll == 2) {                                        // pre-partitioned windows: identity map
      s_rowbase[k] = t * kWinTokens + k;
      s_dw[k] = 1 << 8;
    } else {
      const int hl = h - bd.h0;
      int rb = -1, cls = 0;                                   // zero pad row (or outside this band: never selected)
      if (h < g.H) {
        if (hl >= 0 && hl < bd.hrows) { rb = (z * bd.hrows + hl) * g.W; cls = 1; }
        else if (hl >= bd.hrows && hl < bd.hrows + bd.halo) { rb = (z * bd.halo + (hl - bd.hrows)) * g.W; cls = 2; }
        else if (hl < 0 && hl >= -bd.halo_lo) { rb = (z * bd.halo_lo + (hl + bd.halo_lo)) * g.W; cls = 3; }
      }
      s_rowbase[k] = rb;
      s_dw[k] = dw | (cls << 8);
    }
    s_gid[k] = (uint8_t)shift_group(g, t, k);
  }
  // ---- bias tile of this (t, head) -> smem (bf16)
  {
    const TB* src = earth_bias + ((long long)t * g.heads + head) * kWinTokens * kWinTokens;
    if (sizeof(TB) == 2) {
      for (int i = tid; i < kWinTokens * 18; i += kThreads) {          // 18 chunks of 8 bf16 per row
        const int r = i / 18, c = i - r * 18;
        cp_async16(smem_u32(s_bias + r * kBiasPitch + c * 8), reinterpret_cast<const __nv_bfloat16*>(src) + r * kWinTokens + c * 8);
      }
    } else {
      for (int i = tid; i < kWinTokens * 36; i += kThreads) {          // fp32 table: convert on the fly
        const int r = i / 36, c = i - r * 36;
        const float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + r * kWinTokens + c * 4));
        uint2 o = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        *reinterpret_cast<uint2*>(s_bias + r * kBiasPitch + c * 4) = o;
      }
    }
  }
  __syncthreads();

  // token of window element k in longitude window l: index into the buffer of the element's row class
  auto token_of = [&](int l, int dwc, int rb) -> long long {      // dwc = s_dw[k]
    if (roll == 2) return (long long)l * g.T * kWinTokens + rb;
    int w = 12 * l + (roll == 1 ? 6 : 0) + (dwc & 0xff);
    if (w >= g.W) w -= g.W;
    return (long long)rb + w;
  };
  // issue the gather of window l into buffer b: 144 tokens x {q,k,v} x 4 chunks of 16 B
  auto issue_load = [&](int l, int b) {
    uint8_t* buf = s_buf + b * kBufBytes;
    for (int i = tid; i < kWinTokens * 12; i += kThreads) {
      const int k = i / 12, part = i - k * 12, s = part >> 2, c = part & 3;
      uint8_t* dst = buf + s * kTileBytes + tile_off(k, c);
      const int rb = s_rowbase[k];
      const int dwc = s_dw[k], cls = dwc >> 8;
      if (cls != 0) {
        const __nv_bfloat16* base = cls == 1 ? qkv : (cls == 2 ? halo_qkv : halo_lo_qkv);
        cp_async16(smem_u32(dst), base + token_of(l, dwc, rb) * 3 * C + s * C + head * kHeadDim + c * 8);
      } else {                                            // zero pad row: linear1(0) = bias (layers.py:228,419)
        const float* bsrc = qkv_bias + s * C + head * kHeadDim + c * 8;
        uint4 o;
        o.x = pack_bf16(bsrc[0], bsrc[1]); o.y = pack_bf16(bsrc[2], bsrc[3]);
        o.z = pack_bf16(bsrc[4], bsrc[5]); o.w = pack_bf16(bsrc[6], bsrc[7]);
        *reinterpret_cast<uint4*>(dst) = o;
      }
    }
  };

  issue_load(l_begin, 0);
  cp_async_commit();

  const float sl2 = rsqrtf((float)kHeadDim) * kLog2e;      // scale * log2(e)
  const float mask_l2 = kMaskValue * kLog2e;
  const int gq = lane >> 2, tq = lane & 3;                 // fragment coordinates
  const int row0 = warp * 16;
  const int mi = lane >> 3, mr = lane & 7;                 // ldmatrix: matrix index / row within matrix
  const bool masked_type = roll == 1 && ((t / g.nH == g.nZ - 1) || (t % g.nH == g.nH - 1));

  for (int l = l_begin; l < l_end; ++l) {
    const int b = (l - l_begin) & 1;
    if (l + 1 < l_end) issue_load(l + 1, b ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    uint8_t* sq = s_buf + b * kBufBytes;
    uint8_t* sk = sq + kTileBytes;
    uint8_t* sv = sk + kTileBytes;

    // Q fragments of this warp's 16 rows, two k-steps (d 0..15, 16..31)
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int r = row0 + (mi & 1) * 8 + mr, c = ks * 2 + (mi >> 1);
      ldmatrix_x4(smem_u32(sq + tile_off(r, c)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    }
    const int gid_lo = s_gid[row0 + gq], gid_hi = s_gid[row0 + gq + 8];

    float o_acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
    float m_lo = -INFINITY, m_hi = -INFINITY;
    float l_acc[4] = {0.f, 0.f, 0.f, 0.f};                  // row sums of the bf16 P, from a ones-column MMA

#pragma unroll 1
    for (int kv0 = 0; kv0 < kWinTokens; kv0 += kKvBlock) {
      float s_acc[6][4];
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) {
        if (PRE) {                                          // accumulate on top of the (pre-scaled) bias
          const int j = kv0 + nt * 8 + 2 * tq;
          const __nv_bfloat162 b_lo = *reinterpret_cast<const __nv_bfloat162*>(s_bias + (row0 + gq) * kBiasPitch + j);
          const __nv_bfloat162 b_hi = *reinterpret_cast<const __nv_bfloat162*>(s_bias + (row0 + gq + 8) * kBiasPitch + j);
          s_acc[nt][0] = __low2float(b_lo); s_acc[nt][1] = __high2float(b_lo);
          s_acc[nt][2] = __low2float(b_hi); s_acc[nt][3] = __high2float(b_hi);
        } else {
          s_acc[nt][0] = s_acc[nt][1] = s_acc[nt][2] = s_acc[nt][3] = 0.f;
        }
        uint32_t k0, k1, k2, k3;                            // b0,b1 of k-step 0 ; b0,b1 of k-step 1
        ldmatrix_x4(smem_u32(sk + tile_off(kv0 + nt * 8 + mr, mi)), k0, k1, k2, k3);
        mma_bf16(s_acc[nt], qa[0], k0, k1);
        mma_bf16(s_acc[nt], qa[1], k2, k3);
      }
      // scores in log2 units: s*scale*log2e + bias*log2e (+ mask)
      float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) {
        const int j = kv0 + nt * 8 + 2 * tq;
        if (!PRE) {
          const __nv_bfloat162 b_lo = *reinterpret_cast<const __nv_bfloat162*>(s_bias + (row0 + gq) * kBiasPitch + j);
          const __nv_bfloat162 b_hi = *reinterpret_cast<const __nv_bfloat162*>(s_bias + (row0 + gq + 8) * kBiasPitch + j);
          s_acc[nt][0] = fmaf(s_acc[nt][0], sl2, __low2float(b_lo) * kLog2e);
          s_acc[nt][1] = fmaf(s_acc[nt][1], sl2, __high2float(b_lo) * kLog2e);
          s_acc[nt][2] = fmaf(s_acc[nt][2], sl2, __low2float(b_hi) * kLog2e);
          s_acc[nt][3] = fmaf(s_acc[nt][3], sl2, __high2float(b_hi) * kLog2e);
        }
        if (masked_type) {
          const int g0 = s_gid[j], g1 = s_gid[j + 1];
          if (g0 != gid_lo) s_acc[nt][0] += mask_l2;
          if (g1 != gid_lo) s_acc[nt][1] += mask_l2;
          if (g0 != gid_hi) s_acc[nt][2] += mask_l2;
          if (g1 != gid_hi) s_acc[nt][3] += mask_l2;
        }
        mx_lo = fmaxf(mx_lo, fmaxf(s_acc[nt][0], s_acc[nt][1]));
        mx_hi = fmaxf(mx_hi, fmaxf(s_acc[nt][2], s_acc[nt][3]));
      }
      mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
      mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
      mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
      mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
      const float a_lo = ex2(m_lo - mx_lo), a_hi = ex2(m_hi - mx_hi);     // 0 on the first block (m = -inf)
      m_lo = mx_lo; m_hi = mx_hi;
      l_acc[0] *= a_lo; l_acc[2] *= a_hi;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        o_acc[nt][0] *= a_lo; o_acc[nt][1] *= a_lo; o_acc[nt][2] *= a_hi; o_acc[nt][3] *= a_hi;
      }
      uint32_t pa[3][4];                                    // P as A fragments, 3 k-steps of 16 keys
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) {
        const float p0 = ex2(s_acc[nt][0] - m_lo), p1 = ex2(s_acc[nt][1] - m_lo);
        const float p2 = ex2(s_acc[nt][2] - m_hi), p3 = ex2(s_acc[nt][3] - m_hi);
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
      // O += P V : 3 k-steps x 4 d-tiles; V fragments via transposed ldmatrix.  A fifth "d-tile" of ones
      // (bf16 1.0 pairs, a register constant) accumulates the row sums of P on the tensor pipe.
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        mma_bf16(l_acc, pa[kk], 0x3F803F80u, 0x3F803F80u);
#pragma unroll
        for (int dp = 0; dp < 2; ++dp) {
          uint32_t v0, v1, v2, v3;                          // (b0,b1) of d-tile 2dp ; (b0,b1) of d-tile 2dp+1
          const int r = kv0 + kk * 16 + (mi & 1) * 8 + mr, c = dp * 2 + (mi >> 1);
          ldmatrix_x4_trans(smem_u32(sv + tile_off(r, c)), v0, v1, v2, v3);
          mma_bf16(o_acc[dp * 2], pa[kk], v0, v1);
          mma_bf16(o_acc[dp * 2 + 1], pa[kk], v2, v3);
        }
      }
    }
    const float inv_lo = 1.0f / l_acc[0], inv_hi = 1.0f / l_acc[2];   // every column of the ones tile holds the row sum
    if (lse != nullptr && tq == 0) {                        // log2-sum-exp of every score row, kept for the backward kernel
      float* L = lse + (((long long)l * g.T + t) * g.heads + head) * kWinTokens + row0 + gq;
      L[0] = m_lo + log2f(l_acc[0]);
      L[8] = m_hi + log2f(l_acc[2]);
    }

    // stage O (bf16) into this warp's own, now dead, Q rows; then 64-byte coalesced row stores
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      // element (row, d = nt*8 + 2tq): chunk nt, byte 4*tq inside the chunk
      *reinterpret_cast<uint32_t*>(sq + tile_off(row0 + gq, nt) + 4 * tq) = pack_bf16(o_acc[nt][0] * inv_lo, o_acc[nt][1] * inv_lo);
      *reinterpret_cast<uint32_t*>(sq + tile_off(row0 + gq + 8, nt) + 4 * tq) = pack_bf16(o_acc[nt][2] * inv_hi, o_acc[nt][3] * inv_hi);
    }
    __syncwarp();
    {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = lane + i * 32, r = row0 + (idx >> 2), c = idx & 3;
        const int rb = s_rowbase[r];
        const int dwc = s_dw[r], cls = dwc >> 8;            // pad rows are cropped (layers.py:287-288); halo rows
        __nv_bfloat16* dstp = cls == 1 ? out : (cls == 2 ? halo_out : nullptr);   // belong to a neighbour
        if (dstp != nullptr) {
          const uint4 val = *reinterpret_cast<const uint4*>(sq + tile_off(r, c));
          *reinterpret_cast<uint4*>(dstp + token_of(l, dwc, rb) * C + head * kHeadDim + c * 8) = val;
        }
      }
    }
    __syncthreads();                                        // buffer b may be refilled by the next prefetch
  }
  cp_async_wait<0>();
}

}  // namespace attn

template <typename TB, bool PRE>
static int launch_attn_t(const void* qkv, const void* halo_qkv, const void* halo_lo_qkv, const float* qkv_bias,
                         const void* earth_bias, void* out, void* halo_out, const WinGeom& g, const BandGeom& bd,
                         int roll, int lon_chunk, dim3 grid, cudaStream_t st, float* lse) {
  using namespace attn;
  auto kern = window_attention_bf16_kernel<TB, PRE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) { set_error("attention_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
  kern<<<grid, kThreads, kSmemBytes, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)halo_qkv,
                                           (const __nv_bfloat16*)halo_lo_qkv, qkv_bias, (const TB*)earth_bias,
                                           (__nv_bfloat16*)out, (__nv_bfloat16*)halo_out, g, bd, roll, lon_chunk, lse);
  return check_launch("window_attention_bf16");
}

// tc_attention2.cu: tcgen05 / TMEM formulation (bf16 pre-scaled operands)
int launch_window_attention_tc(const void* qkv, const void* halo_qkv, const void* halo_lo_qkv, const float* qkv_bias,
                               const void* earth_bias, void* out, void* halo_out, const WinGeom& g, const BandGeom& bd,
                               int roll, cudaStream_t st, float* lse, int exact_max);

int launch_window_attention_bf16(const void* qkv, const void* halo_qkv, const void* halo_lo_qkv, const float* qkv_bias, const void* earth_bias,
                                 int bias_dtype, void* out, void* halo_out, const WinGeom& g, const BandGeom& bd,
                                 int roll, int prescaled, cudaStream_t st, float* lse) {
  using namespace attn;
  if (bd.nhw <= 0) return PANGU_OK;
  {   // pre-scaled bf16 operands: the tcgen05 kernel; $PANGU_B200_ATTN_TC=0 keeps the mma.sync kernel below
    static const bool use_tc = []() { const char* e = getenv("PANGU_B200_ATTN_TC"); return e == nullptr || atoi(e) != 0; }();
    if (use_tc && (prescaled & 1) && bias_dtype == PANGU_BF16)
      return launch_window_attention_tc(qkv, halo_qkv, halo_lo_qkv, qkv_bias, earth_bias, out, halo_out, g, bd, roll, st, lse, (prescaled >> 1) & 1);
    prescaled &= 1;
  }
  if (bd.halo_kv) { set_error("attention_bf16: K/V-only halos are only implemented in the tcgen05 kernel"); return PANGU_ERR_UNSUPPORTED; }
  // longitude windows per CTA (they share the staged bias tile): as many as possible (<= 5) while the grid
  // still fills the GPU for at least ~4 waves of 2 CTAs/SM -- small latitude bands get finer CTAs
  int lon_chunk = 1;
  const long long per_l = (long long)g.heads * g.nZ * bd.nhw;          // CTAs per longitude window column
  const int cands[4] = {5, 3, 2, 1};
  for (int c : cands) {
    if (g.nLon % c) continue;
    lon_chunk = c;
    if (per_l * (g.nLon / c) >= 8LL * tc_num_sms()) break;
  }
  dim3 grid((unsigned)g.heads, (unsigned)((g.nLon + lon_chunk - 1) / lon_chunk), (unsigned)(g.nZ * bd.nhw));
  if (bias_dtype == PANGU_BF16)
    return prescaled ? launch_attn_t<__nv_bfloat16, true>(qkv, halo_qkv, halo_lo_qkv, qkv_bias, earth_bias, out, halo_out, g, bd, roll, lon_chunk, grid, st, lse)
                     : launch_attn_t<__nv_bfloat16, false>(qkv, halo_qkv, halo_lo_qkv, qkv_bias, earth_bias, out, halo_out, g, bd, roll, lon_chunk, grid, st, lse);
  if (bias_dtype == PANGU_F32)
    return prescaled ? launch_attn_t<float, true>(qkv, halo_qkv, halo_lo_qkv, qkv_bias, earth_bias, out, halo_out, g, bd, roll, lon_chunk, grid, st, lse)
                     : launch_attn_t<float, false>(qkv, halo_qkv, halo_lo_qkv, qkv_bias, earth_bias, out, halo_out, g, bd, roll, lon_chunk, grid, st, lse);
  set_error("attention_bf16: unknown bias dtype %d", bias_dtype);
  return PANGU_ERR_BAD_ARG;
}

}  // namespace pangu
