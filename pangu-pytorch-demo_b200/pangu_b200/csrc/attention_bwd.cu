// Backward of the bf16 3-D window attention (autograd of models/layers.py:431-478 inside a block's
// pad / roll / partition / reverse / crop, :224-293), the counterpart of tc_attention.cu.
//
// One CTA owns ONE (window type t, head) bias tile and walks a chunk of longitude windows.  Per (window, head),
// with the forward's log2-sum-exp L_i of every score row (written by the forward kernel) and D_i = dO_i . O_i, the
// two passes run CONCURRENTLY on the same staged tiles, nine warps each.  ncu (profiles/r1_attention_bwd.md): the kernel
// sits at ~60 % of the shared-memory wavefront peak (ldmatrix operand fragments re-read by every warp, the transposed
// bias reads, the dBias read-modify-write) with HMMA at 22 %; three row tiles per warp halve the wavefronts but leave
// 6 warps per SM and run slower (latency), so the next step for this kernel is tcgen05 like the forward.
//   row pass   (warps 0-8, 16 query rows each):  S = q k^T + bias (+mask), P = exp2(S - L), dP = dO v^T,
//                                       dS = P (dP - D), dQ = scale * dS k,  dBias += dS into an fp32 [144 x 144] tile
//                                       in shared memory that lives across the whole longitude walk (each warp owns
//                                       its 16 rows: plain read-modify-write; one reduction per CTA at the end)
//   column pass(warps 9-17, 16 key rows each):   S^T, P^T, dP^T, dS^T recomputed in transposed fragments so that
//                                       dV = P^T dO and dK = dS^T q stay warp-local (no smem round trip of the
//                                       144 x 144 matrices, no atomics on the activations)
// After a CTA barrier the q / k / v tiles of the window are dead and serve as the staging tiles of dq / dk / dv.
// Gradients of real tokens are written at their un-rolled token position of dqkv [N, 3C]; zero-pad rows were
// linear1(0) = bias in the forward (layers.py:228,419), so they only feed linear1's bias gradient: the kernel sums
// dq/dk/dv over ALL window rows (real and pad) into dqkv_bias [3C], which IS that bias gradient.  q, k, bias arrive pre-scaled like in the forward (scale*log2e folded into q, log2e
// into the bias table): dq is returned w.r.t. the UN-scaled linear1 output, dk gets the matching ln2 factor.
#include <type_traits>

#include "attn_common.cuh"

namespace pangu {
namespace tc { int num_sms(); }
namespace attn_bwd {

using namespace attn;

constexpr int kPassWarps = 9;                           // 144 rows / 16
constexpr int kWarps = 2 * kPassWarps;
constexpr int kThreads = kWarps * 32;
constexpr int kDbPitch = 152;                           // fp32 per dBias row in smem: pitch % 32 == 24, so the four rows a half-warp
                                                        // touches with one LDS.64 / STS.64 fall into disjoint 8-bank groups
constexpr int kTiles = 5;                               // q, k, v, dO, O
constexpr int kBufBytes = kTiles * kTileBytes;
constexpr int kSmemBytes = kWinTokens * kBiasPitch * 2 + 2 * kBufBytes + kWinTokens * kDbPitch * 4 /*dBias*/ +
                           kWinTokens * 4 /*rowbase*/ + kWinTokens * 4 /*dw*/ + 160 /*gid*/ + 2 * kWinTokens * 4 /*L*/ +
                           kWinTokens * 4 /*D*/ + 16;
static_assert(kSmemBytes <= 227 * 1024, "attention backward: shared memory");

__global__ void __launch_bounds__(kThreads, 1)          // 18 warps: 5 on one SM sub-partition -> 16 K registers / 5 warps = 96 per thread
window_attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ qkv_bias,
                            const __nv_bfloat16* __restrict__ earth_bias, const __nv_bfloat16* __restrict__ o,
                            const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                            __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias, float* __restrict__ dqkv_bias,
                            WinGeom g, int roll, int lon_chunk) {
  extern __shared__ __align__(128) uint8_t smem[];
  __nv_bfloat16* s_bias = reinterpret_cast<__nv_bfloat16*>(smem);                       // [144][152]
  uint8_t* s_buf = smem + kWinTokens * kBiasPitch * 2;                                  // 2 x {q,k,v,dO,O}
  float* s_dbias = reinterpret_cast<float*>(s_buf + 2 * kBufBytes);                     // [144][kDbPitch] fp32
  int* s_rowbase = reinterpret_cast<int*>(s_dbias + kWinTokens * kDbPitch);
  int* s_dw = s_rowbase + kWinTokens;
  uint8_t* s_gid = reinterpret_cast<uint8_t*>(s_dw + kWinTokens);
  float* s_L = reinterpret_cast<float*>(s_gid + 160);                                   // [2][144]
  float* s_D = s_L + 2 * kWinTokens;                                                    // [144]

  const int head = blockIdx.x, lchunk = blockIdx.y, t = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = g.C;
  const int l_begin = lchunk * lon_chunk;
  const int l_end = min(g.nLon, l_begin + lon_chunk);

  for (int k = tid; k < kWinTokens; k += kThreads) {
    const int zw = t / g.nH, hw = t - zw * g.nH;
    const int dz = k / 72, r = k - dz * 72, dh = r / 12, dw = r - dh * 12;
    int z = 2 * zw + dz, h = 6 * hw + dh;
    if (roll == 1) { z += 1; if (z >= g.Z) z -= g.Z; h += 3; if (h >= g.Hp) h -= g.Hp; }
    if (roll == 2) {
      s_rowbase[k] = t * kWinTokens + k;
      s_dw[k] = 1 << 8;
    } else {
      const bool real = h < g.H;
      s_rowbase[k] = real ? (z * g.H + h) * g.W : -1;
      s_dw[k] = dw | ((real ? 1 : 0) << 8);
    }
    s_gid[k] = (uint8_t)shift_group(g, t, k);
  }
  for (int i = tid; i < kWinTokens * kDbPitch; i += kThreads) s_dbias[i] = 0.f;
  {
    const __nv_bfloat16* src = earth_bias + ((long long)t * g.heads + head) * kWinTokens * kWinTokens;
    for (int i = tid; i < kWinTokens * 18; i += kThreads) {
      const int r = i / 18, c = i - r * 18;
      cp_async16(smem_u32(s_bias + r * kBiasPitch + c * 8), src + r * kWinTokens + c * 8);
    }
  }
  __syncthreads();

  auto token_of = [&](int l, int dwc, int rb) -> long long {
    if (roll == 2) return (long long)l * g.T * kWinTokens + rb;
    int w = 12 * l + (roll == 1 ? 6 : 0) + (dwc & 0xff);
    if (w >= g.W) w -= g.W;
    return (long long)rb + w;
  };
  auto issue_load = [&](int l, int b) {
    uint8_t* buf = s_buf + b * kBufBytes;
    for (int i = tid; i < kWinTokens * 20; i += kThreads) {
      const int k = i / 20, part = i - k * 20, s = part >> 2, c = part & 3;
      uint8_t* dst = buf + s * kTileBytes + tile_off(k, c);
      const int rb = s_rowbase[k];
      const int dwc = s_dw[k];
      if ((dwc >> 8) != 0) {
        const long long tok = token_of(l, dwc, rb);
        const __nv_bfloat16* src = s < 3 ? qkv + tok * 3 * C + s * C : (s == 3 ? dout : o) + tok * C;
        cp_async16(smem_u32(dst), src + head * kHeadDim + c * 8);
      } else if (s < 3) {                                   // zero pad row: linear1(0) = bias
        const float* bsrc = qkv_bias + s * C + head * kHeadDim + c * 8;
        uint4 v;
        v.x = pack_bf16(bsrc[0], bsrc[1]); v.y = pack_bf16(bsrc[2], bsrc[3]);
        v.z = pack_bf16(bsrc[4], bsrc[5]); v.w = pack_bf16(bsrc[6], bsrc[7]);
        *reinterpret_cast<uint4*>(dst) = v;
      } else {                                              // its output is cropped: dO = 0
        *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    if (tid < kWinTokens)
      s_L[b * kWinTokens + tid] = __ldg(lse + (((long long)l * g.T + t) * g.heads + head) * kWinTokens + tid);
  };

  issue_load(l_begin, 0);
  cp_async_commit();

  const float scale = rsqrtf((float)kHeadDim);
  const float ln2 = 0.6931471805599453f;
  const float mask_l2 = kMaskValue * kLog2e;
  const int gq = lane >> 2, tq = lane & 3;
  const bool row_role = warp < kPassWarps;                  // warp-uniform
  const int row0 = (row_role ? warp : warp - kPassWarps) * 16;
  const int mi = lane >> 3, mr = lane & 7;
  const bool masked_type = roll == 1 && ((t / g.nH == g.nZ - 1) || (t % g.nH == g.nH - 1));
  const int gid_lo = s_gid[row0 + gq], gid_hi = s_gid[row0 + gq + 8];       // valid after the barrier above

  float colsum[2][4][2];                                    // row role: [0] = dq; column role: [0] = dk, [1] = dv
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) colsum[a][b][0] = colsum[a][b][1] = 0.f;

  for (int l = l_begin; l < l_end; ++l) {
    const int b = (l - l_begin) & 1;
    if (l + 1 < l_end) issue_load(l + 1, b ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    uint8_t* sq = s_buf + b * kBufBytes;
    uint8_t* sk = sq + kTileBytes;
    uint8_t* sv = sk + kTileBytes;
    uint8_t* sdo = sv + kTileBytes;
    uint8_t* so = sdo + kTileBytes;
    const float* sL = s_L + b * kWinTokens;

    if (tid < kWinTokens) {                                  // D_i = dO_i . O_i
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 a4 = *reinterpret_cast<const uint4*>(sdo + tile_off(tid, c));
        const uint4 b4 = *reinterpret_cast<const uint4*>(so + tile_off(tid, c));
        const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a4);
        const __nv_bfloat162* bp = reinterpret_cast<const __nv_bfloat162*>(&b4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = __bfloat1622float2(ap[j]), y = __bfloat1622float2(bp[j]);
          acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
        }
      }
      s_D[tid] = acc;
    }
    __syncthreads();

    // stage a 16 x 32 fragment tile (this warp's rows) through a dead operand tile and write tensor s of dqkv
    auto emit = [&](float (&acc)[4][4], auto s_tag, float mult, uint8_t* stage, float (&cs)[4][2]) {
      constexpr int s = decltype(s_tag)::value;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float v0 = acc[nt][0] * mult, v1 = acc[nt][1] * mult, v2 = acc[nt][2] * mult, v3 = acc[nt][3] * mult;
        *reinterpret_cast<uint32_t*>(stage + tile_off(row0 + gq, nt) + 4 * tq) = pack_bf16(v0, v1);
        *reinterpret_cast<uint32_t*>(stage + tile_off(row0 + gq + 8, nt) + 4 * tq) = pack_bf16(v2, v3);
        cs[nt][0] += v0 + v2;                               // every window row, pad rows included, is a row of linear1's output
        cs[nt][1] += v1 + v3;
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = lane + i * 32, r = row0 + (idx >> 2), c = idx & 3;
        const int dwc = s_dw[r];
        if ((dwc >> 8) != 0) {
          const uint4 val = *reinterpret_cast<const uint4*>(stage + tile_off(r, c));
          *reinterpret_cast<uint4*>(dqkv + token_of(l, dwc, s_rowbase[r]) * 3 * C + s * C + head * kHeadDim + c * 8) = val;
        }
      }
    };

    float acc_a[4][4], acc_b[4][4];                         // row role: dq in acc_a; column role: dk in acc_a, dv in acc_b
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc_a[i][j] = acc_b[i][j] = 0.f;

    // ------------------------------------------------------------------ row pass: dQ, dBias
    if (row_role) {
      uint32_t qa[2][4], da[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int r = row0 + (mi & 1) * 8 + mr, c = ks * 2 + (mi >> 1);
        ldmatrix_x4(smem_u32(sq + tile_off(r, c)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
        ldmatrix_x4(smem_u32(sdo + tile_off(r, c)), da[ks][0], da[ks][1], da[ks][2], da[ks][3]);
      }
      const float L_lo = sL[row0 + gq], L_hi = sL[row0 + gq + 8];
      const float D_lo = s_D[row0 + gq], D_hi = s_D[row0 + gq + 8];
      float (&dq_acc)[4][4] = acc_a;
#pragma unroll 1
      for (int kvb = 0; kvb < 3; ++kvb) {
        const int kv0 = kvb * 48;
        uint32_t dsa[3][4];
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
          const int j = kv0 + nt * 8 + 2 * tq;
          float s_acc[4], dp[4] = {0.f, 0.f, 0.f, 0.f};
          {
            const __nv_bfloat162 b_lo = *reinterpret_cast<const __nv_bfloat162*>(s_bias + (row0 + gq) * kBiasPitch + j);
            const __nv_bfloat162 b_hi = *reinterpret_cast<const __nv_bfloat162*>(s_bias + (row0 + gq + 8) * kBiasPitch + j);
            s_acc[0] = __low2float(b_lo); s_acc[1] = __high2float(b_lo);
            s_acc[2] = __low2float(b_hi); s_acc[3] = __high2float(b_hi);
          }
          uint32_t k0, k1, k2, k3;
          ldmatrix_x4(smem_u32(sk + tile_off(kv0 + nt * 8 + mr, mi)), k0, k1, k2, k3);
          mma_bf16(s_acc, qa[0], k0, k1);
          mma_bf16(s_acc, qa[1], k2, k3);
          ldmatrix_x4(smem_u32(sv + tile_off(kv0 + nt * 8 + mr, mi)), k0, k1, k2, k3);
          mma_bf16(dp, da[0], k0, k1);
          mma_bf16(dp, da[1], k2, k3);
          if (masked_type) {
            const int g0 = s_gid[j], g1 = s_gid[j + 1];
            if (g0 != gid_lo) s_acc[0] += mask_l2;
            if (g1 != gid_lo) s_acc[1] += mask_l2;
            if (g0 != gid_hi) s_acc[2] += mask_l2;
            if (g1 != gid_hi) s_acc[3] += mask_l2;
          }
          const float ds0 = ex2(s_acc[0] - L_lo) * (dp[0] - D_lo), ds1 = ex2(s_acc[1] - L_lo) * (dp[1] - D_lo);
          const float ds2 = ex2(s_acc[2] - L_hi) * (dp[2] - D_hi), ds3 = ex2(s_acc[3] - L_hi) * (dp[3] - D_hi);
          {                                                 // dBias += dS: this warp owns rows row0 .. row0+15 of the tile
            float2* p_lo = reinterpret_cast<float2*>(s_dbias + (row0 + gq) * kDbPitch + j);
            float2* p_hi = reinterpret_cast<float2*>(s_dbias + (row0 + gq + 8) * kDbPitch + j);
            float2 a_lo = *p_lo, a_hi = *p_hi;
            a_lo.x += ds0; a_lo.y += ds1; a_hi.x += ds2; a_hi.y += ds3;
            *p_lo = a_lo; *p_hi = a_hi;
          }
          dsa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(ds0, ds1);
          dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(ds2, ds3);
        }
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
#pragma unroll
          for (int dp2 = 0; dp2 < 2; ++dp2) {
            uint32_t k0, k1, k2, k3;
            const int r = kv0 + kk * 16 + (mi & 1) * 8 + mr, c = dp2 * 2 + (mi >> 1);
            ldmatrix_x4_trans(smem_u32(sk + tile_off(r, c)), k0, k1, k2, k3);
            mma_bf16(dq_acc[dp2 * 2], dsa[kk], k0, k1);
            mma_bf16(dq_acc[dp2 * 2 + 1], dsa[kk], k2, k3);
          }
        }
      }
    } else {
    // ------------------------------------------------------------------ column pass: dK, dV (this warp's 16 KEY rows)
      uint32_t ka[2][4], va[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int r = row0 + (mi & 1) * 8 + mr, c = ks * 2 + (mi >> 1);
        ldmatrix_x4(smem_u32(sk + tile_off(r, c)), ka[ks][0], ka[ks][1], ka[ks][2], ka[ks][3]);
        ldmatrix_x4(smem_u32(sv + tile_off(r, c)), va[ks][0], va[ks][1], va[ks][2], va[ks][3]);
      }
      float (&dk_acc)[4][4] = acc_a;
      float (&dv_acc)[4][4] = acc_b;
      const int jl = row0 + gq, jh = row0 + gq + 8;
#pragma unroll 1
      for (int ib = 0; ib < 3; ++ib) {
        const int i0 = ib * 48;
        uint32_t pta[3][4], dsta[3][4];
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
          const int i = i0 + nt * 8 + 2 * tq;
          float st[4], dpt[4] = {0.f, 0.f, 0.f, 0.f};
          st[0] = __bfloat162float(s_bias[i * kBiasPitch + jl]);
          st[1] = __bfloat162float(s_bias[(i + 1) * kBiasPitch + jl]);
          st[2] = __bfloat162float(s_bias[i * kBiasPitch + jh]);
          st[3] = __bfloat162float(s_bias[(i + 1) * kBiasPitch + jh]);
          uint32_t q0, q1, q2, q3;
          ldmatrix_x4(smem_u32(sq + tile_off(i0 + nt * 8 + mr, mi)), q0, q1, q2, q3);
          mma_bf16(st, ka[0], q0, q1);
          mma_bf16(st, ka[1], q2, q3);
          ldmatrix_x4(smem_u32(sdo + tile_off(i0 + nt * 8 + mr, mi)), q0, q1, q2, q3);
          mma_bf16(dpt, va[0], q0, q1);
          mma_bf16(dpt, va[1], q2, q3);
          if (masked_type) {
            const int g0 = s_gid[i], g1 = s_gid[i + 1];
            if (g0 != gid_lo) st[0] += mask_l2;
            if (g1 != gid_lo) st[1] += mask_l2;
            if (g0 != gid_hi) st[2] += mask_l2;
            if (g1 != gid_hi) st[3] += mask_l2;
          }
          const float2 Li = *reinterpret_cast<const float2*>(sL + i), Di = *reinterpret_cast<const float2*>(s_D + i);   // i is even
          const float Li0 = Li.x, Li1 = Li.y, Di0 = Di.x, Di1 = Di.y;
          const float p0 = ex2(st[0] - Li0), p1 = ex2(st[1] - Li1), p2 = ex2(st[2] - Li0), p3 = ex2(st[3] - Li1);
          pta[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
          pta[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
          dsta[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0 * (dpt[0] - Di0), p1 * (dpt[1] - Di1));
          dsta[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2 * (dpt[2] - Di0), p3 * (dpt[3] - Di1));
        }
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
#pragma unroll
          for (int dp2 = 0; dp2 < 2; ++dp2) {
            uint32_t x0, x1, x2, x3;
            const int r = i0 + kk * 16 + (mi & 1) * 8 + mr, c = dp2 * 2 + (mi >> 1);
            ldmatrix_x4_trans(smem_u32(sdo + tile_off(r, c)), x0, x1, x2, x3);
            mma_bf16(dv_acc[dp2 * 2], pta[kk], x0, x1);
            mma_bf16(dv_acc[dp2 * 2 + 1], pta[kk], x2, x3);
            ldmatrix_x4_trans(smem_u32(sq + tile_off(r, c)), x0, x1, x2, x3);
            mma_bf16(dk_acc[dp2 * 2], dsta[kk], x0, x1);
            mma_bf16(dk_acc[dp2 * 2 + 1], dsta[kk], x2, x3);
          }
        }
      }
    }
    __syncthreads();                                        // both passes are done with q / k / v: the tiles become staging
    if (row_role) {
      emit(acc_a, std::integral_constant<int, 0>{}, scale, sq, colsum[0]);
    } else {
      emit(acc_a, std::integral_constant<int, 1>{}, ln2, sk, colsum[0]);
      emit(acc_b, std::integral_constant<int, 2>{}, 1.0f, sv, colsum[1]);
    }
    __syncthreads();                                        // buffer b may be refilled by the next prefetch
  }
  cp_async_wait<0>();

  // ---- per-CTA reductions: Earth-specific bias gradient (dS summed over this CTA's longitude windows) ...
  {
    float* dst = dbias + ((long long)t * g.heads + head) * kWinTokens * kWinTokens;
    for (int i = tid; i < kWinTokens * (kWinTokens / 4); i += kThreads) {     // 16-byte vector reductions: 5184 per CTA
      const int r = i / (kWinTokens / 4), c = (i - r * (kWinTokens / 4)) * 4;
      const float4 v = *reinterpret_cast<const float4*>(s_dbias + r * kDbPitch + c);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + r * kWinTokens + c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
  }
  // ---- ... and linear1's bias gradient: column sums of dq / dk / dv over all rows this CTA produced
  {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      if (row_role && a == 1) break;
      const int sidx = row_role ? 0 : 1 + a;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float v = colsum[a][nt][e];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (gq == 0) atomicAdd(dqkv_bias + sidx * C + head * kHeadDim + nt * 8 + 2 * tq + e, v);
        }
    }
  }
}

}  // namespace attn_bwd

int launch_window_attention_bwd(const void* qkv, const float* qkv_bias, const void* earth_bias, const void* o,
                                const void* dout, const float* lse, void* dqkv, float* dbias, float* dpad,
                                const WinGeom& g, int roll, cudaStream_t st) {
  using namespace attn_bwd;
  static unsigned long long configured = 0;
  {
    cudaError_t e = pangu::set_max_smem_once(configured, window_attention_bwd_kernel, kSmemBytes);
    if (e != cudaSuccess) { set_error("attention_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
  }
  // longitude windows per CTA: they share the staged bias tile and the fragment-resident dBias accumulators (one
  // reduction of 144 x 144 values per CTA), so chunks are long; still >= ~10 waves of one CTA per SM
  int lon_chunk = g.nLon;
  const long long per_l = (long long)g.heads * g.T;
  for (int c = g.nLon; c >= 1; --c) {
    if (g.nLon % c) continue;
    lon_chunk = c;
    if (per_l * (g.nLon / c) >= 10LL * tc::num_sms()) break;
  }
  dim3 grid((unsigned)g.heads, (unsigned)(g.nLon / lon_chunk), (unsigned)g.T);
  window_attention_bwd_kernel<<<grid, kThreads, kSmemBytes, st>>>(
      (const __nv_bfloat16*)qkv, qkv_bias, (const __nv_bfloat16*)earth_bias, (const __nv_bfloat16*)o,
      (const __nv_bfloat16*)dout, lse, (__nv_bfloat16*)dqkv, dbias, dpad, g, roll, lon_chunk);
  return check_launch("window_attention_bwd");
}

}  // namespace pangu
