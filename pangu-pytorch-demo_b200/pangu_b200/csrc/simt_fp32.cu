// fp32 parity path: CUDA-core FFMA kernels (no tensor cores, so rel-L2 <= 1e-5 against the
// reference's fp32 is reachable).  This is the correctness path, not the benchmarked one.
#include "common.cuh"

namespace pangu {

// ------------------------------------------------------------------------------------------
// SGEMM: out[M,N] = act(A[M,K] . W[N,K]^T + bias).  Both operands K-major (nn.Linear layout).
// 128x64 CTA tile, BK=16, 256 threads, 8x4 outputs per thread, float4 global loads.
// ------------------------------------------------------------------------------------------
constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16;

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <int kAct>
__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W,
                const float* __restrict__ bias, float* __restrict__ out, long long ldo,
                long long M, int K, int N) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // global->smem mapping: A tile 128 rows x 4 float4; thread loads rows r, r+64
  const int a_row = tid >> 2, a_k4 = (tid & 3) * 4;
  const int b_row = tid >> 2, b_k4 = (tid & 3) * 4;   // 64 rows x 4 float4

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = a_row + h * 64;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) v = __ldg(reinterpret_cast<const float4*>(A + (m0 + r) * lda + k0 + a_k4));
      As[a_k4 + 0][r] = v.x; As[a_k4 + 1][r] = v.y; As[a_k4 + 2][r] = v.z; As[a_k4 + 3][r] = v.w;
    }
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + b_row < N) v = __ldg(reinterpret_cast<const float4*>(W + (long long)(n0 + b_row) * K + k0 + b_k4));
      Bs[b_k4 + 0][b_row] = v.x; Bs[b_k4 + 1][b_row] = v.y; Bs[b_k4 + 2][b_row] = v.z; Bs[b_k4 + 3][b_row] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  float bv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx * 4 + j;
    if (bias != nullptr && n < N) bv[j] = __ldg(bias + n);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + ty * 8 + i;
    if (m >= M) continue;
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[i][j] + bv[j];
      if (kAct == PANGU_ACT_GELU_ERF) v = gelu_erf(v);
      r[j] = v;
    }
    const int n = n0 + tx * 4;
    if (n + 3 < N && (ldo % 4 == 0)) {
      *reinterpret_cast<float4*>(out + m * ldo + n) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < N) out[m * ldo + n + j] = r[j];
    }
  }
}

int launch_sgemm(const float* A, long long lda, const float* W, const float* bias, float* out,
                 long long ldo, long long M, int K, int N, int act, cudaStream_t st) {
  if (K % SG_BK != 0 || lda % 4 != 0) { set_error("sgemm: K=%d must be a multiple of 16 and lda of 4", K); return PANGU_ERR_BAD_ARG; }
  if (M <= 0 || N <= 0) return PANGU_OK;
  dim3 grid((unsigned)((M + SG_BM - 1) / SG_BM), (unsigned)((N + SG_BN - 1) / SG_BN));
  if (act == PANGU_ACT_GELU_ERF)
    sgemm_nt_kernel<PANGU_ACT_GELU_ERF><<<grid, 256, 0, st>>>(A, lda, W, bias, out, ldo, M, K, N);
  else
    sgemm_nt_kernel<PANGU_ACT_NONE><<<grid, 256, 0, st>>>(A, lda, W, bias, out, ldo, M, K, N);
  return check_launch("sgemm");
}

// ------------------------------------------------------------------------------------------
// x_out = residual + LN(y) * gamma + beta ; one warp per row, two-pass statistics in registers.
// ------------------------------------------------------------------------------------------
template <typename TY, int kPerLane>
__global__ void __launch_bounds__(256)
ln_residual_kernel(const TY* __restrict__ y, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ residual,
                   float* __restrict__ x_out, __nv_bfloat16* __restrict__ x_out_bf16, long long M,
                   float eps) {
  constexpr int C = kPerLane * 32;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float v[kPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    v[i] = to_f32<TY>(y[row * C + i * 32 + lane]);
    s += v[i];
  }
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    const int c = i * 32 + lane;
    float o = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    if (residual != nullptr) o += residual[row * C + c];
    if (x_out != nullptr) x_out[row * C + c] = o;
    if (x_out_bf16 != nullptr) x_out_bf16[row * C + c] = __float2bfloat16_rn(o);
  }
}

template <typename TY>
static int launch_ln_t(const TY* y, const float* gamma, const float* beta, const float* residual,
                       float* x_out, __nv_bfloat16* xb, long long M, int C, float eps, cudaStream_t st) {
  const int warps = 8;
  const unsigned blocks = (unsigned)((M + warps - 1) / warps);
  switch (C) {
    case 192: ln_residual_kernel<TY, 6><<<blocks, warps * 32, 0, st>>>(y, gamma, beta, residual, x_out, xb, M, eps); break;
    case 384: ln_residual_kernel<TY, 12><<<blocks, warps * 32, 0, st>>>(y, gamma, beta, residual, x_out, xb, M, eps); break;
    case 768: ln_residual_kernel<TY, 24><<<blocks, warps * 32, 0, st>>>(y, gamma, beta, residual, x_out, xb, M, eps); break;
    default: set_error("ln_residual: unsupported C=%d (192/384/768)", C); return PANGU_ERR_UNSUPPORTED;
  }
  return check_launch("ln_residual");
}

int launch_ln_residual(const void* y, int y_dtype, const float* gamma, const float* beta,
                       const float* residual, float* x_out, void* xb, long long M, int C, float eps,
                       cudaStream_t st) {
  if (M <= 0) return PANGU_OK;
  if (y_dtype == PANGU_F32)
    return launch_ln_t<float>((const float*)y, gamma, beta, residual, x_out, (__nv_bfloat16*)xb, M, C, eps, st);
  return launch_ln_t<__nv_bfloat16>((const __nv_bfloat16*)y, gamma, beta, residual, x_out, (__nv_bfloat16*)xb, M, C, eps, st);
}

// ------------------------------------------------------------------------------------------
// fp32 window attention, one CTA per (window, head).  q/k/v gathered straight from the token-order
// qkv tensor with pad/roll folded into the address; pad rows take the linear1 bias.
// ------------------------------------------------------------------------------------------
constexpr int ATT_WARPS = 8;

__global__ void __launch_bounds__(ATT_WARPS * 32)
window_attention_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias,
                            const float* __restrict__ earth_bias, float* __restrict__ out,
                            WinGeom g, int roll) {
  extern __shared__ __align__(16) float smem[];
  float* sq = smem;                                  // [144][32]
  float* sk = sq + kWinTokens * 32;                  // [144][33]
  float* sv = sk + kWinTokens * 33;                  // [144][32]
  float* sp = sv + kWinTokens * 32;                  // [ATT_WARPS][144]
  long long* ssrc = reinterpret_cast<long long*>(sp + ATT_WARPS * kWinTokens);   // [144]
  int* sgid = reinterpret_cast<int*>(ssrc + kWinTokens);                          // [144]

  const int lt = blockIdx.x;
  const int l = lt / g.T, t = lt - l * g.T;
  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = g.C;
  const float scale = rsqrtf((float)kHeadDim);       // (dim // heads) ** -0.5, layers.py:338

  for (int k = warp; k < kWinTokens; k += ATT_WARPS) {
    const long long n = window_source(g, l, t, k, roll);
    float q, kk, v;
    if (n >= 0) {
      const float* row = qkv + n * 3 * C + head * kHeadDim + lane;
      q = row[0]; kk = row[C]; v = row[2 * C];
    } else {                                         // zero pad row -> linear1(0) = bias
      q = qkv_bias[head * kHeadDim + lane];
      kk = qkv_bias[C + head * kHeadDim + lane];
      v = qkv_bias[2 * C + head * kHeadDim + lane];
    }
    sq[k * 32 + lane] = q * scale;                   // query = query * scale (layers.py:431)
    sk[k * 33 + lane] = kk;
    sv[k * 32 + lane] = v;
    if (lane == 0) { ssrc[k] = n; sgid[k] = shift_group(g, t, k); }
  }
  __syncthreads();

  const float* brow = earth_bias + ((long long)t * g.heads + head) * kWinTokens * kWinTokens;
  float* myp = sp + warp * kWinTokens;
  for (int i = warp; i < kWinTokens; i += ATT_WARPS) {
    float s[5];
    float mx = -INFINITY;
    const int gi = sgid[i];
#pragma unroll
    for (int jj = 0; jj < 5; ++jj) {
      const int j = jj * 32 + lane;
      float a = -INFINITY;
      if (j < kWinTokens) {
        a = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) a = fmaf(sq[i * 32 + d], sk[j * 33 + d], a);
        a += __ldg(brow + i * kWinTokens + j);
        if (roll == 1 && sgid[j] != gi) a += kMaskValue;
      }
      s[jj] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 5; ++jj) {
      const int j = jj * 32 + lane;
      const float e = (j < kWinTokens) ? expf(s[jj] - mx) : 0.f;
      s[jj] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int jj = 0; jj < 5; ++jj) {
      const int j = jj * 32 + lane;
      if (j < kWinTokens) myp[j] = s[jj] * inv;
    }
    __syncwarp();
    float o = 0.f;
#pragma unroll 8
    for (int j = 0; j < kWinTokens; ++j) o = fmaf(myp[j], sv[j * 32 + lane], o);
    const long long n = ssrc[i];
    if (n >= 0) out[n * C + head * kHeadDim + lane] = o;
    __syncwarp();
  }
}

int launch_window_attention_f32(const float* qkv, const float* qkv_bias, const float* earth_bias,
                                float* out, const WinGeom& g, int roll, cudaStream_t st) {
  const size_t smem = (size_t)(kWinTokens * (32 + 33 + 32) + ATT_WARPS * kWinTokens) * sizeof(float) +
                      kWinTokens * (sizeof(long long) + sizeof(int));
  static unsigned long long configured = 0;
  {
    cudaError_t e = pangu::set_max_smem_once(configured, window_attention_f32_kernel, (int)smem);
    if (e != cudaSuccess) { set_error("attention_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PANGU_ERR_CUDA; }
  }
  dim3 grid((unsigned)(g.nLon * g.T), (unsigned)g.heads);
  window_attention_f32_kernel<<<grid, ATT_WARPS * 32, smem, st>>>(qkv, qkv_bias, earth_bias, out, g, roll);
  return check_launch("window_attention_f32");
}

}  // namespace pangu
