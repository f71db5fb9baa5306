// extern "C" entry points of libpangu_b200.so that dispatch between the fp32 SIMT path and the
// bf16 tcgen05 path, plus error plumbing.  See include/pangu_b200.h for the contract.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace pangu {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return PANGU_ERR_CUDA;
  }
  return PANGU_OK;
}

// simt_fp32.cu
int launch_sgemm(const float* A, long long lda, const float* W, const float* bias, float* out,
                 long long ldo, long long M, int K, int N, int act, cudaStream_t st);
int launch_ln_residual(const void* y, int y_dtype, const float* gamma, const float* beta,
                       const float* residual, float* x_out, void* xb, long long M, int C, float eps,
                       cudaStream_t st);
int launch_window_attention_f32(const float* qkv, const float* qkv_bias, const float* earth_bias,
                                float* out, const WinGeom& g, int roll, cudaStream_t st);
// tc_gemm.cu
int launch_tc_linear(const void* A, long long lda, const void* W, const float* bias, void* out,
                     long long ldo, long long M, int K, int N, int act, int out_dtype, cudaStream_t st,
                     void* shadow = nullptr, const void* A2 = nullptr, long long lda2 = 0, int K1 = 0,
                     const float* addend = nullptr);
int launch_tc_linear_ln(const void* A, long long lda, const void* W, const float* bias,
                        const float* gamma, const float* beta, const float* residual, float* x_out,
                        void* x_out_bf16, long long M, int K, int C, float eps, cudaStream_t st);
// tc_gemm2.cu
int launch_tc_linear_pair(const void* A, long long lda, const void* W, const float* bias, void* out,
                          long long ldo, long long M, int K, int N, int act, int out_dtype, void* shadow, cudaStream_t st,
                          void* aux, int aux_mode, float* colsum);
// tc_mlp.cu
int launch_tc_attn_proj_mlp(const void* o, const void* wp, const float* bp, const float* gamma1, const float* beta1,
                            const float* x_in, const void* w1, const float* b1, const void* w2, const float* b2,
                            const float* gamma2, const float* beta2, float* scratch, long long scratch_rows, float* x_out,
                            void* x_out_bf16, long long M, int C, float eps1, float eps2, cudaStream_t st);
int launch_tc_mlp(const void* x, const void* w1, const float* b1, const void* w2, const float* b2,
                  const float* gamma, const float* beta, const float* residual, float* x_out,
                  void* x_out_bf16, long long M, int C, float eps, cudaStream_t st);
int debug_read_mlp_trace(long long* out, int n);
// tc_attention.cu
int launch_window_attention_bf16(const void* qkv, const void* halo_qkv, const void* halo_lo_qkv, const float* qkv_bias, const void* earth_bias,
                                 int bias_dtype, void* out, void* halo_out, const WinGeom& g, const BandGeom& bd,
                                 int roll, int prescaled, cudaStream_t st, float* lse = nullptr);

// tc_attention2.cu
int debug_read_attn_trace(long long* out, int n);
// attention_bwd.cu
int launch_window_attention_bwd(const void* qkv, const float* qkv_bias, const void* earth_bias, const void* o,
                                const void* dout, const float* lse, void* dqkv, float* dbias, float* dpad,
                                const WinGeom& g, int roll, cudaStream_t st);

}  // namespace pangu

namespace pangu { namespace tc { int g_pdl_runtime = 0; } }

using namespace pangu;

extern "C" const char* pangu_last_error(void) { return g_err; }
extern "C" int pangu_abi_version(void) { return 1; }
extern "C" int pangu_has_tcgen05(void) { return 1; }
extern "C" int pangu_set_pdl(int on) {
  const int prev = pangu::tc::g_pdl_runtime;
  pangu::tc::g_pdl_runtime = on != 0;
  return prev;
}

extern "C" int pangu_linear(const void* A, int64_t lda, const void* W, const float* bias, void* out,
                            int64_t ldo, int64_t M, int32_t K, int32_t N, int act, int dtype,
                            int out_dtype, void* stream) {
  if (!A || !W || !out || M < 0 || K <= 0 || N <= 0) { set_error("linear: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (act != PANGU_ACT_NONE && act != PANGU_ACT_GELU_ERF) { set_error("linear: unknown activation %d", act); return PANGU_ERR_BAD_ARG; }
  if (dtype == PANGU_F32) {
    if (out_dtype != PANGU_F32) { set_error("linear: fp32 path writes fp32"); return PANGU_ERR_UNSUPPORTED; }
    return launch_sgemm((const float*)A, lda, (const float*)W, bias, (float*)out, ldo, M, K, N, act, as_stream(stream));
  }
  if (dtype == PANGU_BF16)
    return launch_tc_linear(A, lda, W, bias, out, ldo, M, K, N, act, out_dtype, as_stream(stream));
  set_error("linear: unknown dtype %d", dtype);
  return PANGU_ERR_BAD_ARG;
}

extern "C" int pangu_linear_bf16_ex(const void* A, int64_t lda, const void* A2, int64_t lda2, int32_t K1,
                                    const void* W, const float* bias, void* out, void* out_bf16_shadow,
                                    int64_t ldo, int64_t M, int32_t K, int32_t N, int act, int out_dtype,
                                    void* stream) {
  if (!A || !W || !out || M < 0 || K <= 0 || N <= 0) { set_error("linear_ex: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (act != PANGU_ACT_NONE && act != PANGU_ACT_GELU_ERF) { set_error("linear_ex: unknown activation %d", act); return PANGU_ERR_BAD_ARG; }
  return launch_tc_linear(A, lda, W, bias, out, ldo, M, K, N, act, out_dtype, as_stream(stream), out_bf16_shadow, A2, lda2, K1);
}

extern "C" int pangu_linear_bf16_aux(const void* A, int64_t lda, const void* W, const float* bias, void* out, void* aux,
                                     int64_t ldo, int64_t M, int32_t K, int32_t N, int act, int aux_mode, float* colsum,
                                     void* stream) {
  if (!A || !W || !out || !aux || M < 0 || K <= 0 || N <= 0) { set_error("linear_bf16_aux: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (aux_mode != PANGU_AUX_PRE_OUT && aux_mode != PANGU_AUX_GELU_BWD) { set_error("linear_bf16_aux: unknown aux mode %d", aux_mode); return PANGU_ERR_BAD_ARG; }
  if (act != PANGU_ACT_NONE && act != PANGU_ACT_GELU_ERF) { set_error("linear_bf16_aux: unknown activation %d", act); return PANGU_ERR_BAD_ARG; }
  if (M == 0) return PANGU_OK;
  const int rc = launch_tc_linear_pair(A, lda, W, bias, out, ldo, M, K, N, act, PANGU_BF16, nullptr, as_stream(stream), aux, aux_mode, colsum);
  if (rc == PANGU_ERR_UNSUPPORTED) set_error("linear_bf16_aux: shape M=%lld K=%d N=%d is not covered by the CTA-pair GEMM (K in {192,384}, N %% 192 == 0, M >= 2048)", (long long)M, K, N);
  return rc;
}

extern "C" int pangu_ln_residual(const void* y, int y_dtype, const float* gamma, const float* beta,
                                 const float* residual, float* x_out, void* x_out_bf16, int64_t M,
                                 int32_t C, float eps, void* stream) {
  if (!y || !gamma || !beta || (!x_out && !x_out_bf16) || M < 0) { set_error("ln_residual: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (y_dtype != PANGU_F32 && y_dtype != PANGU_BF16) { set_error("ln_residual: unknown dtype"); return PANGU_ERR_BAD_ARG; }
  return launch_ln_residual(y, y_dtype, gamma, beta, residual, x_out, x_out_bf16, M, C, eps, as_stream(stream));
}

extern "C" int pangu_linear_ln_residual_bf16(const void* A, int64_t lda, const void* W, const float* bias,
                                             const float* gamma, const float* beta, const float* residual,
                                             float* x_out, void* x_out_bf16, int64_t M, int32_t K,
                                             int32_t C, float eps, void* stream) {
  if (!A || !W || !gamma || !beta || !x_out || M < 0 || K <= 0) { set_error("linear_ln_residual: bad argument"); return PANGU_ERR_BAD_ARG; }
  return launch_tc_linear_ln(A, lda, W, bias, gamma, beta, residual, x_out, x_out_bf16, M, K, C, eps, as_stream(stream));
}

extern "C" int pangu_debug_mlp_trace(int64_t* out, int32_t n) {
  return debug_read_mlp_trace(reinterpret_cast<long long*>(out), n);
}

extern "C" int pangu_mlp_ln_residual_bf16(const void* x, const void* w1, const float* b1, const void* w2,
                                          const float* b2, const float* gamma, const float* beta,
                                          const float* residual, float* x_out, void* x_out_bf16, int64_t M,
                                          int32_t C, float eps, void* stream) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !gamma || !beta || !x_out || M < 0) { set_error("mlp_ln_residual: bad argument"); return PANGU_ERR_BAD_ARG; }
  return launch_tc_mlp(x, w1, b1, w2, b2, gamma, beta, residual, x_out, x_out_bf16, M, C, eps, as_stream(stream));
}

extern "C" int pangu_attn_proj_mlp_bf16(const void* o, const void* w_proj, const float* b_proj, const float* gamma1,
                                       const float* beta1, const float* x_in, const void* w1, const float* b1,
                                       const void* w2, const float* b2, const float* gamma2, const float* beta2,
                                       float* scratch, int64_t scratch_rows, float* x_out, void* x_out_bf16, int64_t M,
                                       int32_t C, float eps1, float eps2, void* stream) {
  if (!o || !w_proj || !b_proj || !gamma1 || !beta1 || !x_in || !w1 || !b1 || !w2 || !b2 || !gamma2 || !beta2 || !scratch ||
      !x_out || M < 0) { set_error("attn_proj_mlp: bad argument"); return PANGU_ERR_BAD_ARG; }
  return launch_tc_attn_proj_mlp(o, w_proj, b_proj, gamma1, beta1, x_in, w1, b1, w2, b2, gamma2, beta2, scratch, scratch_rows,
                                 x_out, x_out_bf16, M, C, eps1, eps2, as_stream(stream));
}

extern "C" int pangu_window_attention(const void* qkv, const float* qkv_bias, const void* earth_bias,
                                      int bias_dtype, void* out, const pangu_geom* gg, int roll, int dtype,
                                      void* stream) {
  WinGeom g;
  if (!make_geom(gg, g) || !qkv || !qkv_bias || !earth_bias || !out) { set_error("window_attention: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (g.C != g.heads * kHeadDim) { set_error("window_attention: C=%d must equal heads*32", g.C); return PANGU_ERR_BAD_ARG; }
  if (dtype == PANGU_F32) {
    if (bias_dtype != PANGU_F32) { set_error("window_attention: fp32 path needs an fp32 bias table"); return PANGU_ERR_UNSUPPORTED; }
    return launch_window_attention_f32((const float*)qkv, qkv_bias, (const float*)earth_bias, (float*)out, g, roll, as_stream(stream));
  }
  if (dtype == PANGU_BF16) {
    const BandGeom full{0, g.H, 0, g.nH, 0, 0, 0};
    return launch_window_attention_bf16(qkv, nullptr, nullptr, qkv_bias, earth_bias, bias_dtype, out, nullptr, g, full, roll, 0, as_stream(stream));
  }
  set_error("window_attention: unknown dtype %d", dtype);
  return PANGU_ERR_BAD_ARG;
}

extern "C" int pangu_window_attention_band(const void* qkv, const void* halo_qkv, const void* halo_lo_qkv,
                                           const float* qkv_bias, const void* earth_bias, int bias_dtype,
                                           void* out, void* halo_out, const pangu_geom* gg,
                                           const pangu_band* band, int roll, int prescaled, void* stream) {
  WinGeom g;
  if (!make_geom(gg, g) || !band || !qkv || !qkv_bias || !earth_bias || !out) { set_error("window_attention_band: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (g.C != g.heads * kHeadDim) { set_error("window_attention_band: C=%d must equal heads*32", g.C); return PANGU_ERR_BAD_ARG; }
  if (roll != 0 && roll != 1) { set_error("window_attention_band: roll must be 0 or 1"); return PANGU_ERR_BAD_ARG; }
  BandGeom bd{band->h0, band->hrows, band->hw0, band->nhw, band->wrap, band->halo, band->halo_lo};
  if (prescaled & PANGU_ATTN_HALO_KV) {
    if (!(prescaled & 1) || bias_dtype != PANGU_BF16 || halo_out != nullptr) {
      set_error("window_attention_band: K/V-only halos need the pre-scaled bf16 path and no halo output");
      return PANGU_ERR_BAD_ARG;
    }
    bd.halo_kv = 1;
  }
  const int last_hw = bd.hw0 + bd.nhw - (bd.wrap ? 1 : 0);          // one past the last regular window
  if (bd.h0 < 0 || bd.hrows <= 0 || bd.h0 + bd.hrows > g.H || bd.hw0 < 0 || bd.nhw < 0 || last_hw > g.nH ||
      bd.halo < 0 || (bd.halo > 0 && !halo_qkv) || bd.halo_lo < 0 || (bd.halo_lo > 0 && !halo_lo_qkv) ||
      (bd.wrap && (roll != 1 || bd.h0 != 0))) {
    set_error("window_attention_band: inconsistent band (h0 %d rows %d hw0 %d nhw %d wrap %d halo %d/%d)", bd.h0, bd.hrows, bd.hw0, bd.nhw, bd.wrap, bd.halo, bd.halo_lo);
    return PANGU_ERR_BAD_ARG;
  }
  // every source row of the selected windows must be a pad row, an own row or a halo row
  const int shift = roll ? 3 : 0;
  for (int i = 0; i < bd.nhw - (bd.wrap ? 1 : 0); ++i) {
    const int lo = 6 * (bd.hw0 + i) + shift, hi = lo + 5;             // no wrap for regular windows (hi < Hp)
    const int hi_real = hi < g.H ? hi : g.H - 1;
    if (lo < g.H && (lo < bd.h0 - bd.halo_lo || hi_real >= bd.h0 + bd.hrows + bd.halo)) {
      set_error("window_attention_band: window %d needs rows [%d,%d] outside band [%d,%d)+%d", bd.hw0 + i, lo, hi_real, bd.h0, bd.h0 + bd.hrows, bd.halo);
      return PANGU_ERR_BAD_ARG;
    }
  }
  return launch_window_attention_bf16(qkv, halo_qkv, halo_lo_qkv, qkv_bias, earth_bias, bias_dtype, out, halo_out, g, bd, roll, prescaled & 3, as_stream(stream));
}

extern "C" int pangu_linear_bf16_add(const void* A, int64_t lda, const void* W, const float* bias, const float* addend,
                                     float* out, int64_t ldo, int64_t M, int32_t K, int32_t N, void* stream) {
  if (!A || !W || !out || M < 0 || K <= 0 || N <= 0) { set_error("linear_add: bad argument"); return PANGU_ERR_BAD_ARG; }
  return launch_tc_linear(A, lda, W, bias, out, ldo, M, K, N, PANGU_ACT_NONE, PANGU_F32, as_stream(stream), nullptr, nullptr, 0, 0, addend);
}

extern "C" int pangu_window_attention_train(const void* qkv, const float* qkv_bias, const void* earth_bias, void* out,
                                            float* lse, const pangu_geom* gg, int roll, void* stream) {
  WinGeom g;
  if (!make_geom(gg, g) || !qkv || !qkv_bias || !earth_bias || !out || !lse) { set_error("window_attention_train: bad argument"); return PANGU_ERR_BAD_ARG; }
  if (g.C != g.heads * kHeadDim) { set_error("window_attention_train: C=%d must equal heads*32", g.C); return PANGU_ERR_BAD_ARG; }
  const int exact = (roll >> 8) & 1;                              // PANGU_ROLL_EXACT_MAX
  roll &= 0xff;
  if (roll < 0 || roll > 2) { set_error("window_attention_train: roll must be 0, 1 or 2"); return PANGU_ERR_BAD_ARG; }
  const BandGeom full{0, g.H, 0, g.nH, 0, 0, 0};
  return launch_window_attention_bf16(qkv, nullptr, nullptr, qkv_bias, earth_bias, PANGU_BF16, out, nullptr, g, full, roll, 1 | (exact << 1), as_stream(stream), lse);
}

extern "C" int pangu_window_attention_backward(const void* qkv, const float* qkv_bias, const void* earth_bias,
                                               const void* out, const void* d_out, const float* lse, void* d_qkv,
                                               float* d_earth_bias, float* d_qkv_bias, const pangu_geom* gg,
                                               int roll, void* stream) {
  WinGeom g;
  if (!make_geom(gg, g) || !qkv || !qkv_bias || !earth_bias || !out || !d_out || !lse || !d_qkv || !d_earth_bias || !d_qkv_bias) {
    set_error("window_attention_backward: bad argument");
    return PANGU_ERR_BAD_ARG;
  }
  if (g.C != g.heads * kHeadDim) { set_error("window_attention_backward: C=%d must equal heads*32", g.C); return PANGU_ERR_BAD_ARG; }
  if (roll < 0 || roll > 2) { set_error("window_attention_backward: roll must be 0, 1 or 2"); return PANGU_ERR_BAD_ARG; }
  return launch_window_attention_bwd(qkv, qkv_bias, earth_bias, out, d_out, lse, d_qkv, d_earth_bias, d_qkv_bias, g, roll, as_stream(stream));
}

extern "C" int pangu_debug_attn_trace(int64_t* out, int32_t n) { return debug_read_attn_trace(reinterpret_cast<long long*>(out), n); }
