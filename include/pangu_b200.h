/*
 * pangu_b200.h -- C ABI of libpangu_b200.so: the B200 (sm_100a) kernels behind the Pangu-Weather
 * forward path of comdaze/pangu-pytorch-demo.
 *
 * The reference has NO foreign-function interface: its hot path is `PanguModel.forward`
 * (models/pangu_model.py:61-104) over the nn.Modules of models/layers.py, executed by ATen.
 * The drop-in boundary is therefore the Python class API (pangu-pytorch-demo_b200/models/), and
 * THIS header is what those classes bind with ctypes -- each entry point cites the reference lines
 * whose arithmetic it replaces.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch's allocator); `stream` is a
 *     cudaStream_t passed as void*; no entry point synchronises the host or allocates memory;
 *   - returns 0 on success, a negative pangu_status otherwise; pangu_last_error() gives the text
 *     (thread-local);
 *   - token tensors are row-major [N, C] with n = (z*H + h)*W + w (models/layers.py:116-119);
 *   - dtype: 0 = fp32 (SIMT FFMA path, the <=1e-5 parity path), 1 = bf16 operands with fp32
 *     accumulation (tcgen05/TMEM/TMA path);
 *   - the window is fixed at (2, 6, 12) = 144 tokens, head_dim 32, pad 5 rows
 *     (models/layers.py:168,178,338).
 */
#ifndef PANGU_B200_H_
#define PANGU_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PANGU_OK = 0,
  PANGU_ERR_BAD_ARG = -1,
  PANGU_ERR_CUDA = -2,
  PANGU_ERR_UNSUPPORTED = -3
} pangu_status;

typedef enum { PANGU_F32 = 0, PANGU_BF16 = 1 } pangu_dtype;

/* epilogue activation of pangu_linear */
typedef enum { PANGU_ACT_NONE = 0, PANGU_ACT_GELU_ERF = 1 } pangu_act;

/* Token grid of one stage: stage A = {8,181,360,192,6}, stage B = {8,91,180,384,12}. */
typedef struct {
  int32_t Z, H, W;   /* token grid (un-padded)                    */
  int32_t C;         /* channels                                  */
  int32_t heads;     /* C / 32                                    */
} pangu_geom;

/* Latitude band of one stage's token grid held by one rank (pangu_b200/dist.py; new -- the reference has no
 * spatial sharding).  The rank's tensors hold global rows [h0, h0+hrows); the launch covers the h-windows
 * [hw0, hw0+nhw) of the (rolled or un-rolled) padded grid, the last of them being the global wrap-around
 * window nH-1 when wrap != 0 (rolled blocks, rank 0); `halo` rows of the next rank follow the own rows in
 * separate halo buffers. */
typedef struct {
  int32_t h0, hrows;
  int32_t hw0, nhw;
  int32_t wrap;
  int32_t halo;      /* rows of the southern neighbour available after the own rows  */
  int32_t halo_lo;   /* rows of the northern neighbour available before the own rows */
} pangu_band;

const char* pangu_last_error(void);
int pangu_abi_version(void);
/* 1 if the library was built with tcgen05/TMA kernels for sm_100a (always, for this build). */
int pangu_has_tcgen05(void);
/* Programmatic dependent launch between the persistent tensor-core kernels (their set-up overlaps the predecessor's tail).
 * Process-wide switch, off by default; returns the previous setting.  Turn it on around work that is alone on the device
 * (a captured inference step); leave it off when NCCL kernels run concurrently on another stream (DDP fine-tune). */
int pangu_set_pdl(int on);

/* ------------------------------------------------------------------ index kernels (bit-exact) */

/* win[l,t,k,:] = x[src(l,t,k),:] or 0 -- F.pad + torch.roll(-1,-3,-6) + window partition,
 * models/layers.py:224-262.  x [Z*H*W, C] -> win [nLon, T, 144, C]. elem_bytes in {2,4}. */
int pangu_window_partition(const void* x, void* win, const pangu_geom* g, int roll,
                           int elem_bytes, void* stream);
/* x[src(l,t,k),:] = win[l,t,k,:] for real tokens -- window reverse + roll back + crop,
 * models/layers.py:269-293. */
int pangu_window_reverse(const void* win, void* x, const pangu_geom* g, int roll,
                         int elem_bytes, void* stream);
/* idx[l,t,k] = source token of window element or -1 (int64 [nLon,T,144]); the map the two
 * kernels above apply, exported for the bit-exact tests. */
int pangu_window_source_index(int64_t* idx, const pangu_geom* g, int roll, void* stream);
/* mask[t,i,j] = region(i) != region(j) ? -100 : 0, float32 [T,144,144] -- EarthSpecificBlock.gen_mask,
 * models/layers.py:187-216 (identical for every longitude window). */
int pangu_shift_mask(float* mask, const pangu_geom* g, void* stream);
/* EarthAttention3D._construct_index, models/layers.py:371-411: int64 [144*144]. */
int pangu_position_index(int64_t* idx, void* stream);
/* Compact Earth-specific bias (the paper's parameterisation, commented out in the reference at
 * models/layers.py:355,442-449): full[t, h, i, j] = table[position_index[i*144+j], t, h], table fp32 [3312, T, heads],
 * full fp32 [T, heads, 144, 144] (the reference's parameter without its leading 1).  Bit-exact gather. */
int pangu_bias_table_expand(const float* table, float* full, int32_t T, int32_t heads, void* stream);
/* Its adjoint: d_table[idx, t, h] += sum of d_full[t, h, i, j] over the pairs (i, j) with position_index = idx (the gradient of
 * a compact table from the dense bias gradient the attention backward produces); deterministic, no atomics. */
int pangu_bias_table_reduce(const float* d_full, float* d_table, int32_t T, int32_t heads, void* stream);

/* ------------------------------------------------------------------ dense linears */

/* out[M,N] = act(A[M,K] . W[N,K]^T + bias[N])      nn.Linear / Conv1d(k=1):
 * models/layers.py:88,113 (embed), :312,315 (Mlp), :419,481 (attention), :522 (down),
 * :542,566 (up), :591,608 (recover).
 * dtype = PANGU_F32 : A, W, out fp32 (FFMA).  dtype = PANGU_BF16: A, W bf16, fp32 accumulate in
 * TMEM (tcgen05.mma), out bf16 or fp32 per out_dtype.  bias fp32 or NULL.  lda/ldo in elements. */
int pangu_linear(const void* A, int64_t lda, const void* W, const float* bias, void* out,
                 int64_t ldo, int64_t M, int32_t K, int32_t N, int act, int dtype,
                 int out_dtype, void* stream);

/* bf16 tensor-core linear with two fusions around it (both optional):
 *   - A2 != NULL: the A operand is the channel concat cat(A[M,K1], A2[M,K-K1]) read from the two tensors -- the skip
 *     concat in front of the output layer (models/pangu_model.py:98, models/layers.py:591,608); K1 % 64 == 0;
 *   - out_bf16_shadow != NULL (out_dtype fp32): a bf16 copy of the output is written as well (operand of the next
 *     tensor-core GEMM), same row pitch ldo -- replaces the cast pass after embed / down-sample / up-sample. */
int pangu_linear_bf16_ex(const void* A, int64_t lda, const void* A2, int64_t lda2, int32_t K1, const void* W,
                         const float* bias, void* out, void* out_bf16_shadow, int64_t ldo, int64_t M,
                         int32_t K, int32_t N, int act, int out_dtype, void* stream);

/* x_out = residual + LayerNorm_C(y) * gamma + beta   (post-norm residual, models/layers.py:296-297;
 * eps 1e-5).  y is fp32 or bf16 (y_dtype); residual/x_out fp32; x_out_bf16 optional shadow copy
 * (operand of the next GEMM) or NULL.  residual may be NULL (plain LayerNorm). */
int pangu_ln_residual(const void* y, int y_dtype, const float* gamma, const float* beta,
                      const float* residual, float* x_out, void* x_out_bf16, int64_t M,
                      int32_t C, float eps, void* stream);

/* Fused bf16 linear + bias + LayerNorm + residual (N == C in {192,384}):
 *   x_out = residual + LN(A . W^T + bias) * gamma + beta, plus bf16 shadow.
 * attention.linear2 + norm1 + shortcut (models/layers.py:481,296) and Mlp.linear2 + norm2 +
 * residual (:315,297). */
int pangu_linear_ln_residual_bf16(const void* A, int64_t lda, const void* W, const float* bias,
                                  const float* gamma, const float* beta, const float* residual,
                                  float* x_out, void* x_out_bf16, int64_t M, int32_t K, int32_t C,
                                  float eps, void* stream);

/* Fused Mlp + norm2 + residual, bf16 operands (models/layers.py:311-317 and :297):
 *   x_out = residual + LN( GELU(x . W1^T + b1) . W2^T + b2 ) * gamma + beta,  plus bf16 shadow.
 * x [M,C] bf16, W1 [4C,C] bf16, W2 [C,4C] bf16, biases/affine fp32, residual/x_out fp32 [M,C]; C in {192,384}.
 * One kernel: the 4C-wide hidden activation stays in TMEM (tcgen05 cta_group::2, CTA pairs). */
int pangu_mlp_ln_residual_bf16(const void* x, const void* w1, const float* b1, const void* w2,
                               const float* b2, const float* gamma, const float* beta,
                               const float* residual, float* x_out, void* x_out_bf16, int64_t M,
                               int32_t C, float eps, void* stream);

/* The whole tail of an EarthSpecificBlock after the window attention in ONE kernel (models/layers.py:481,296-297):
 *   x1 = x_in + LayerNorm(o . w_proj^T + b_proj) * gamma1 + beta1;   x_out = x1 + LayerNorm(Mlp(x1)) * gamma2 + beta2
 * i.e. pangu_linear_ln_residual_bf16 followed by pangu_mlp_ln_residual_bf16, bit-identical to that pair, but x1 and its bf16
 * shadow never reach HBM: the fp32 x1 tile of a CTA lives in `scratch` (fp32 [scratch_rows, C], scratch_rows >= 128 * the
 * number of SMs; it stays L2-resident), the bf16 x1 tile is written straight into the Mlp's operand tile in shared memory.
 * o bf16 [M, C] (attention output), w_proj bf16 [C, C], w1 bf16 [4C, C], w2 fp16 [C, 4C]; C = 384. */
int pangu_attn_proj_mlp_bf16(const void* o, const void* w_proj, const float* b_proj, const float* gamma1,
                             const float* beta1, const float* x_in, const void* w1, const float* b1, const void* w2,
                             const float* b2, const float* gamma2, const float* beta2, float* scratch,
                             int64_t scratch_rows, float* x_out, void* x_out_bf16, int64_t M, int32_t C, float eps1,
                             float eps2, void* stream);

/* Bring-up aid: copies the fused-Mlp kernel's pipeline timeline (clock64 stamps recorded by CTA 0 when
 * $PANGU_MLP_DBG has bit 16 set) to HOST memory `out` (n <= 512 int64).  Synchronises the device. */
int pangu_debug_mlp_trace(int64_t* out, int32_t n);
/* Same for the tcgen05 window-attention kernel (CTA (0,0,0), $PANGU_ATTN_DBG != 0): [8 windows][16 stamps]. */
int pangu_debug_attn_trace(int64_t* out, int32_t n);

/* ------------------------------------------------------------------ 3-D window attention */

/* EarthAttention3D.forward between linear1 and linear2 (models/layers.py:422-478) with the block's
 * pad/roll/partition/mask/reverse/crop folded into the addressing (models/layers.py:224-293):
 *   qkv  [Z*H*W, 3C] : linear1 output in TOKEN order (channel = s*C + head*32 + d, :422-427);
 *   qkv_bias [3C]    : linear1.bias -- the value of q/k/v on zero pad rows (:228-229,419);
 *   earth_bias [T, heads, 144, 144] (fp32 or bf16 per bias_dtype), added to the scores (:450-453);
 *   roll == 1        : shifted block: source (z+1,h+3,w+6) mod (Z,H+5,W), -100 shift mask (:237,457-464);
 *   roll == 2        : qkv/out are already in window order [nLon*T*144, .] (identity map, no mask) --
 *                      the stand-alone EarthAttention3D.forward(x_window, mask) call;
 *   out  [Z*H*W, C]  : softmax(q*scale k^T + bias + mask) v, heads merged (:476-478), written at the
 *                      un-rolled token position; pad rows dropped.
 * dtype selects fp32 (SIMT) or bf16 (tensor cores) for qkv/out. */
int pangu_window_attention(const void* qkv, const float* qkv_bias, const void* earth_bias,
                           int bias_dtype, void* out, const pangu_geom* g, int roll, int dtype,
                           void* stream);

/* Band-sharded variant (bf16 only): qkv/out hold the band's own rows [Z*hrows*W, .]; halo_qkv [Z*halo*W, .] the
 * first rows of the southern neighbour and halo_lo_qkv [Z*halo_lo*W, .] the last rows of the northern neighbour
 * (read by the windows that straddle a band edge in a rolled block).  Attention output of own rows goes to
 * `out`; output of southern-halo rows goes to halo_out when it is not NULL (to be returned to the neighbour)
 * and is dropped otherwise (both neighbours compute the straddling window).  g is the GLOBAL geometry;
 * bias/mask types are global.  roll in {0,1}.
 * prescaled != 0: the caller folded scale*log2(e) = 32^-0.5 * 1.442695 into the q rows of linear1 (weights AND
 * the qkv_bias given here) and log2(e) into earth_bias, i.e. q k^T + bias is already the exponent in log2 units
 * (the host mirror does this once per weight update); 0 = plain reference semantics (models/layers.py:431-453).
 * Bit 1 of `prescaled` (PANGU_ATTN_EXACT_MAX = 2, pre-scaled bf16 path): take the exact row maximum of S + bias in every
 * window type instead of the bound max(S) + max(bias row) -- for bias tables whose rows spread over more than ~100 log2
 * units (the bound would underflow every exponent); shift-masked window types always take the exact maximum. */
#define PANGU_ATTN_EXACT_MAX 2
/* Bit 2 of `prescaled` (pre-scaled bf16 path, halo_out == NULL): the halo buffers hold only the K and V columns of the
 * neighbours' rows, [Z*halo*W, 2C] -- the neighbours' queries are never needed when each rank keeps only its own output
 * rows (the "redundant" exchange scheme of pangu_b200/dist.py), so a third of the halo bytes does not travel. */
#define PANGU_ATTN_HALO_KV 4
int pangu_window_attention_band(const void* qkv, const void* halo_qkv, const void* halo_lo_qkv,
                                const float* qkv_bias, const void* earth_bias, int bias_dtype, void* out,
                                void* halo_out, const pangu_geom* g, const pangu_band* band, int roll,
                                int prescaled, void* stream);

/* ------------------------------------------------------------------ layout / bandwidth kernels */

/* PatchEmbedding_pretrain.forward up to the two convs (models/layers.py:56-112): normalise
 * (surface :65; upper air with level-flipped stats :95-99), concat constants (:75,:101), zero-pad
 * (:37,:49), patchify.  Writes the GEMM operands
 *   patches_surface [181*360, 112]  feature = c*16 + ph*4 + pw
 *   patches_upper   [7*181*360, 192] feature = c*32 + pz*16 + ph*4 + pw
 * in out_dtype.  input [5,13,721,1440], input_surface [4,721,1440], maps [3,724,1440],
 * const_h [13,721,1440], surface_mean/std [4], upper_mean/std [13,5] (as stored: index 12-level). */
int pangu_patch_embed_gather(const float* input, const float* input_surface,
                             const float* surface_mean, const float* surface_std,
                             const float* upper_mean, const float* upper_std, const float* maps,
                             const float* const_h, void* patches_surface, void* patches_upper,
                             int out_dtype, void* stream);

/* Same for a latitude band: the arrays hold lat_rows valid pixel rows (input [5,13,lat_rows,1440], ...),
 * maps holds map_rows rows, and tok_rows = ceil(lat_rows/4) token rows are produced (rows beyond lat_rows
 * are the zero padding of models/layers.py:37,49).  Full grid: 721, 181, 724. */
int pangu_patch_embed_gather_rows(const float* input, const float* input_surface,
                                  const float* surface_mean, const float* surface_std,
                                  const float* upper_mean, const float* upper_std, const float* maps,
                                  const float* const_h, void* patches_surface, void* patches_upper,
                                  int out_dtype, int32_t lat_rows, int32_t tok_rows, int32_t map_rows,
                                  void* stream);

/* PatchRecovery_pretrain.forward after the two convs (models/layers.py:593-619): un-patchify + crop.
 *   y_upper [7*181*360, 160] (ch = v*32+pz*16+ph*4+pw), y_surface [181*360, 64] (ch = v*16+ph*4+pw), fp32
 *   -> output [5,13,721,1440], output_surface [4,721,1440] fp32. */
int pangu_patch_recover_scatter(const float* y_upper, const float* y_surface, float* output,
                                float* output_surface, void* stream);

/* Same for a latitude band of tok_rows token rows / lat_rows output pixel rows. */
int pangu_patch_recover_scatter_rows(const float* y_upper, const float* y_surface, float* output,
                                     float* output_surface, int32_t lat_rows, int32_t tok_rows,
                                     void* stream);

/* Same with the de-normalisation of era5_data/utils_data.py:540-546 (normBackData: x * std + mean) folded into
 * the scatter -- the step between two chained forecasts (inference/inference_mix_multiOutput.py:238).
 * upper_std/mean [5*13] indexed v*13 + level in DATA level order (weatherStatistics_output,
 * era5_data/utils_data.py:395-421), surface_std/mean [4]; all four NULL = normalised output. */
int pangu_patch_recover_scatter_denorm(const float* y_upper, const float* y_surface, float* output,
                                       float* output_surface, int32_t lat_rows, int32_t tok_rows,
                                       const float* upper_std, const float* upper_mean,
                                       const float* surface_std, const float* surface_mean, void* stream);

/* DownSample.forward before the linear (models/layers.py:501-519): pad H to even, 2x2 merge
 * (feature = dh*2C + dw*C + c), LayerNorm(4C).  x fp32 [Z*H*W, C] -> out [Z*ceil(H/2)*(W/2), 4C]. */
int pangu_downsample_merge_ln(const float* x, const float* gamma, const float* beta, void* out,
                              int out_dtype, int32_t Z, int32_t H, int32_t W, int32_t C, float eps,
                              void* stream);

/* UpSample.forward between linear1 and linear2 (models/layers.py:546-563): pixel-shuffle
 * (in-feature = dh*2C' + dw*C' + c), crop to H rows, LayerNorm(C').
 * y [Z*H2*W2, 4C'] (y_dtype) -> out [Z*H*(2*W2), C'] (out_dtype). */
int pangu_upsample_shuffle_ln(const void* y, int y_dtype, const float* gamma, const float* beta,
                              void* out, int out_dtype, int32_t Z, int32_t H2, int32_t W2, int32_t H,
                              int32_t Cout, float eps, void* stream);

/* fp32 -> bf16 cast of n elements (operand staging for the bf16 path). */
int pangu_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream);
/* out[n, 0:C1] = a[n,:], out[n, C1:C1+C2] = b[n,:] as bf16 (skip concat, models/pangu_model.py:98). */
int pangu_concat_cast_bf16(const float* a, const float* b, void* out, int64_t n, int32_t C1,
                           int32_t C2, void* stream);

/* ------------------------------------------------------------------ fine-tune backward (autograd of the above)
 * The reference obtains these by torch.autograd over models/layers.py (loss.backward(), models/pangu_sample.py:226,
 * under DDP, finetune/finetune_fully.py:220); each entry cites the forward lines it differentiates.  bf16 operands,
 * fp32 accumulation; parameter gradients are fp32 and ACCUMULATED (+=) into caller-zeroed buffers. */

/* out = A . W^T + bias + addend (fp32 out; addend fp32 [M,N] with the pitch of out, or NULL): a dgrad GEMM
 * (dX = dY . W, with W given transposed) fused with the residual-gradient add of models/layers.py:296-297. */
int pangu_linear_bf16_add(const void* A, int64_t lda, const void* W, const float* bias, const float* addend,
                          float* out, int64_t ldo, int64_t M, int32_t K, int32_t N, void* stream);

/* bf16 linear with a second bf16 tensor `aux` [M, N] (row pitch ldo) attached to its epilogue -- the two Mlp fusions of
 * the fine-tune step (models/layers.py:311-317 and its autograd):
 *   PANGU_AUX_PRE_OUT : aux <- A.W^T + bias, out <- act(A.W^T + bias): Mlp.linear1 leaves the pre-activation the backward
 *                       needs and the GELU output in ONE pass;
 *   PANGU_AUX_GELU_BWD: out <- (A.W^T + bias) * GELU'(aux), colsum[n] += sum_m out[m, n] (fp32, or NULL): the dgrad of
 *                       Mlp.linear2 fused with the GELU backward and linear1's bias gradient.
 * Shapes of the CTA-pair GEMM only (K in {192, 384}, N % 192 == 0, M >= 2048); PANGU_ERR_UNSUPPORTED otherwise (the
 * caller then runs the separate kernels). */
typedef enum { PANGU_AUX_PRE_OUT = 1, PANGU_AUX_GELU_BWD = 2 } pangu_aux_mode;
int pangu_linear_bf16_aux(const void* A, int64_t lda, const void* W, const float* bias, void* out, void* aux,
                          int64_t ldo, int64_t M, int32_t K, int32_t N, int act, int aux_mode, float* colsum,
                          void* stream);

/* dW[n_out, k_in] += dY[M, n_out]^T . X[M, k_in] -- weight gradient of nn.Linear / Conv1d(k=1)
 * (models/layers.py:88,113,312,315,419,481,522,542,566,591,608).  dY, X bf16 row-major [tokens, channels] exactly as
 * the passes leave them (read MN-major by tcgen05.mma; no transposes), dW fp32 with row pitch ldw.  Split over
 * token ranges, reduced with red.global.add.v4.f32. */
int pangu_linear_wgrad_bf16(const void* dy, int64_t ldy, const void* x, int64_t ldx, float* dw, int64_t ldw,
                            int64_t M, int32_t n_out, int32_t k_in, void* stream);

/* out[c] += sum_m x[m, c] -- bias gradients.  dtype of x: PANGU_F32 / PANGU_BF16; ld in elements. */
int pangu_colsum(const void* x, int dtype, int64_t ld, int64_t M, int32_t C, float* out, void* stream);

/* Backward of out = scale * (LayerNorm_C(y) * gamma + beta) (models/layers.py:296-297 with the DropPath factor):
 *   dy (bf16) = d out / d y applied to dout (+ dout2 when not NULL: the two gradient streams that meet at a residual),
 *   dgamma += scale * sum dout * yhat, dbeta += scale * sum dout, dcolsum += sum dy (bias gradient of the linear
 *   that produced y); any of the three may be NULL.  y fp32 or bf16 [M, C], C in {192, 384}. */
int pangu_ln_backward(const float* dout, const float* dout2, const void* y, int y_dtype, const float* gamma,
                      float scale, void* dy, float* dgamma, float* dbeta, float* dcolsum, int64_t M, int32_t C,
                      float eps, void* stream);

/* Backward of pangu_upsample_shuffle_ln (models/layers.py:546-563): dout fp32 [Z*H*2*W2, Cout], y bf16 [Z*H2*W2, 4*Cout]
 * -> dy bf16 (same layout as y; the caller zero-fills it: the cropped row gets no gradient), dgamma/dbeta +=. */
int pangu_upsample_shuffle_ln_backward(const float* dout, const void* y, const float* gamma, void* dy, float* dgamma,
                                       float* dbeta, int32_t Z, int32_t H2, int32_t W2, int32_t H, int32_t Cout,
                                       float eps, void* stream);

/* Backward of pangu_downsample_merge_ln (models/layers.py:501-519): dout fp32 [Z*ceil(H/2)*(W/2), 4C], x fp32
 * [Z*H*W, C] -> dx fp32 (every real token is written exactly once; the pad row is dropped), dgamma/dbeta +=. */
int pangu_downsample_merge_ln_backward(const float* dout, const float* x, const float* gamma, float* dx,
                                       float* dgamma, float* dbeta, int32_t Z, int32_t H, int32_t W, int32_t C,
                                       float eps, void* stream);

/* h = GELU(h_pre), exact erf form, bf16, n % 8 == 0 (recompute of models/layers.py:313 for the backward). */
int pangu_gelu_bf16(const void* h_pre, void* h, int64_t n, void* stream);
/* dh_pre = dh * GELU'(h_pre) (bf16 [M, F]; dh_pre may alias dh), dcolsum[F] += sum_m dh_pre (or NULL). */
int pangu_gelu_backward_bf16(const void* dh, const void* h_pre, void* dh_pre, float* dcolsum, int64_t M, int32_t F,
                             void* stream);

/* pangu_window_attention_band on the whole grid with pre-scaled operands (see there), additionally writing
 * lse [nLon, T, heads, 144] fp32 = log2-sum-exp of every score row, which the backward kernel consumes.
 * roll | PANGU_ROLL_EXACT_MAX: exact row maximum everywhere (see PANGU_ATTN_EXACT_MAX). */
#define PANGU_ROLL_EXACT_MAX 0x100
int pangu_window_attention_train(const void* qkv, const float* qkv_bias, const void* earth_bias, void* out,
                                 float* lse, const pangu_geom* g, int roll, void* stream);

/* Backward of the window attention (models/layers.py:431-478 inside :224-293).  qkv / qkv_bias / earth_bias (bf16)
 * pre-scaled exactly as given to pangu_window_attention_train, out its output, d_out bf16 [N, C] the incoming
 * gradient, lse from the forward.  Writes d_qkv bf16 [N, 3C] (w.r.t. the UN-scaled linear1 output, token order);
 * accumulates d_earth_bias fp32 [T, heads, 144, 144] (sum over longitude windows of dS; gradient of the fp32
 * earth_specific_bias parameter) and d_qkv_bias fp32 [3C] = linear1's bias gradient: the column sums of dq/dk/dv over
 * ALL window rows, including the zero-pad rows, which equal the bias in the forward (models/layers.py:228,419).
 * roll as in pangu_window_attention. */
int pangu_window_attention_backward(const void* qkv, const float* qkv_bias, const void* earth_bias, const void* out,
                                    const void* d_out, const float* lse, void* d_qkv, float* d_earth_bias,
                                    float* d_qkv_bias, const pangu_geom* g, int roll, void* stream);

/* Inverse of pangu_patch_recover_scatter_rows (models/layers.py:593-619): gradients of the output fields ->
 * gradients of the two conv outputs as bf16 [7*tok_rows*360, 160] / [tok_rows*360, 64]; cropped positions get 0. */
int pangu_patch_recover_gather_backward(const float* d_output, const float* d_output_surface, void* dy_upper,
                                        void* dy_surface, int32_t lat_rows, int32_t tok_rows, void* stream);

/* The reference's training loss and its gradient in ONE pass (models/pangu_sample.py:163-218, default branch; SURVEY 8f
 * rank 2): target normalised per plane ((x - mean[plane]) / std[plane], era5_data/utils_data.py normData; mean = std =
 * NULL: already normalised), loss_sum += scale * sum w[plane / planes_per_var] * |out - target|, and, when d_out is
 * not NULL, d_out = scale * w * sign(out - target).  out / target / d_out fp32 [planes, plane_elems]; the caller
 * passes scale = loss_weight / numel so that loss_sum is loss_weight * mean(L1 * w). */
int pangu_weighted_l1_loss(const float* out, const float* target, const float* mean, const float* stdv,
                           const float* weight, int32_t planes, int32_t planes_per_var, int64_t plane_elems,
                           float scale, float* loss_sum, float* d_out, void* stream);

/* The same with the reference's custom mask (models/pangu_sample.py:120-127,194-199: `use_custom_mask`): every L1 term
 * is multiplied by mask[h][w] (fp32 [plane_elems], the same for all planes); the caller passes
 * scale = loss_weight / valid_points (train(), :198-199) or loss_weight / (valid_points * channels) (test(), :467). */
int pangu_weighted_l1_loss_masked(const float* out, const float* target, const float* mean, const float* stdv,
                                  const float* weight, const float* mask, int32_t planes, int32_t planes_per_var,
                                  int64_t plane_elems, float scale, float* loss_sum, float* d_out, void* stream);

/* The reference's wind-speed loss (models/pangu_sample.py:74-93 get_wind_speed, :184-193 `only_use_wind_speed_loss`) and
 * its gradient in one pass: ws = sqrt(u^2 + v^2) of the output planes and of the (normalised on the fly, mean/std per
 * plane or NULL) target planes, loss_sum += scale * sum mask * |ws(out) - ws(target)|; d_u / d_v (NULL: no gradient)
 * = scale * mask * sign(.) * u / ws(out) (0 where ws(out) = 0).  u / v [planes][plane_elems] fp32, plane p of u and of
 * v being the same level; mask [plane_elems] or NULL. */
int pangu_wind_speed_l1_loss(const float* out_u, const float* out_v, const float* tgt_u, const float* tgt_v,
                             const float* mean_u, const float* std_u, const float* mean_v, const float* std_v,
                             const float* mask, int32_t planes, int64_t plane_elems, float scale, float* loss_sum,
                             float* d_u, float* d_v, void* stream);

/* Latitude-weighted verification scores in one pass (era5_data/score.py:126-161
 * weighted_rmse_torch_channels with its optional mask, :181-201 weighted_acc_torch_channels).  pred / target fp32
 * [planes][H][W]; mask [H][W] or NULL (= 1); clim [planes] or NULL (= 0): the climatology subtracted from both fields for the
 * anomaly correlation (models/pangu_sample.py:549-556); lat_weight [H] = the reference's latitude_weighting_factor_torch.
 * sums [planes][5] fp64 (zeroed by the call) = { sum w m (p-t)^2, sum w m, sum w a b, sum w a^2, sum w b^2 } with a = p - clim,
 * b = t - clim; RMSE = sqrt(sums[0] / (H W)) without a mask, sqrt(sums[0] / sums[1]) with one; ACC = sums[2] / sqrt(sums[3] sums[4]). */
int pangu_lat_weighted_score_sums(const float* pred, const float* target, const float* mask, const float* clim,
                                  const float* lat_weight, int32_t planes, int32_t H, int32_t W, double* sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PANGU_B200_H_ */
