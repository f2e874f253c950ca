/* libcsts_b200.so — C ABI of the B200 (sm_100a) kernels behind the CSTS hot path.
 *
 * The reference (BolinLai/CSTS) is pure Python: the "FFI" of this path is the set of
 * torch.nn / torch.nn.functional calls made by slowfast/models/{attention,av_attention,common,
 * stem_helper,custom_multimodal_builder,losses}.py and slowfast/utils/utils.py.  Each entry point
 * below names the reference call site(s) it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C: raw device pointers, sizes and strides in ELEMENTS, `void* stream` = cudaStream_t.
 *   - every function returns 0 on success and a non-zero code otherwise; csts_last_error() returns the
 *     calling thread's message.  No exception crosses the ABI, there is no fallback path.
 *   - asynchronous on `stream`, no internal synchronisation, no device allocation, no retained pointers:
 *     the caller (PyTorch's caching allocator) owns all memory.  Re-entrant / thread-safe (backward runs
 *     on autograd's worker thread).
 *   - dtype codes: 0 = f32, 1 = bf16, 2 = f16 (IEEE half).  The residual stream, statistics and parameter
 *     gradients are f32.  Every entry point takes the type of its 16-bit tensors explicitly.  The host layer
 *     runs one type per precision mode: bf16 by default (fp32 range, no loss scaling), f16 under
 *     TRAIN.MIXED_PRECISION (the reference's fp16 autocast + GradScaler contract, tools/train_avgaze_net.py:70,
 *     99-109; 10 mantissa bits).  The two operands of a tcgen05 GEMM must share one type (a kind::f16 MMA with
 *     different A / B formats faults on sm_100a); the mma.sync kernel accepts mixed pairs.
 *   - the library is built for sm_100a only; csts_check_device() reports anything else.
 */
#ifndef CSTS_B200_H
#define CSTS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------------- */
int csts_version(void);
int csts_last_error(char* buf, int len);
int csts_check_device(void);
long long csts_launch_count(int reset);          /* kernels launched since load / last reset */

/* ---- GEMM -------------------------------------------------------------------------------------------
 * C[z][m][n] = epi( alpha * sum_k opA(A[z])[m][k] * opB(B[z])[n][k] )
 * Replaces every nn.Linear (attention.py:130,159; common.py:19-31; attention.py:246; custom_multimodal_
 * builder.py:495-496), the q@k^T / attn@v batched matmuls (attention.py:154,158; av_attention.py:137,141,
 * 334,351), the PatchEmbed Conv3d after im2col (stem_helper.py:27-36), the dense (1,8,8) Conv3d pools
 * (custom_multimodal_builder.py:227-229) and all of their autograd backward products.
 * Large token-major problems run on the tcgen05/TMEM/TMA kernel, odd shapes on the mma.sync kernel. */
typedef struct csts_gemm_args {
  const void* A;          /* bf16 or f16 (a_dtype) */
  const void* B;          /* bf16 or f16 (b_dtype) */
  void* C;                /* f32, bf16 or f16 (c_dtype) */
  void* Z;                /* bf16 or f16 (z_dtype), same shape as C (pitch ldz): act==1 -> receives GELU'(pre-activation),
                                                                act==2 -> is read (C = result * Z),
                                                                act==4 -> the softmax probabilities P */
  const float* bias;      /* [N] or NULL */
  const float* residual;  /* f32 [rows, N] (pitch ldr) or NULL; row = m % res_mod when res_mod > 0 */
  const float* row_scale; /* [ceil(M / rows_per_scale)] or NULL: row m is multiplied by
                             row_scale[m / rows_per_scale] before the residual is added (DropPath, common.py:46-59) */
  float* rowsum;          /* [M] f32 or NULL: rowsum[m] += sum_k opA(A)[m][k].  For a weight-gradient product
                             dW = dY^T . X (a_kmajor = 0) this is the bias gradient sum_tokens dY, produced by one extra
                             N = 16 MMA per k-step against an all-ones tile instead of a separate pass over dY.
                             Needs a_kmajor = 0 and a single batch; accumulated into (zeroed by the caller) */
  int64_t lda, ldb, ldc, ldz, ldr;
  int64_t sA1, sA2, sB1, sB2, sC1, sC2;   /* batch strides: z -> (z / batch2, z % batch2) */
  int32_t M, N, K;
  int32_t batch1, batch2;
  int32_t a_kmajor;       /* 1: A[m*lda + k]   0: A[k*lda + m] */
  int32_t b_kmajor;       /* 1: B[n*ldb + k]   0: B[k*ldb + n] */
  int32_t c_dtype;        /* 0 f32, 1 bf16, 2 f16 */
  int32_t act;            /* 0 none, 1 erf GELU (nn.GELU(), common.py:21; Z <- GELU'), 2 times Z (GELU backward),
                             3 row softmax of alpha*acc (N <= 256, 16-bit C = P), 4 softmax backward: C = alpha*Z o (acc - rowsum(acc o Z)),
                             two-pass attention for N > 256 (neither the f32 scores nor dP reach memory):
                             5 C[z][m] = logsumexp_n(alpha*acc) (f32, one value per row), 6 C = exp(alpha*acc - rowvec[z][m]) (16-bit P),
                             7 C = alpha * Z o (acc - rowvec[z][m]) (16-bit dS; Z = P with C's layout, rowvec = rowsum(dO o O)) */
  int32_t accumulate;     /* C += result */
  int32_t res_mod;
  int32_t split_k;        /* > 1: partial sums combined with f32 atomics (C must be f32); < 0: the library picks the factor */
  float alpha;
  int32_t backend;        /* 0 auto, 1 mma.sync, 2 tcgen05 */
  int32_t rows_per_scale;
  int32_t a_dtype;        /* 1 bf16, 2 f16 */
  int32_t b_dtype;        /* 1 bf16, 2 f16 */
  int32_t z_dtype;        /* 1 bf16, 2 f16 (ignored when Z is NULL) */
  int32_t tile_n;         /* tuning override of the tcgen05 tile width (96 / 128 / 192 / 256); 0: the library picks */
  int32_t ctas;           /* tuning override of the kernel build: 1 = one CTA per SM (12 epilogue warps, 512 TMEM columns),
                             2 = two CTAs per SM (4 epilogue warps, 256 TMEM columns each); 0: the library picks */
  const float* rowvec;    /* act 6 / 7: f32 [batch][M] per-row input (row logsumexp; rowsum(dO o O)) */
} csts_gemm_args;
int csts_gemm(const csts_gemm_args* a, void* stream);
int csts_gemm_backend(const csts_gemm_args* a);  /* 2 = tcgen05, 1 = mma.sync for this problem */
int csts_gemm_plan(const csts_gemm_args* a, int* tile_n, int* ctas, int* splits);   /* what the tcgen05 launcher would choose */

/* D[b][head][q] = sum_d dO[b][q][head][d] * O[b][q][head][d] for two 16-bit (B, Lq, heads*d) tensors: the row term of the
 * softmax backward (rowsum(dP o P) = dO . O), computed without dP (attention.py:154-158 backward) */
int csts_rowdot(const void* dO, const void* O, int dtype, float* D, int B, int Lq, int heads, int d, void* stream);

/* ---- LayerNorm: nn.LayerNorm (attention.py:239,243 norm1/norm2 eps 1e-6; :42-43 norm_q/k/v eps 1e-5) --- */
int csts_layernorm_fwd(const void* x, int x_dtype, void* y, int y_dtype, const float* gamma, const float* beta, float* mean,
                       float* rstd, int64_t rows, int width, float eps, void* stream);
/* dx = [add +] LN'(dy); dgamma/dbeta (f32, zeroed by the caller) are accumulated into.  dx16 (optional): a second,
 * 16-bit copy of dx with row m multiplied by row_scale[m / rows_per_scale] (the DropPath-scaled operand of the next
 * backward GEMM), written in the same pass */
int csts_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean, const float* rstd,
                       const float* gamma, const float* add, void* dx, int dx_dtype, float* dgamma, float* dbeta, int64_t rows,
                       int width, void* dx16, int dx16_dtype, const float* row_scale, int rows_per_scale, void* stream);

/* two LayerNorm backward problems of identical geometry and types in one launch (norm_k / norm_v of a block, attention.py:135-138) */
int csts_layernorm_bwd_pair(const void* dy0, const void* dy1, int dy_dtype, const void* x0, const void* x1, int x_dtype, const float* mean0,
                            const float* mean1, const float* rstd0, const float* rstd1, const float* gamma0, const float* gamma1, void* dx0,
                            void* dx1, int dx_dtype, float* dgamma0, float* dgamma1, float* dbeta0, float* dbeta1, int64_t rows, int width,
                            void* stream);

/* ---- attention softmax: attn.softmax(dim=-1) (attention.py:155), with the in-frame mask of
 * SpatialAttention (av_attention.py:336-348) when mask_hw > 0.  P is bf16 / f16 (p_dtype), pad columns
 * [n, ldp) zero; dS is bf16 / f16 (ds_dtype). */
int csts_softmax_fwd(const float* S, void* P, int p_dtype, int64_t rows, int n, int lds, int ldp, int nq, int mask_hw, int mask_t,
                     void* stream);
int csts_softmax_bwd(const void* P, int p_dtype, const float* dP, void* dS, int ds_dtype, int64_t rows, int n, int ldp, int lddp,
                     float scale, void* stream);

/* ---- casts / layout / elementwise ------------------------------------------------------------------- */
/* f32 (rows, cols) -> bf16 / f16 (rows, ld_out), zero padded; row m scaled by row_scale[m / rows_per_scale] if given */
int csts_cast16(const float* src, void* dst, int dst_dtype, int64_t rows, int cols, int ld_out, const float* row_scale,
                int rows_per_scale, void* stream);
int csts_permute_021(const float* src, void* dst, int dst_dtype, int a, int b, int c, void* stream); /* [a][b][c] -> [a][c][b] */
int csts_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream);     /* decoder skips, custom_multimodal_builder.py:467-473 */
int csts_scale_f32(const float* a, const float* device_scalar, float* out, int64_t n, void* stream);
int csts_colsum(const void* X, int x_dtype, float* out, int64_t M, int N, int64_t ld, void* stream);  /* bias gradients */

/* ---- attention_pool / attention_upsample: depthwise 3x3x3 Conv3d / ConvTranspose3d over the token grid
 * + LayerNorm(head_dim) (attention.py:11-49, :105-116; :251-292, :344-348).  Element (b, head, pos, c) of
 * `in` lives at in + b*in_sB + head*in_sH + pos*in_sP + c: the kernels read the (B,N,3,heads,d) qkv
 * tensor in place, so the reference's permute+contiguous copies (attention.py:31,37) do not exist. */
typedef struct csts_pool_args {
  const void* in;         /* bf16 or f16 (dtype) */
  void* out;              /* same type as in */
  const float* w;         /* (d,1,3,3,3) parameter, f32 */
  const float* gamma;     /* LayerNorm(d) weight or NULL (no norm: `out` = raw conv) */
  const float* beta;
  void* pre;              /* same type, dense (B, heads, Lo, d): raw conv output kept for backward (norm only) */
  float* mean;            /* [B*heads*Lo] (norm only) */
  float* rstd;
  int64_t in_sB, in_sH, in_sP;
  int64_t out_sB, out_sH, out_sP;
  int32_t B, heads, d;
  int32_t Ti, Hi, Wi;     /* input grid */
  int32_t To, Ho, Wo;     /* output grid */
  int32_t st, sh, sw;     /* stride of the (un-transposed) convolution, powers of two */
  int32_t transposed;     /* 0: out[o] = sum_tap w[tap] in[o*s+tap-1]; 1: out[o] = sum_tap w[tap] in[(o+1-tap)/s] */
  float eps;
  int32_t dtype;          /* 1 bf16, 2 f16: type of in / out / pre */
  /* optional second problem of identical geometry run by the same launch (the k and v pools of a block share
   * everything but their tensors): used when in2 != NULL */
  const void* in2;
  void* out2;
  const float* w2;
  const float* gamma2;
  const float* beta2;
  void* pre2;
  float* mean2;
  float* rstd2;
} csts_pool_args;
int csts_dwconv(const csts_pool_args* p, void* stream);

/* dw[c][tap] += sum small[o][c] * big[o*s + tap - 1][c]   (conv: small = d(out), big = in; transposed: swapped) */
typedef struct csts_wgrad_args {
  const void* small;      /* bf16 or f16 (small_dtype), grid (Ts,Hs,Ws) */
  const void* big;        /* bf16 or f16 (big_dtype), grid (Tb,Hb,Wb) */
  float* dw;              /* (d,1,3,3,3) f32, accumulated into */
  int64_t small_sB, small_sH, small_sP;
  int64_t big_sB, big_sH, big_sP;
  int32_t B, heads, d;
  int32_t Ts, Hs, Ws;
  int32_t Tb, Hb, Wb;
  int32_t st, sh, sw;
  int32_t small_dtype, big_dtype;
  /* optional second problem of identical geometry in the same launch (used when small2 != NULL) */
  const void* small2;
  const void* big2;
  float* dw2;
} csts_wgrad_args;
int csts_dwconv_wgrad(const csts_wgrad_args* p, void* stream);

/* ---- residual-path resampling on the token-major f32 stream (B, T*H*W, C) -------------------------------
 * pool_skip = nn.MaxPool3d((1,3,3),(1,2,2),(0,1,1)) (attention.py:225-236,240);
 * upsample_skip = nn.Upsample(trilinear, align_corners=False) (attention.py:463-467,471). */
int csts_maxpool_fwd(const float* x, float* y, void* argmax_u8, int B, int T, int H, int W, int C, void* stream);
int csts_maxpool_bwd(const float* dy, const void* argmax_u8, float* dx, int B, int T, int H, int W, int C, void* stream);
int csts_upsample_fwd(const float* x, float* y, int B, int T, int H, int W, int C, int ft, int fh, int fw, void* stream);
int csts_upsample_bwd(const float* dy, float* dx, int B, int T, int H, int W, int C, int ft, int fh, int fw, int accumulate, void* stream);

/* ---- stem / fusion glue / head --------------------------------------------------------------------- */
/* PatchEmbed Conv3d k(3,7,7) s(2,4,4) p(1,3,3) as im2col (stem_helper.py:27-38): x f32 (B,Cin,T,H,W) ->
 * bf16 / f16 [B*T/2*H/4*W/4, Kp], column ((c*3+kt)*7+kh)*7+kw, zero padded to Kp */
int csts_im2col_patch(const float* x, void* patches, int dtype, int B, int Cin, int T, int H, int W, int Kp, void* stream);
/* separable position embedding (custom_multimodal_builder.py:362-370) and its gradient */
int csts_pos_embed(const float* spatial, const float* temporal, float* pos, int T, int HW, int C, void* stream);
int csts_pos_embed_bwd(const float* dY, float* dspatial, float* dtemporal, int B, int T, int HW, int C, void* stream);
/* fusion re-weighting x * w[:, t] (custom_multimodal_builder.py:454-461) */
int csts_reweight_fwd(const float* x, const float* w, float* out, int B, int T, int S, int C, int64_t w_sB, void* stream);
int csts_reweight_bwd(const float* dout, const float* x, const float* w, float* dx, float* dw, int B, int T, int S, int C, int64_t w_sB,
                      void* stream);
/* x.mean(dim=1) feeding vision_proj / audio_proj (custom_multimodal_builder.py:493-496) */
int csts_token_mean_fwd(const float* x, void* out, int out_dtype, int B, int N, int C, void* stream);
int csts_token_mean_bwd(const float* dout, float* dx, int B, int N, int C, int accumulate, void* stream);
/* classifier(feat + F.interpolate(stem, T -> 2T, trilinear)) (custom_multimodal_builder.py:476-481) */
int csts_classifier_fwd(const float* feat, const float* stem, const float* w, const float* bias, float* logits, int B, int Ti, int S, int C,
                        void* stream);
int csts_classifier_bwd(const float* dlogits, const float* feat, const float* stem, const float* w, float* dfeat, float* dstem, float* dw,
                        float* dbias, int B, int Ti, int S, int C, void* stream);

/* ---- losses -------------------------------------------------------------------------------------------
 * frame_softmax (slowfast/utils/utils.py:5-12) fused with KLDiv (slowfast/models/losses.py:59-82): emits the
 * soft-maxed heat-maps, the scalar loss and d loss / d logits in one pass over the logits. */
int csts_kldiv_frame_softmax(const float* logits, const float* target, float* prob, float* frame_kl, float* loss, float* dlogits,
                             int frames, int HW, int T, float temperature, void* stream);
/* sim_matrix (slowfast/utils/utils.py:15-24) and its gradient */
int csts_sim_matrix_fwd(const float* a, const float* b, float* sim, float* na, float* nb, int n, int D, float eps, void* stream);
int csts_sim_matrix_bwd(const float* a, const float* b, const float* sim, const float* dsim, const float* na, const float* nb, float* da,
                        float* db, int n, int D, void* stream);
/* EgoNCE (slowfast/models/losses.py:157-170): loss and d loss / d sim; lse_scratch holds 2*n floats */
int csts_egonce(const float* sim, float* loss, float* dsim, float* lse_scratch, int n, float temperature, void* stream);

/* ---- optimizer step ------------------------------------------------------------------------------------
 * The tail of the training step (tools/train_avgaze_net.py:101-109; slowfast/models/optimizer.py:98-104):
 * scaler.unscale_ + clip_grad_norm_(max_norm) + AdamW (torch semantics: decoupled weight decay, eps outside
 * the bias-corrected sqrt) + refresh of the 16-bit operand copy of each weight, as two multi-tensor launches.
 * `tensors` is a DEVICE array describing every parameter; `chunks` a DEVICE array of (tensor index, chunk
 * index) pairs, one per csts_mt_chunk_elems() elements of each tensor.  All state is f32.  When the gradient
 * norm is not finite the step is skipped and *found_inf = 1 (GradScaler semantics).  *step is the number of
 * steps already taken; the caller increments it afterwards (by 1 - *found_inf). */
typedef struct csts_mt_tensor {
  void* param;            /* f32, updated in place */
  const void* grad;       /* f32, read only (left scaled / un-clipped) */
  void* exp_avg;          /* f32 */
  void* exp_avg_sq;       /* f32 */
  void* w16;              /* bf16 / f16 copy of the updated parameter, same element order, or NULL */
  int64_t numel;
  float weight_decay;
  int32_t group;          /* 0 / 1: which learning-rate scalar applies */
  int32_t w16_dtype;      /* 1 bf16, 2 f16 */
  int32_t pad_;
} csts_mt_tensor;
int csts_mt_chunk_elems(void);
int csts_grad_sqnorm(const csts_mt_tensor* tensors, const int32_t* chunks, int n_chunks, double* out_sq, void* stream);
int csts_clip_adamw_step(const csts_mt_tensor* tensors, const int32_t* chunks, int n_chunks, const double* total_sq, const float* grad_scale,
                         float* found_inf, const float* step, const float* lr0, const float* lr1, double beta1, double beta2, float eps,
                         float max_norm, void* stream);

/* ---- metric: adaptive-threshold F1 (slowfast/utils/metrics.py:9-74), with the per-frame min-max rescale of the loops
 * (tools/train_avgaze_net.py:125-127) fused in when rescale != 0.  preds / labels_hm: f32 [frames][HW]; labels: f32
 * [frames][3] (x, y, gaze type); thresholds: f32 [n_thr <= 32], ascending.  counts: scratch f32 [frames * (2*n_thr + 1)];
 * out: f32 [5] = f1, recall, precision, best threshold, its index.  No host synchronisation. */
int csts_adaptive_f1(const float* preds, const float* labels_hm, const float* labels, const float* thresholds, int n_thr, int frames, int HW,
                     int fixation_idx, int rescale, float* counts, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSTS_B200_H */
