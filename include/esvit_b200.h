/* esvit_b200 C ABI — the drop-in boundary of the B200-native EsViT multi-crop self-distillation step.
 *
 * Every entry point takes raw DEVICE pointers + sizes + a cudaStream_t (as void*), launches sm_100a kernels on
 * that stream and returns an int status: 0 = ok, otherwise a cudaError_t value or ESVIT_ERR_BAD_ARG (1001) for an
 * unsupported shape.  No entry point allocates, frees or synchronises; all buffers (inputs, outputs, workspaces)
 * are owned by the caller (PyTorch's caching allocator on the host side) and must stay alive in stream order.
 * Entry points are re-entrant and keep no mutable global state.
 *
 * The reference (microsoft/esvit) is pure Python/PyTorch and has no FFI of its own; each function below names the
 * reference code it replaces (file:line relative to the reference tree).  INTEGRATION.md shows the ctypes binding
 * and the module-level swap a maintainer would add to main_esvit.py.
 *
 * Conventions: "bf16" = __nv_bfloat16 bits; token-major activations [B, H*W, C] (same order as the reference);
 * fp32 residual stream; bf16 GEMM operands.  "ACCUMULATED" outputs must be zero-filled by the caller.
 */
#ifndef ESVIT_B200_H
#define ESVIT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ESVIT_ERR_BAD_ARG 1001

/* ---- residual add + LayerNorm ------------------------------------------------------------------------------
 * replaces: x = shortcut + drop_path(branch); y = norm(x)      models/swin_transformer.py:329-331, :283, :687
 * xout = x + keep[row / tokens_per_sample] * delta (delta/keep/xout may be NULL; delta = proj / fc2 GEMM output
 * including its bias); y = LN(xout) (y may be NULL).  x may be NULL (= 0) when delta is given: xout = fp32(delta), the
 * first LN after PatchMerging's reduction GEMM (:411-415) without a separate bf16 -> fp32 pass.
 * x fp32 [T,C]; delta bf16 [T,C]; keep fp32 [B]; y bf16 or fp32 [T,C]; mean/rstd fp32 [T] (saved for backward). */
int esvit_add_ln_fwd(const float* x, const void* delta, const float* keep, int tokens_per_sample,
                     const float* gamma, const float* beta, float eps, float* xout, void* y, int y_is_bf16,
                     float* mean, float* rstd, long long T, int C, void* stream);
/* dx = dxo + LNbwd(dy); ddelta = keep * dx (bf16); dgamma/dbeta/ddelta_bias ACCUMULATED (ddelta_bias = column sums of
 * ddelta = gradient of the proj / fc2 bias).  dy / dxo / dx / ddelta / ddelta_bias may be NULL. */
int esvit_add_ln_bwd(const void* dy, int dy_is_bf16, const float* dxo, const float* xs, const float* mean,
                     const float* rstd, const float* gamma, const float* keep, int tokens_per_sample, float* dx,
                     void* ddelta, float* dgamma, float* dbeta, float* ddelta_bias, long long T, int C, void* stream);

/* ---- PatchMerging gather + LayerNorm(4C) ---------------------------------- models/swin_transformer.py:393-417
 * x fp32 [B,H,W,C] -> y bf16 [B,ceil(H/2)*ceil(W/2),4C] (the 4C->2C reduction GEMM follows as a library GEMM). */
int esvit_patch_merge_ln_fwd(const float* x, const float* gamma, const float* beta, float eps, void* y, float* mean,
                             float* rstd, int B, int H, int W, int C, void* stream);
int esvit_patch_merge_ln_bwd(const void* dy, const float* x, const float* mean, const float* rstd,
                             const float* gamma, float* dx, float* dgamma, float* dbeta, int B, int H, int W, int C,
                             void* stream);

/* ---- token mean (AdaptiveAvgPool1d(1)) ------------------------------------- models/swin_transformer.py:688-689 */
int esvit_token_mean_fwd(const float* region, float* pooled, int B, int N, int C, void* stream);
int esvit_token_mean_bwd(const float* dpooled, const float* dregion_in, float* dregion, int B, int N, int C,
                         void* stream);

/* ---- PatchEmbed: 4x4/4 conv (3->E) + LayerNorm -------------------------------- models/swin_transformer.py:537-547
 * img fp32 [B,3,H,W]; w fp32 [E,3,4,4]; out fp32 [B,(H/4)*(W/4),E].  bwd ACCUMULATES dw/dbias/dgamma/dbeta. */
int esvit_patch_embed_fwd(const float* img, const float* w, const float* bias, const float* gamma, const float* beta,
                          float eps, float* out, float* mean, float* rstd, int B, int H, int W, int E, void* stream);
int esvit_patch_embed_bwd(const float* img, const float* w, const float* bias, const float* gamma, const float* mean,
                          const float* rstd, const float* dout, float* dw, float* dbias, float* dgamma, float* dbeta,
                          int B, int H, int W, int E, void* stream);

/* ---- (shifted-)window attention core ----------------------------- models/swin_transformer.py:120-152, :283-325
 * Folds pad / roll / window_partition / rel-pos bias / -100 shift mask / softmax / PV / window_reverse / roll / crop.
 * qkv bf16 [B,H,W,3C] ([q|k|v][head][32]) is the qkv GEMM output including its bias; qkv_bias bf16 [3C] is what a
 * padded slot holds (the bias alone); bias_table fp32 [(2ws-1)^2, nH]; out bf16 [B,H,W,C]; lse fp32
 * [B*nWindows, nH, ws*ws].  ws in {7,14}; head_dim 32.
 * bias_ws fp32 [nH*8192]: caller-owned scratch (ws 7: the rel-pos bias expanded to [nH][64][64]; ws 14 backward: the
 * lane-expanded bias-gradient accumulator [nH][27][6][32], cleared and folded into dbias_table inside the call).
 * bias_ready (ws 7): 1 = bias_ws already holds the expansion written by esvit_window_attn_expand_bias for this table
 * (one expansion per table per step instead of one per call), 0 = the call expands it itself.
 * bwd: dqkv fully written; dbias_table fp32 and dqkv_bias fp32 [3C] (complete qkv-bias gradient) ACCUMULATED.
 * ws 7 backward: tcgen05 / TMEM kernel (five tensor-core GEMMs per window pair, dq/dk/dv through 4-D bulk tensor stores:
 * qkv / dqkv / out rows must be 16-byte aligned, which C % 8 == 0 and a 16-byte aligned base guarantee); ws 7 forward
 * and ws 14: mma.sync kernels (ESVIT_ATTN_TC selects). */
int esvit_window_attn_expand_bias(const float* bias_table, float* bias_ws, int nH, int ws, void* stream);
int esvit_window_attn_fwd(const void* qkv, const void* qkv_bias, const float* bias_table, float* bias_ws, int bias_ready,
                          void* out, float* lse, int B, int H, int W, int C, int nH, int ws, int shift, float scale,
                          void* stream);
int esvit_window_attn_bwd(const void* qkv, const void* qkv_bias, const float* bias_table, float* bias_ws, int bias_ready,
                          const void* out, const void* dout, const float* lse, void* dqkv, float* dbias_table,
                          float* dqkv_bias, int B, int H, int W, int C, int nH, int ws, int shift, float scale,
                          void* stream);

/* ---- tcgen05 / TMA GEMM with fused epilogue ------------------------- nn.Linear + Mlp.act, models/swin_transformer.py:31-33
 * out[M,N] (bf16) = act(a[M,K] @ w[N,K]^T + bias[N]); act 0 = identity, 1 = exact GELU (then `pre`, if not NULL, gets
 * gelu'(pre-activation): the backward is dh = dy * pre, see esvit_mul_bwd_dbias).  a, w bf16 row-major (K contiguous), K % 8 == 0, N % 8 == 0, bias fp32 or NULL.
 * TMA-staged 128B-swizzled tiles, tcgen05.mma with the fp32 accumulator in TMEM, persistent over output tiles. */
int esvit_gemm_bias_act(const void* a, const void* w, const float* bias, void* out, void* pre, long long M, int N, int K,
                        int act, void* stream);
/* out[M,N] (bf16) = (a[M,K] @ w[N,K]^T) * mult[M,N]; colsum[N] (fp32) += column sums of out (caller zero-fills).
 * Same kernel, third epilogue: the fc2 input-gradient GEMM (a = dy, w = W2^T) fused with the GELU backward of fc1
 * (mult = gelu'(pre-activation) from esvit_gemm_bias_act act = 1) and the fc1 bias gradient - what autograd runs as
 * mm + GeluBackward + sum(0) for Mlp.forward, models/swin_transformer.py:31-35.  The multiplier tile is TMA-loaded
 * into the store-staging slot, so it is read once, coalesced, and never touches registers before use.
 * ws fp32 [160*N]: caller-owned scratch (per-CTA partial column sums, folded into colsum inside the call). */
int esvit_gemm_mul_colsum(const void* a, const void* w, const void* mult, void* out, float* colsum, float* ws,
                          long long M, int N, int K, void* stream);

/* ---- second-generation tcgen05 GEMM family (csrc/gemm2_tcgen05.cu): every nn.Linear of the step, forward, input
 * gradient and weight gradient -------- models/swin_transformer.py:21-37,88-91,125,150,393-420; vision_transformer.py:385-418
 * CTA pairs (tcgen05.mma.cta_group::2, 256 x 256 tiles) or single CTAs; operands K-major or MN-major (the same row-major
 * matrices read "transposed" by TMA + UMMA MN-major descriptors: no transposed copies).
 * gemm_bf16: out[M,N] (bf16) = act(opA(a) . opB(b) + bias[N]).  a: a_mn = 0 [M,K] | a_mn = 1 [K,M];  b: b_mn = 0 [N,K]
 *   (Linear weight, forward) | b_mn = 1 [K,N] (Linear weight [out = K, in = N], input gradient).  act 0 identity, 1 exact
 *   GELU (pre != NULL also receives gelu'(pre-activation)).  tile: 0 = automatic, else cta_group * 1000 + BN.
 * gemm_mul_colsum2: out = (a . opB(b)) * mult, colsum ACCUMULATED (see esvit_gemm_mul_colsum); ws fp32 [160 * N].
 * gemm_wgrad: dw[N,K] (fp32) (+)= dy[T,N]^T . x[T,K], split over T, deterministic fold of fp32 partial tiles held in ws
 *   (esvit_gemm_wgrad_ws_floats(N, K) fp32 elements).  All of M / N / K / T multiples of 8. */
int esvit_gemm_bf16(const void* a, const void* b, const float* bias, void* out, void* pre, long long M, int N, int K,
                    int a_mn, int b_mn, int act, int tile, void* stream);
int esvit_gemm_mul_colsum2(const void* a, const void* b, const void* mult, void* out, float* colsum, float* ws,
                           long long M, int N, int K, int b_mn, int tile, void* stream);
int esvit_gemm_wgrad_ws_floats(int N, int K);
int esvit_gemm_wgrad(const void* dy, const void* x, float* dw, float* ws, long long T, int N, int K, int accumulate,
                     int tile, void* stream);

/* ---- GELU (exact erf), bf16 ------------------------------------------------ models/swin_transformer.py:21-37 */
int esvit_gelu_fwd(const void* x, void* y, long long n, void* stream);
int esvit_gelu_bwd(const void* x, const void* dy, void* dx, long long n, void* stream);
/* gelu backward that also ACCUMULATES dbias fp32 [N] = column sums of dx for x bf16 [R,N]: the gradient of the fc1
 * bias (added by the GEMM epilogue) without a separate reduction kernel. */
int esvit_gelu_bwd_dbias(const void* x, const void* dy, void* dx, float* dbias, long long R, int N, void* stream);
/* dx = dy * gp (gp = stored local derivative, bf16 [R,N]); ACCUMULATES dbias fp32 [N] = column sums of dx. */
int esvit_mul_bwd_dbias(const void* gp, const void* dy, void* dx, float* dbias, long long R, int N, void* stream);

/* ---- DINOHead pieces ------------------------------------------------------ models/vision_transformer.py:403-417
 * l2norm: y = x / max(||x||, eps) rows (bf16); weight_norm: w(bf16) = v * g / ||v||_row (fp32 v [K,D], g [K]). */
int esvit_l2norm_fwd(const void* x, void* y, float* inv, float eps, long long R, int D, void* stream);
int esvit_l2norm_bwd(const void* x, const void* dy, const float* inv, void* dx, long long R, int D, void* stream);
int esvit_weight_norm_fwd(const float* v, const float* g, void* w, float* norm, long long K, int D, void* stream);
/* dw: bf16 [K,D], or fp32 when dw_is_f32 (the fp32 weight gradient of esvit_gemm_wgrad) */
int esvit_weight_norm_bwd(const float* v, const float* g, const float* norm, const void* dw, int dw_is_f32, float* dv,
                          float* dg, long long K, int D, void* stream);

/* ---- DINOLoss / DDINOLoss ---------------------------------------------------- main_esvit.py:620-648, :683-750
 * row_lse: lse[r] = log sum_k exp((x[r,k] - center[k]) * inv_temp)   (center NULL for student rows).
 * dino_ce_fwd: row_loss[r] = n_r*lse_s[r] - sum_j <softmax((t[trow[r][j]]-center)*inv_temp_t), s[r]*inv_tau_s>;
 *   lse_s[r] = LSE(s[r]*inv_tau_s) is an OUTPUT (computed in the same pass, kept for the backward)
 * dino_ce_bwd: ds[r] = gscale[0]*w[r]*inv_tau_s * (n_r*softmax(s[r]*inv_tau_s) - sum_j q_j)   (bf16 out)
 * trow int32 [R,2], -1 = no pair.  s/t bf16 [R,K]/[Rt,K], K % 8 == 0.
 * order int32 [R] or NULL: CTA i works on row order[i] (a permutation; image-major keeps the paired teacher rows in L2). */
int esvit_row_lse(const void* x, const float* center, float inv_temp, float* lse, long long R, int K, void* stream);
int esvit_dino_ce_fwd(const void* s, const void* t, const float* center, float* lse_s, const float* lse_t,
                      const int* trow, const int* order, float inv_temp_t, float inv_tau_s, float* row_loss, long long R,
                      int K, void* stream);
int esvit_dino_ce_bwd(const void* s, const void* t, const float* center, const float* lse_s, const float* lse_t,
                      const int* trow, const int* order, const float* w, const float* gscale, float inv_temp_t,
                      float inv_tau_s, void* ds, long long R, int K, void* stream);
int esvit_weighted_sum(const float* v, const float* w, int R, float* out, void* stream);
/* The same loss with the teacher probabilities stored once per teacher row (every teacher row is paired with ~3.5 student
 * rows; recomputing its exponentials per pairing made both CE kernels SFU-bound):
 * row_softmax_q: lse[r] as esvit_row_lse, q[r,k] = 2^12 * softmax((x[r] - center) * inv_temp)_k in fp16 [R,K]
 *   (K <= esvit_row_softmax_q_max_k()).
 * dino_ce_q_fwd / bwd: esvit_dino_ce_fwd / bwd with q (that fp16 tensor) in place of (t, center, lse_t, inv_temp_t). */
int esvit_row_softmax_q_max_k(void);
int esvit_row_softmax_q(const void* x, const float* center, float inv_temp, float* lse, void* q, long long R, int K,
                        void* stream);
int esvit_dino_ce_q_fwd(const void* s, const void* q, float* lse_s, const int* trow, const int* order, float inv_tau_s,
                        float* row_loss, long long R, int K, void* stream);
int esvit_dino_ce_q_bwd(const void* s, const void* q, const float* lse_s, const int* trow, const int* order,
                        const float* w, const float* gscale, float inv_tau_s, void* ds, long long R, int K, void* stream);

/* ---- update_center ----------------------------------------------------------- main_esvit.py:650-660, :752-770
 * colsum: out[k] = sum_r t[r,k] (deterministic two-stage); workspace fp32 [esvit_colsum_workspace_rows()*K].
 * center_ema: center_out = center*m + (colsum/rows_total)*(1-m)  (after the caller's SUM all-reduce of colsum);
 * out-of-place like the reference's rebinding, because the loss backward still reads the old center. */
int esvit_colsum_workspace_rows(void);
int esvit_colsum(const void* t, long long R, int K, float* workspace, float* out, void* stream);
int esvit_center_ema(const float* center, const float* colsum, float rows_total, float momentum, float* center_out,
                     int K, void* stream);

/* ---- DDINOLoss region match -------------------------------------------------------- main_esvit.py:735-736
 * normalize_rows: y = x / max(||x||, eps), fp32 [R,P].
 * region_match: for every student region token of every crop v != iq, the FIRST arg-max over the Tg teacher tokens
 * of view iq (same image) of the cosine similarity.  sn fp32 [Rs,P] rows ordered (crop, image, token) with 2 global
 * crops of Tg tokens then ncrops-2 local crops of Tl tokens; tn fp32 [2*B*Tg, P].
 * idx_out int64 [2, ncrops, B, Tg] (slots of v == iq or i >= T_v untouched); trow int32 [Rs,2] teacher region rows. */
int esvit_normalize_rows(const float* x, float* y, long long R, int P, float eps, void* stream);
int esvit_region_match(const float* sn, const float* tn, int B, int ncrops, int Tg, int Tl, int P,
                       long long* idx_out, int* trow, void* stream);

/* ---- optimiser-side multi-tensor kernels (host arrays of device pointers) ----------------------------------
 * ema_multi: teacher = teacher*m + student*(1-m), bit-exact with main_esvit.py:587-590.
 * clip_multi: per-tensor L2 clip of utils.py:106-115; sumsq_ws double[n] workspace; norms fp32[n] or NULL. */
int esvit_ema_multi(void* const* teacher, const void* const* student, const long long* numel, int n, double momentum,
                    void* stream);
int esvit_clip_multi(void* const* grads, const long long* numel, int n, float clip, double* sumsq_ws, float* norms,
                     void* stream);
/* Fused optimiser pass, CUDA-graph friendly (every step-varying scalar is read from device memory):
 * grad_sumsq_multi: sumsq[i] = ||grads[i]||^2 (double[n], zeroed inside).
 * adamw_ema_multi: per-tensor clip (utils.py:106-115) folded into torch.optim.AdamW's update (main_esvit.py:411) and the
 *   teacher EMA (main_esvit.py:587-590, bit-exact two-rounding form) in one sweep.
 *   hyper fp32[8] = {lr, wd(group 0), beta1, beta2, eps, ema_m, 1-ema_m, clip(<=0: off)};
 *   state fp32[2n] = per tensor {step count, flags: bit0 weight-decayed, bit1 skip (= reference's p.grad=None)};
 *   teacher may be NULL; param_bf16 / teacher_bf16 (arrays or entries may be NULL) receive bf16 copies of the updated
 *   values = the GEMM operands of the next step, so no per-step cast kernels are needed. */
int esvit_grad_sumsq_multi(void* const* grads, const long long* numel, int n, double* sumsq, void* stream);
int esvit_adamw_ema_multi(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                          void* const* teacher, void* const* param_bf16, void* const* teacher_bf16,
                          const long long* numel, int n, const float* hyper, float* state, const double* sumsq,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ESVIT_B200_H */
