/* libvlmb200 — C ABI of the B200-native (sm_100a) kernels behind ViLMedic's vision-language hot path.
 *
 * Boundary contract: extern "C", plain pointers + sizes, no torch types.  The caller (PyTorch, via
 * ctypes — see vilmedic_b200/_lib.py — or any other host) owns every buffer and passes its CUDA stream as `void*`
 * (cudaStream_t).  No compute entry point allocates or synchronises; state kept by the library: per-device caches, the two launch
 * policy knobs (vlm_set_sm_margin, vlm_set_background) and the peer-visible buffers a caller explicitly requests with vlm_ipc_*.  Every
 * function returns 0 on success, a negative code on failure; the message is available from vlm_last_error()
 * (thread-local).  Device pointers must be 16-byte aligned unless stated otherwise.
 *
 * The reference (jbdel/vilmedic @ /root/reference) is 100% Python and ships no native interface; each entry point
 * below cites the reference call site (and the HuggingFace arithmetic it delegates to) that it replaces.
 * "HF:" = transformers/models/..., as pinned by the reference's setup.py:28.
 */
#ifndef VLM_B200_H_
#define VLM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLM_B200_ABI_VERSION 4

/* ---- runtime ---------------------------------------------------------------------------------------------------- */
const char* vlm_last_error(void);
int vlm_abi_version(void);
/* SMs that every persistent kernel of this library leaves free (grid = SM count - margin) for concurrent work such as the NCCL
 * kernels of the data-parallel gradient exchange (vilmedic_b200/ddp.py).  Default: env VLM_SM_MARGIN, else 0.  Set it before a
 * CUDA graph is captured — grid sizes are baked into the graph. */
int vlm_set_sm_margin(int margin);
int vlm_get_sm_margin(void);
/* Background mode (process-wide flag, returns the previous value): while on, vlm_optim_step and vlm_colsum_bf16 launch as many
 * small CTAs (128 threads, few registers) that fit on an SM next to a resident persistent GEMM / attention CTA — for work the
 * host issues on a side stream under the backward pass (vilmedic_b200/ddp.py, nn.py). Results are identical in either mode. */
int vlm_set_background(int on);
/* 0 iff the current CUDA device is sm_100 (B200). */
int vlm_device_check(void);

/* ---- GEMM (tcgen05 + TMA) --------------------------------------------------------------------------------------- */
/* C[M,N] = epi(alpha * A'[M,K] * B'[N,K]^T), bf16 operands, fp32 accumulate in TMEM.
 *   a_mn_major=0: A' stored [M][lda] (k contiguous); 1: stored [K][lda] (m contiguous).  Same for b / N.
 *   epi: (+bias[N] fp32) -> act (0 none | 1 GELU-erf; if aux_out is non-null the DERIVATIVE GELU'(pre-activation) is
 *        stashed there (bf16), so that the backward epilogue is a plain multiply | 2 multiply by aux_in, i.e. by the
 *        stashed GELU') -> dropout(p_drop; Philox(seed, offset, (row*N+col)/4), same stream as
 *        vlm_dropout_bf16 on the flat [M,N] tensor) -> (+residual, dtype of C) -> (accumulate into C) -> store.
 *   alpha_ptr: optional device scalar multiplied into alpha (e.g. the upstream loss gradient, no host sync).
 *   batch>1: strided-batched; operands advance by *_batch_stride ELEMENTS per batch (bias is shared).
 *   force_bn: 0 = heuristic, else N-tile in {64,128,192,256}.  max_ctas: 0 = one per SM.
 * Replaces nn.Linear / torch.mm behind: HF:vit/modeling_vit.py:228-230,265-268,296-312 (ViT Q/K/V, out, FFN),
 *   HF:bert_generation/modeling_bert_generation.py:52-56,89-153,181-232,265-293,593-601 (decoder projections, LM head),
 *   vilmedic/blocks/vision/visual_encoder.py:119-122,139 (visual_projection),
 *   vilmedic/blocks/losses/selfsup/ConVIRTLoss.py:31, InfoNCELoss.py:13 (similarity matrix). */
int vlm_gemm_bf16(const void* a, long long lda, int a_mn_major, const void* b, long long ldb, int b_mn_major, void* c,
                  long long ldc, int c_is_fp32, int M, int N, int K, const float* bias, const void* residual,
                  long long ldr, int act, const void* aux_in, void* aux_out, long long ld_aux, float alpha,
                  const float* alpha_ptr, int accumulate, int batch, long long a_batch_stride, long long b_batch_stride,
                  long long c_batch_stride, long long aux_batch_stride, long long res_batch_stride, float p_drop,
                  unsigned long long seed, unsigned long long offset, const unsigned long long* rng_offset_ptr, int force_bn,
                  int max_ctas, void* stream);

/* ---- LayerNorm -------------------------------------------------------------------------------------------------- */
/* y = LN(x) * gamma + beta, biased variance, one warp per row; x bf16 or fp32, y bf16; mean/rstd (fp32 [M]) optional.
 * Replaces nn.LayerNorm in HF:vit/modeling_vit.py:333,340,455 and HF:bert_generation/modeling_bert_generation.py:52-56,
 * 288-292,410-429.  D % 8 == 0, D <= 2048. */
int vlm_layernorm_fwd(const void* x, int x_is_fp32, const float* gamma, const float* beta, void* y, float* mean,
                      float* rstd, int M, int D, float eps, void* stream);
/* dx = LN'(dy) (+ dres), dtype of x; dgamma/dbeta (fp32 [D]) are ACCUMULATED with atomics (caller zeroes).
 * Fused extras (all optional): dx_drop = dropout(dx) with the same Philox stream as the forward GEMM epilogue that
 * produced the LN input (p_drop, seed, offset, rng_offset_ptr); colsum[D] += column sums of dx_drop (or dx) — the
 * bias gradient of the Linear whose output fed this LayerNorm. */
int vlm_layernorm_bwd(const void* dy, const void* x, int x_is_fp32, const float* mean, const float* rstd,
                      const float* gamma, const void* dres, void* dx, float* dgamma, float* dbeta, int M, int D,
                      void* dx_drop, float p_drop, unsigned long long seed, unsigned long long offset,
                      const unsigned long long* rng_offset_ptr, float* colsum, void* stream);

/* ---- attention -------------------------------------------------------------------------------------------------- */
/* O = softmax(scale * Q K^T + mask) V per (batch, head); bf16 in/out, fp32 softmax.  q/k/v/o are addressed as
 * base + b*bs + row*rs + h*DH (elements), so packed QKV projections are consumed in place.  kmask: uint8 [B,Sk],
 * 1 = attend (key padding / encoder_attention_mask), may be null.  causal!=0 adds the j<=i mask.  lse: fp32 [B,H,Tq]
 * (natural log), needed by the backward.  p_drop: dropout on the probabilities (Philox seed/offset).
 * Replaces ALL_ATTENTION_FUNCTIONS['sdpa'|'eager'] behind HF:vit/modeling_vit.py:232-247 and
 * HF:bert_generation/modeling_bert_generation.py:114-153 (self, create_causal_mask :568-574), :181-232 (cross,
 * create_bidirectional_mask :582-588).  DH in {48, 64, 96}. */
int vlm_attention_fwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                      const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse,
                      const uint8_t* kmask, int B, int H, int Tq, int Sk, int DH, int causal, float scale, float p_drop,
                      unsigned long long seed, unsigned long long offset, const unsigned long long* rng_offset_ptr,
                      void* stream);
/* Gradients dq/dk/dv (bf16, same addressing scheme with their own strides); delta: fp32 [B,H,Tq] scratch. */
int vlm_attention_bwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                      const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                      const void* d_o, long long do_bs, long long do_rs, const float* lse, float* delta, void* dq,
                      long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs, void* dv,
                      long long dv_bs, long long dv_rs, const uint8_t* kmask, int B, int H, int Tq, int Sk, int DH,
                      int causal, float scale, float p_drop, unsigned long long seed, unsigned long long offset,
                      const unsigned long long* rng_offset_ptr, void* stream);

/* Same contract as vlm_attention_fwd, on the tcgen05 kernel (DH = 64, Sk <= 256: S = Q K^T and O = P V on the 5th-gen tensor
 * cores, S / O in TMEM, single-pass softmax).  vlm_attention_fwd routes supported shapes here when VLM_ATTN_TC=1. */
int vlm_attention_fwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                         const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse,
                         const uint8_t* kmask, int B, int H, int Tq, int Sk, int DH, int causal, float scale, float p_drop,
                         unsigned long long seed, unsigned long long offset, const unsigned long long* rng_offset_ptr,
                         void* stream);
/* Same contract as vlm_attention_bwd, on the tcgen05 kernel (DH = 64, Tq <= 256, Tq <= 128 with dropout): five tensor-core
 * products per 128-key tile with all accumulators in TMEM, delta computed in-kernel.  vlm_attention_bwd routes supported
 * shapes here when VLM_ATTN_TC=1. */
int vlm_attention_bwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                         const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                         const void* d_o, long long do_bs, long long do_rs, const float* lse, float* delta, void* dq,
                         long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs, void* dv,
                         long long dv_bs, long long dv_rs, const uint8_t* kmask, int B, int H, int Tq, int Sk, int DH,
                         int causal, float scale, float p_drop, unsigned long long seed, unsigned long long offset,
                         const unsigned long long* rng_offset_ptr, void* stream);

/* ---- softmax cross-entropy (LM head loss, label-smoothing CE) ---------------------------------------------------- */
/* One pass per row: loss_rows[r] (0 for ignored rows), lse_rows[r] (optional), and dlogits = (softmax - target) *
 * grad_scale (optional, same dtype as logits, may alias logits; columns [V, ldd) are zeroed).
 * shift_T > 0: ids is input_ids [R = B*T]; the label of row (b,t) is ids[b,t+1], the last position of each sequence
 * is ignored — HF:loss/loss_utils.py:45-66 reached via vilmedic/blocks/huggingface/decoder/decoder_model.py:46
 * (labels=input_ids, pads are NOT masked).  shift_T == 0: ids are labels [R], negative = ignore.
 * smoothing: vilmedic/blocks/losses/mvqa/LabelSmoothingCrossEntropyLoss.py:38-48.
 * row_weight (optional fp32 [R]): multiplies the gradient of row r (reward-weighted log-likelihood of vilmedic/blocks/rl/SCST.py:12-44);
 * -inf logits (filtered tokens) contribute nothing to the log-sum-exp and get a zero gradient. */
int vlm_softmax_ce(const void* logits, int logits_fp32, long long ld, const long long* ids, int shift_T, int R, int V,
                   float smoothing, float grad_scale, void* dlogits, long long ldd, float* loss_rows, float* lse_rows,
                   const float* row_weight, void* stream);

/* ---- helpers around the core ------------------------------------------------------------------------------------ */
int vlm_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
/* images fp32 [B,C,H,W] -> bf16 [B, n_prefix+(H/P)(W/P), C*P*P], the first n_prefix rows of every image = 0 (CLS slot; DeiT: CLS +
 * distillation token, HF:deit/modeling_deit.py embeddings); column order = Conv2d weight flatten (HF:vit/modeling_vit.py:151,166). */
int vlm_patchify(const float* images, void* patches, int B, int C, int H, int W, int P, int n_prefix, void* stream);
/* x[b,row,:] = tok + pos[row]  (HF:vit/modeling_vit.py:117-124; row 0 = CLS, DeiT row 1 = distillation token). */
int vlm_vit_cls_pos(void* x, int x_is_fp32, const float* tok, const float* pos, int B, int S, int D, int row, void* stream);
/* dpos[s] += sum_b dx[b,s]; dcls += sum_b dx[b,0]; ddist += sum_b dx[b,1] (n_prefix == 2); dbias += sum_{b,s>=n_prefix} dx[b,s]
 * (any of the outputs may be null). */
int vlm_vit_embed_bwd(const void* dx, int dx_is_fp32, float* dpos, float* dcls, float* ddist, float* dbias, int B, int S, int D,
                      int n_prefix, void* stream);
/* out[n] += (scale_ptr ? *scale_ptr : 1) * sum_m x[m,n]  (bias gradients; caller zeroes). */
int vlm_colsum_bf16(const void* x, long long ld, float* out, int M, int N, const float* scale_ptr, void* stream);
/* mask[r] = (sum_d |f[r,d]| != 0)  — vilmedic/blocks/vision/visual_encoder.py:138. */
int vlm_features_mask(const void* feats, uint8_t* mask, int R, int D, void* stream);
/* z[r] = word[ids[r]] (+ tt_row) + pos[pos_ids ? pos_ids[r] : pos_offset + r % T]  (HF:bert_generation/modeling_bert_generation.py:
 * 410-429, pre-LN).  pos_ids (optional int32 [R]): explicit positions — RoBERTa's padding-aware ids (HF:roberta/modeling_roberta.py
 * create_position_ids_from_input_ids); tt_row (optional fp32 [D]): token_type_embeddings[0] of BERT / RoBERTa checkpoints loaded through
 * `proto` (vilmedic/blocks/huggingface/encoder/encoder_model.py:20-22, decoder/decoder_model.py:17-21). */
int vlm_embed_fwd(const long long* ids, const float* word, const float* pos, void* z, int R, int T, int D, int V,
                  int pos_offset, const int* pos_ids, const float* tt_row, void* stream);
/* scatter-add of dz into dword / dpos (fp32, atomics; either may be null); rows with id == padding_idx (< 0: none) get no
 * gradient, as nn.Embedding(padding_idx=pad_token_id) in HF:bert_generation/modeling_bert_generation.py:400. */
int vlm_embed_bwd(const long long* ids, const void* dz, float* dword, float* dpos, int R, int T, int D, int V,
                  int pos_offset, int padding_idx, const int* pos_ids, void* stream);
/* y = x * keep / (1-p), keep ~ Philox(seed, offset, element); same call on grads is the backward.  n % 8 == 0. */
int vlm_dropout_bf16(const void* x, void* y, long long n, float p, unsigned long long seed, unsigned long long offset,
                     const unsigned long long* rng_offset_ptr, void* stream);
/* *counter += delta on the device.  Every dropout-capable entry point takes `rng_offset_ptr`: a device uint64 added to
 * `offset`, so a CUDA-graph replay of a training step draws fresh masks (advance the counter once per step). */
int vlm_rng_advance(unsigned long long* counter, unsigned long long delta, void* stream);
/* y[r,:] = mask[r / rows_per_mask] ? x[r,:] : 0 — multi-image masking, vilmedic/blocks/vision/visual_encoder.py:170-171. */
int vlm_mask_rows_bf16(const void* x, void* y, const uint8_t* mask, int R, int D, int rows_per_mask, void* stream);
/* fp32 activations: kind 0 = tanh (BertPooler, vilmedic/blocks/huggingface/encoder/encoder_model.py:58-60), 1 = ReLU
 * (ConVIRT projection heads, vilmedic/models/selfsup/conVIRT.py:58-67).  The backward takes the forward output y. */
int vlm_act_fwd_f32(const float* x, float* y, long long n, int kind, void* stream);
int vlm_act_bwd_f32(const float* dy, const float* y, float* dx, long long n, int kind, void* stream);
/* out[0] = scale * sum(x[0..n))  (deterministic single-block reduction; mean of per-row losses). */
int vlm_sum_scale_f32(const float* x, int n, float scale, float* out, void* stream);

/* ---- contrastive losses (ConVIRT / InfoNCE / GLoRIA-global) ------------------------------------------------------ */
/* x fp32 [N,D] (optionally L2-normalised per row, clamp eps — ConVIRTLoss.py:25-31) split into bf16 hi/lo and packed for the
 * 3-term tensor-core product: xa [N,3D] = [hi|hi|lo], xb [N,3D] = [hi|lo|hi], xh [N,D] = hi; any output may be null. */
int vlm_rownorm_split(const float* x, void* xa, void* xb, void* xh, float* inv_norm, int N, int D, int normalize, float eps,
                      void* stream);
/* backward of the normalisation: dx = inv_norm * (dxh - xhat (xhat . dxh)). */
int vlm_rownorm_bwd(const float* x, const float* inv_norm, const float* dxh, float* dx, int N, int D, int normalize,
                    void* stream);
/* S fp32 [N,N]: lse_row[i] = LSE_j(scale*S_ij), lse_col[i] = LSE_j(scale*S_ji), loss_row/col = lse - scale*S_ii.
 * ConVIRTLoss.py:13-21 (scale 1/tau), InfoNCELoss.py:15-16 (scale 1), GLoRIALoss.py:72-75 (scale temp3). */
int vlm_sym_lse(const float* S, int N, long long ld, float scale, float* lse_row, float* lse_col, float* loss_row,
                float* loss_col, void* stream);
/* dS (bf16 [N,ldd]) = g * scale * (w_row (softmax_row - I) + w_col (softmax_col - I)); g_ptr = upstream scalar or null. */
int vlm_sym_lse_bwd(const float* S, int N, long long ld, float scale, const float* lse_row, const float* lse_col, float w_row,
                    float w_col, const float* g_ptr, void* dS, long long ldd, void* stream);

/* ---- GLoRIA local (word x region) loss — vilmedic/blocks/losses/selfsup/GLoRIALoss.py:13-51,78-129 ---------------- */
/* All (image i, caption j) pairs are evaluated at once instead of the reference's Python loop over captions (:86-119).
 * Layouts: Xc bf16 [B,S,D] regions; Ww bf16 / Q fp32 [NL = B*L, D] words (row j*L+w; rows w >= cap_lens[j] zero);
 * A/P1 fp32 [B*S, NL]; P2 fp32 + bf16 [B,S,NL]; WC fp32 [B,NL,D]; cos / wnorm fp32 [B,NL]; sims fp32 [B,B].
 * The GEMMs between these kernels are vlm_gemm_bf16 calls (see vilmedic_b200/blocks/losses/gloria.py). */
/* out[b][c][r] = in[b][r][c] (fp32 in; bf16 or fp32 out); rows c in [C,Cout) and c >= row_limit[b] are zero-filled.
 * Converts the reference's [B,D,ih*iw] / [B,D,Lw] feature layouts (:19-25, :92) to K-major rows and back.  With bf16 output,
 * out_lo (optional) receives bf16(x - hi) so that hi*hi + lo*hi + hi*lo tensor-core products carry ~16 mantissa bits. */
int vlm_transpose_cast(const float* in, void* out, void* out_lo, int out_bf16, int batch, int R, int C, long long in_ld, long long in_bs,
                       long long out_ld, long long out_bs, int Cout, const int* row_limit, void* stream);
/* P1 = softmax over the words of each caption (:32-33). */
int vlm_gloria_word_softmax(const float* A, float* P1, const int* cap_lens, int rows, int NB, int L, long long ld, void* stream);
/* P2 = softmax over regions of temp1 * P1, per (image, caption, word) (:38-43); fp32 and bf16 hi/lo copies. */
int vlm_gloria_region_softmax(const float* P1, float* P2, void* P2h, void* P2l, const int* cap_lens, int NI, int S, int NB, int L,
                              float temp1, void* stream);
/* cos[i,(j,w)] = q.wc / max(|q||wc|, eps)  (cosine_similarity, :5-10, called at :109); also the two norms. */
int vlm_gloria_cos(const float* WC, const float* Q, const int* cap_lens, float* cosv, float* wnorm, float* qnorm, int NI, int NB,
                   int L, int D, float eps, void* stream);
/* sims[i,j] = temp3 * log sum_w exp(temp2 * cos[i,(j,w)])  (:112-122, agg "sum"). */
int vlm_gloria_sims(const float* cosv, const int* cap_lens, float* sims, int NB, int L, float temp2, float temp3, void* stream);
/* Backward through CE (both directions, means over B; g0/g1 = upstream scalars or null = 1), log-sum-exp and cosine:
 * dWC bf16 [B,NL,D] and the direct word gradient dQ fp32 [NL,D]. */
int vlm_gloria_cos_bwd(const float* WC, const float* Q, const int* cap_lens, const float* cosv, const float* wnorm,
                       const float* qnorm, const float* sims, const float* lse_row, const float* lse_col, const float* g0,
                       const float* g1, void* dWC, float* dQ, int NB, int L, int D, float temp2, float temp3, float eps,
                       void* stream);
/* G [B,S,ld] (= dL/dP2) <- temp1 * P2 (G - sum_s P2 G), in place (= dL/dP1). */
int vlm_gloria_region_softmax_bwd(const float* P2, float* G, int NI, int S, long long ld, float temp1, void* stream);
/* dA bf16 = P1 (G - sum_w P1 G) per caption segment. */
int vlm_gloria_word_softmax_bwd(const float* P1, const float* G, void* dA, const int* cap_lens, int rows, int NB, int L,
                                long long ld, void* stream);

/* ---- CNN backbone pieces (torchvision ResNet-18/50 of vilmedic/blocks/vision/visual_encoder.py:71-83) --
 * Convolutions = vlm_gemm_bf16 over NHWC bf16 activations [B*H*W, C] (C % 8 == 0) and the im2col matrix
 * col[m, (kh*KW + kw)*C + c]; weights are packed from torchvision's OIHW fp32 masters into [Cout, Kp] bf16 in the same
 * (kh, kw, ci) order (Kp = kh*kw*Cin rounded up to 8, zero padded). */
int vlm_conv_weight_pack(const float* w_oihw, void* wm_bf16, int Cout, int Cin, int KH, int KW, int Kp, void* stream);
/* gw_oihw[co,ci,kh,kw] += dwm[co, (kh*KW + kw)*Cin + ci]  (dwm fp32 [Cout, Kp] from the wgrad GEMM). */
int vlm_conv_wgrad_unpack(const float* dwm, float* gw_oihw, int Cout, int Cin, int KH, int KW, int Kp, void* stream);
/* x bf16 [B,H,W,C] -> col bf16 [B*Ho*Wo, KH*KW*C], Ho = (H + 2 pad - KH)/stride + 1; zero outside the image. */
int vlm_im2col_nhwc(const void* x, void* col, int B, int H, int W, int C, int KH, int KW, int stride, int pad, void* stream);
/* Stem: fp32 NCHW images (what the reference's datasets hand over) -> col bf16 [B*Ho*Wo, Kp]. */
int vlm_im2col_nchw_f32(const float* img, void* col, int B, int Cin, int H, int W, int KH, int KW, int stride, int pad, int Kp,
                        void* stream);
/* Transposed gather: dx bf16 [B,H,W,C] = (add or 0) + sum of the dcol entries every input pixel contributed to. */
int vlm_col2im_nhwc(const void* dcol, const void* add, void* dx, int B, int H, int W, int C, int KH, int KW, int stride, int pad,
                    void* stream);
/* nn.BatchNorm2d, training mode, on x bf16 [M = B*H*W, C]: batch statistics (fp32), y = relu?(bn(x) (+ res)), running statistics
 * updated as torch does (momentum, unbiased variance, num_batches_tracked).  mean/rstd/scale/shift: fp32 [C] outputs (saved for
 * the backward); sum_ws: fp32 [2C] workspace. */
int vlm_bn_train_fwd(const void* x, const void* res, void* y, const float* gamma, const float* beta, float* mean, float* rstd,
                     float* scale, float* shift, float* sum_ws, float* running_mean, float* running_var, long long* num_batches,
                     int M, int C, float eps, float momentum, int relu, void* stream);
/* Evaluation mode: running statistics. */
int vlm_bn_eval_fwd(const void* x, const void* res, void* y, const float* gamma, const float* beta, const float* running_mean,
                    const float* running_var, float* scale, float* shift, int M, int C, float eps, int relu, void* stream);
/* Backward of relu?(bn(x) (+ res)): g = dy masked by y > 0 (if relu); dx (bf16), dres = g (bf16, optional), dgamma/dbeta +=. */
int vlm_bn_train_bwd(const void* dy, const void* y, const void* x, const float* mean, const float* rstd, const float* gamma,
                     float* dgamma, float* dbeta, float* sum_ws, void* dx, void* dres, int M, int C, int relu, void* stream);
/* nn.MaxPool2d(3, 2, 1) on NHWC bf16; idx uint8 [B*Ho*Wo, C] = window position (kh*3 + kw) of the first maximum (ATen's rule). */
int vlm_maxpool3x3s2_fwd(const void* x, void* y, uint8_t* idx, int B, int H, int W, int C, void* stream);
int vlm_maxpool3x3s2_bwd(const void* dy, const uint8_t* idx, void* dx, int B, int H, int W, int C, void* stream);
/* nn.AdaptiveAvgPool2d((1, 1)): y[b, c] = mean over the HW positions. */
int vlm_avgpool_fwd(const void* x, void* y, int B, int HW, int C, void* stream);
int vlm_avgpool_bwd(const void* dy, void* dx, int B, int HW, int C, void* stream);

/* ---- incremental decoding + device-side beam search (csrc/decode.cu) ------------------------------------------------------ */
/* One decode step of the loop behind vilmedic/blocks/huggingface/decoder/evaluation.py:73-78 (HF cached generate) and the
 * ensemble search of vilmedic/blocks/huggingface/decoder/beam_search.py:243-320, with every per-step scalar in device memory
 * (counters[0] = t = position of the token being consumed), so that the whole step replays as one CUDA graph.
 * vlm_embed_step: z[r] = bf16(word[tok[r]] (+ tt_row) + pos[min(*t + pos_shift, max_pos-1)])
 *   (HF:bert_generation/modeling_bert_generation.py:395-429; pos_shift = padding_idx + 1 and tt_row for RoBERTa-family checkpoints). */
int vlm_embed_step(const long long* tok, const float* word, const float* pos, void* z, int R, int D, int V, const int* t_ptr,
                   int max_pos, int pos_shift, const float* tt_row, void* stream);
/* T_q = 1 attention for R rows x H heads (DH in {48, 64, 96}); fp32 scores / softmax / PV, bf16 out [R, H*DH].
 *   self-attention (kv_new != NULL): cache bf16 [R, max_len, 2*H*DH] ([K | V] per position); row r's history at position j < *t is
 *     read from physical row row_map[r*map_ld + j]; the new key/value kv_new[r] ([K | V], pitch ld_new) is used for position *t,
 *     appended at (r, *t), and row_map[r][*t] = r is recorded.  Beam reordering permutes row_map, never the cache
 *     (replaces the index_select cache reorder of beam_search.py:317-319).
 *   cross-attention (kv_new == NULL): fixed_len keys, K/V of row r read from cache row r / row_div (one projection per image,
 *     shared by its beams); kmask uint8 [cache rows, kmask_ld] (0 = masked) optional. */
int vlm_decode_attention(const void* q, long long ldq, const void* kv_new, long long ld_new, void* cache, long long cache_row_stride,
                         int two_d, int* row_map, int map_ld, const int* t_ptr, int fixed_len, int row_div, const uint8_t* kmask,
                         int kmask_ld, void* out, long long ldo, int R, int H, int DH, float scale, int max_len, void* stream);
/* Per row r: x = sum_m logits[m][r, :V] (beam_search.py:254), score = log_softmax(x) + beam_scores[r] (:260-265); writes the row's
 * top-2k (score desc, token asc) to cand_score / cand_tok [R, 2k].  logits: HOST array of n_models (<= 8) device pointers, fp32,
 * row pitch ld.  k in 1..8. */
int vlm_beam_rows(const float* const* logits, int n_models, long long ld, int V, const float* beam_scores, float* cand_score,
                  int* cand_tok, int R, int k, void* stream);
/* Per batch element: global top-2k over its k rows' candidates (== torch.topk over k*V, :289-294; ties: lower beam*V+token first),
 * then BeamSearchScorer.process (:297-304; legacy BeamHypotheses semantics: EOS candidates ranked >= k are skipped, a finished
 * hypothesis scores sum_logprobs / len**length_penalty (double), is_done when k are finished and the worst kept score >= the best
 * running score / len**length_penalty).  Writes next_tok / parent / beam_scores [B*k]; k == 1 is greedy argmax decoding (finished
 * rows emit pad; forced_last_token >= 0 is emitted at the last position like HF's ForcedEOSTokenLogitsProcessor, -1 = off).
 * counters: {t, number of finished batch elements, sequence length at which the last one finished, -}. */
int vlm_beam_select(const float* cand_score, const int* cand_tok, int k, int V, int B, int max_len, const long long* ids,
                    float* beam_scores, uint8_t* done, long long* next_tok, int* parent, double* hyp_score, int* hyp_len,
                    long long* hyp_tok, int* hyp_count, double* hyp_worst, int* counters, int eos, int pad, double length_penalty,
                    int forced_last_token, void* stream);
/* Sampling rollouts of SCST (vilmedic/blocks/rl/SCST.py:139-153, generate(do_sample=True, top_k, bad_words_ids)).
 * vlm_logits_filter (in place, bf16 or fp32 rows): HF NoBadWordsLogitsProcessor for <= 8 single-token ids (HOST array) -> -inf, then
 *   TopKLogitsWarper: scores below the k-th largest -> -inf (ties kept), top_k = 0 off.
 * vlm_sample_rows: one token per row ~ softmax(logits / temperature) (Gumbel-max over Philox(seed, offset + counters[0] + (counters[3] << 16),
 *   row, v); t_ptr = the search's 4-int counters: [0] step, [3] per-rollout nonce), written
 *   with its log-probability to slot 0 of cand_tok / cand_score [R, 2]; vlm_beam_select(k = 1) then does the EOS / pad bookkeeping. */
int vlm_logits_filter(void* logits, int logits_fp32, long long ld, int R, int V, const int* bad_ids, int n_bad, int top_k, void* stream);
int vlm_sample_rows(const float* logits, long long ld, int V, float temperature, unsigned long long seed, unsigned long long offset,
                    const int* t_ptr, float* cand_score, int* cand_tok, int R, void* stream);
/* ids[r] <- ids[parent[r]] + next_tok[r]; row_map[r] <- row_map[parent[r]] (through the tmp buffers); counters[0] += 1. */
int vlm_beam_advance(long long* ids, long long* ids_tmp, int* row_map, int* map_tmp, const int* parent, const long long* next_tok, int R,
                     int max_len, int* counters, void* stream);

/* ---- input pipeline (vilmedic/datasets/base/ImageDataset.py:97-104) ------------------------------------ */
/* RandomCrop + RandomHorizontalFlip + ToTensor + Normalize of the reference's train transform, after its (host-side) Resize:
 * in  uint8 [B, Hin, Win, 3] (HWC, what PIL / numpy hand over; device memory), top/left int32 [B] crop origins,
 * flip uint8 [B]; out fp32 [B, 3, crop, crop] = ((in / 255) - mean[c]) / std[c] with IEEE fp32 division — bit-identical to
 * torchvision's ToTensor().div(255) -> Normalize sub_().div_().  mean3 / std3 are HOST arrays of 3 floats.  Byte work,
 * HBM-bound: reads 3 B, writes 12 B per pixel. */
int vlm_image_crop_flip_normalize(const uint8_t* in, float* out, const int* top, const int* left, const uint8_t* flip, int B,
                                  int Hin, int Win, int crop, const float* mean3, const float* std3, void* stream);

/* Resize of the same transform on the device: one pass of Pillow's separable fixed-point resampling (what torchvision's Resize does to
 * a PIL image) along one axis of uint8 [B, H, W, 3] images; axis 0 resamples columns (Hout == Hin), axis 1 rows (Wout == Win).
 * bounds int32 [out, 2] = (first source index, tap count), coefs int32 [out, ksize] = round(k * 2^22), computed on the host exactly as
 * Pillow does (vilmedic_b200/blocks/vision/preprocess.py: pil_resample_tables); out = clip8((2^21 + sum pix * coef) >> 22).
 * Horizontal pass first, then vertical, like Pillow — bit-identical to PIL.Image.resize(size, BILINEAR). */
int vlm_image_resample_u8(const uint8_t* in, uint8_t* out, const int* bounds, const int* coefs, int ksize, int B, int Hin, int Win,
                          int Hout, int Wout, int axis, void* stream);

/* ---- optimizer (vilmedic/executors/trainor.py:119-124) ------------------------------------------ */
/* out[0] += sum(g^2)  (caller zeroes). */
int vlm_sumsq_f32(const float* g, long long n, float* out, void* stream);
/* same over a bf16 buffer (the all-reduced gradient payload of the data-parallel step, vilmedic_b200/ddp.py). */
int vlm_sumsq_bf16(const void* g, long long n, float* out, void* stream);
/* AdamW over flat fp32 buffers (torch.optim.AdamW semantics), writes the bf16 mirror of p, optional global-norm clip
 * (gnorm_sq_ptr + max_norm), grad pre-scale (1/world, 1/grad_accu, 1/loss_scale) and fused zero_grad. */
int vlm_adamw_step(float* p, float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int* step_ptr, int increment_step, const float* lr_scale_ptr,
                   float grad_scale, const float* gnorm_sq_ptr, float max_norm, int zero_grad, void* stream);

/* The optimizers the reference's configs select by name (vilmedic/executors/utils.py:81-86 `getattr(torch.optim, name)`;
 * config/ uses RAdam and Adam): kind 0 = AdamW, 1 = Adam (L2 weight decay), 2 = RAdam (torch defaults: L2 decay, rectified once
 * rho_t > 5) — torch/optim/{adamw,adam,radam}.py single-tensor semantics over a flat span.  Same fusions as vlm_adamw_step.
 * Device-side skip instead of the host syncs of vilmedic/executors/trainor.py:109-112: when *loss_ptr or *gnorm_sq_ptr is
 * NaN/Inf the span is left untouched (gradients still zeroed).  Call vlm_optim_step_begin ONCE per optimizer step before the
 * span launches: it advances *step_ptr (or bumps *skip_count when the step is skipped).
 * g_bf16 (nullable): when set, the gradient VALUES are read from this bf16 buffer (the exchanged payload of the data-parallel
 * step — the reference's DDP all-reduce, vilmedic/executors/trainor_accelerate.py:132, at half the bytes) and `g` is only zeroed. */
int vlm_optim_step_begin(int* step_ptr, const float* gnorm_sq_ptr, const float* loss_ptr, int* skip_count, void* stream);
int vlm_optim_step(int kind, float* p, float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, const int* step_ptr, const float* lr_scale_ptr, float grad_scale,
                   const float* gnorm_sq_ptr, float max_norm, const float* loss_ptr, int zero_grad, const void* g_bf16,
                   void* stream);

/* ---- data-parallel gradient exchange over peer memory (NVLink / NVSwitch), fused into the optimizer ----------------- */
/* Replaces the NCCL all-reduce of DistributedDataParallel / accelerate (vilmedic/executors/trainor_accelerate.py:122,132) followed
 * by torch.optim.step (vilmedic/executors/trainor.py:119-124): every rank publishes the bf16 cast of a gradient bucket in a buffer
 * its peers have mapped with CUDA IPC and raises READY[bucket][rank] = epoch in every rank's flag block; vlm_optim_step_p2p polls
 * the `world` READY flags of its bucket in LOCAL memory, then reads the bucket of every rank with peer loads, sums in fp32 in rank
 * order (replicas stay bit-identical) and applies the update of vlm_optim_step in the same pass.  Flag block of a rank: int32
 * [slots][world].  Epoch: device int, advanced once per step (vlm_p2p_epoch_inc) so that the step replays as a CUDA graph.
 * Waits are bounded (90 s): on expiry *err_flag is set to 1 and the kernel returns without touching the parameters.
 *   vlm_ipc_alloc  cudaMalloc (zero-filled) + the 64-byte cudaIpcMemHandle_t of the allocation
 *   vlm_ipc_open   map a peer's allocation into this process (peer access enabled lazily); vlm_ipc_close unmaps it
 *   vlm_p2p_signal flags[w][slot][rank] = *epoch for every rank w (system-scope release after a fence); peer_flags = host array of
 *                  `world` device pointers;  vlm_p2p_wait: until my flags[slot][w] >= *epoch + epoch_delta for all w */
int vlm_ipc_alloc(long long bytes, void** dev_ptr, void* handle64);
int vlm_ipc_open(const void* handle64, void** dev_ptr);
int vlm_ipc_close(void* dev_ptr);
int vlm_ipc_free(void* dev_ptr);
int vlm_p2p_epoch_inc(int* epoch_ptr, void* stream);
int vlm_p2p_signal(void* const* peer_flags, int world, int rank, int slot, const int* epoch_ptr, void* stream);
int vlm_p2p_wait(const int* my_flags, int world, int slot, const int* epoch_ptr, int epoch_delta, int* err_flag, void* stream);
/* peer_g16: host array of `world` device pointers to every rank's bf16 gradient buffer (whole arena); elem_offset: first element
 * of this span in those buffers (multiple of 4); p/g/m/v/p_bf16 point at the span itself.  No clipping / loss skip on this path.
 * One-shot (peer_r32 == NULL): the kernel sums the `world` bf16 buckets itself — (N-1) * 2 bytes per parameter over NVLink.
 * Two-shot (peer_r32 = host array of every rank's fp32 buffer of reduced slices): vlm_p2p_reduce_slice has summed this rank's slice
 * of the bucket (units of 4 elements [rank * units_per_rank, ...) counted from the bucket start) into its own fp32 buffer and
 * raised the bucket's REDUCED flag; the update kernel waits for the REDUCED flags (`slot`) and reads unit i of the span from rank
 * (unit_base + i) / units_per_rank (unit_base = span start - bucket start, in units) — (N-1)/N * 6 bytes per parameter. */
int vlm_p2p_reduce_slice(const void* const* peer_g16, long long elem_offset, float* out, long long n, int world,
                         const int* my_flags, int slot, const int* epoch_ptr, int* err_flag, void* stream);
int vlm_optim_step_p2p(int kind, float* p, float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, const int* step_ptr, const float* lr_scale_ptr,
                       float grad_scale, const void* const* peer_g16, const void* const* peer_r32, long long elem_offset,
                       long long unit_base, long long units_per_rank, int world, const int* my_flags,
                       int slot, const int* epoch_ptr, int* err_flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VLM_B200_H_ */
