/* libvlmb200 — C ABI of the B200-native (sm_100a) kernels behind ViLMedic's vision-language hot path.
 *
 * Boundary contract: extern "C", plain pointers + sizes, no torch types.  The caller (PyTorch, via
 * ctypes — see vilmedic_b200/_lib.py — or any other host) owns every buffer and passes its CUDA stream as `void*`
 * (cudaStream_t).  No entry point allocates, synchronises or keeps global state beyond per-device caches.  Every
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

#define VLM_B200_ABI_VERSION 1

/* ---- runtime ---------------------------------------------------------------------------------------------------- */
const char* vlm_last_error(void);
int vlm_abi_version(void);
/* 0 iff the current CUDA device is sm_100 (B200). */
int vlm_device_check(void);

/* ---- GEMM (tcgen05 + TMA) --------------------------------------------------------------------------------------- */
/* C[M,N] = epi(alpha * A'[M,K] * B'[N,K]^T), bf16 operands, fp32 accumulate in TMEM.
 *   a_mn_major=0: A' stored [M][lda] (k contiguous); 1: stored [K][lda] (m contiguous).  Same for b / N.
 *   epi: (+bias[N] fp32) -> act (0 none | 1 GELU-erf, pre-activation stashed to aux_out if non-null |
 *        2 multiply by GELU'(aux_in)) -> (+residual, dtype of C) -> (accumulate into C) -> store bf16 or fp32.
 *   batch>1: strided-batched; operands advance by *_batch_stride ELEMENTS per batch (bias is shared).
 *   force_bn: 0 = heuristic, else N-tile in {64,128,192,256}.  max_ctas: 0 = one per SM.
 * Replaces nn.Linear / torch.mm behind: HF:vit/modeling_vit.py:228-230,265-268,296-312 (ViT Q/K/V, out, FFN),
 *   HF:bert_generation/modeling_bert_generation.py:52-56,89-153,181-232,265-293,593-601 (decoder projections, LM head),
 *   vilmedic/blocks/vision/visual_encoder.py:119-122,139 (visual_projection),
 *   vilmedic/blocks/losses/selfsup/ConVIRTLoss.py:31, InfoNCELoss.py:13 (similarity matrix). */
int vlm_gemm_bf16(const void* a, long long lda, int a_mn_major, const void* b, long long ldb, int b_mn_major, void* c,
                  long long ldc, int c_is_fp32, int M, int N, int K, const float* bias, const void* residual,
                  long long ldr, int act, const void* aux_in, void* aux_out, long long ld_aux, float alpha,
                  int accumulate, int batch, long long a_batch_stride, long long b_batch_stride,
                  long long c_batch_stride, long long aux_batch_stride, long long res_batch_stride, int force_bn,
                  int max_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VLM_B200_H_ */
