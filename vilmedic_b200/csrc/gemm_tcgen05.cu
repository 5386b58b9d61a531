// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accumulators
// in TMEM, double buffered) -> tcgen05.ld epilogue with fused bias / GELU / GELU' / residual / accumulate.
//
//   C[M,N] = epilogue(alpha * sum_k A'[m,k] * B'[n,k])
//
// A' is given either K-major  (memory A[M][lda],  k contiguous)  or MN-major (memory A[K][lda], m contiguous);
// B' likewise: K-major = memory B[N][ldb] (the nn.Linear weight layout), MN-major = memory B[K][ldb].
// That covers all three products of a Linear layer without any transposed copies:
//   forward  Y  = X W^T          A=X  K-major,  B=W  K-major
//   dgrad    dX = dY W           A=dY K-major,  B=W  MN-major   (k' = out features)
//   wgrad    dW = dY^T X         A=dY MN-major, B=X  MN-major   (k' = tokens)
// Replaces the cuBLAS calls behind nn.Linear in HF ViT / BertGeneration (transformers modeling_vit.py:228-230,
// 265-268, 296-312; modeling_bert_generation.py:52-56, 89-153, 265-293) used by
// vilmedic/blocks/vision/visual_encoder.py:56-58 and vilmedic/blocks/huggingface/decoder/decoder_model.py:23-26.
//
// Warp roles (320 threads): warp0 = TMA producer, warp1 = TMEM allocator + MMA issuer, warps2..9 = epilogue
// (warp w drains TMEM lane quadrant w%4, column slice (w-2)/4).
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "vlm_b200.h"
namespace vlm {

static constexpr int GEMM_BM = 128;
static constexpr int GEMM_BK = 64;  // 64 bf16 = 128 B = one swizzle row

// kernel + launchers live in gemm_kernel.cuh and are instantiated per tile width in gemm_tcgen05_bn{64,128,192,256}.cu
template <int BN>
int gemm_launch_bn(int a_mn_major, int b_mn_major, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                   const CUtensorMap& tx, int mode, int M, int N, int K, int batch, int a_bmul, int b_bmul, int split_k, long long c_bs,
                   long long aux_bs, long long res_bs, const GemmEpilogue& epi, int max_ctas, cudaStream_t stream);
extern template int gemm_launch_bn<64>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, int, int, int,
                                       int, int, int, int, int, long long, long long, long long, const GemmEpilogue&, int, cudaStream_t);
extern template int gemm_launch_bn<128>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, int, int, int,
                                        int, int, int, int, int, long long, long long, long long, const GemmEpilogue&, int, cudaStream_t);
extern template int gemm_launch_bn<192>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, int, int, int,
                                        int, int, int, int, int, long long, long long, long long, const GemmEpilogue&, int, cudaStream_t);
extern template int gemm_launch_bn<256>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, int, int, int,
                                        int, int, int, int, int, long long, long long, long long, const GemmEpilogue&, int, cudaStream_t);

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(err));
      return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 3-D bf16 tensor map: dims {inner, rows, batch}; box {64, box_rows, 1}; 128B swizzle; OOB -> zero.
int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t rows, uint64_t batch,
                   uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -1;
  cuuint64_t dims[3] = {inner, rows, batch};
  cuuint64_t strides[2] = {row_stride_elems * 2, (batch > 1 ? batch_stride_elems : row_stride_elems * rows) * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu rows=%llu batch=%llu ld=%llu bstride=%llu box=%u",
              (int)r, ptr, (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)batch,
              (unsigned long long)row_stride_elems, (unsigned long long)batch_stride_elems, box_rows);
    return -1;
  }
  return 0;
}

// Generic bf16 tiled tensor map (rank <= 5), 128B swizzle, zero OOB fill; strides in ELEMENTS for dims 1..rank-1.
int make_tmap_bf16_nd(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                      const uint32_t* box, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -1;
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) st[i - 1] = strides_elems[i - 1] * 2;
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                               : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), d, st, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(rank %d) failed (%d): ptr=%p dims=%llu,%llu,%llu,%llu", rank, (int)r, ptr,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0));
    return -1;
  }
  return 0;
}

// make sure the calling host thread has the primary context bound (cuTensorMapEncodeTiled is a driver call)
void bind_context_for_driver_calls() {
  static thread_local int bound_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (bound_dev != dev) {
    cudaSetDevice(dev);
    cudaFree(0);
    bound_dev = dev;
  }
}

static int pick_bn(int M, int N, int batch, int force_bn) {
  if (force_bn == 64 || force_bn == 128 || force_bn == 192 || force_bn == 256) return force_bn;
  const int sms = num_sms();
  const long long m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  int best = 128;
  double best_cost = 1e30;
  const int cands[4] = {256, 192, 128, 64};
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (bn > 64 && N <= bn / 2) continue;  // mostly padding
    const long long tiles = m_tiles * ((N + bn - 1) / bn) * batch;
    const long long waves = (tiles + sms - 1) / sms;
    // per-tile cost ~ BN (MMA time) + a fixed part (epilogue drain / pipeline fill)
    const double cost = (double)waves * (bn + 24.0);
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

int gemm2_dispatch(const void* a, long long lda, int a_mn, const void* b, long long ldb, int b_mn, int M, int N, int K,
                   int bn, const GemmEpilogue& e, const CUtensorMap* tc, const CUtensorMap* tx, int mode, cudaStream_t s);

// 0 = 1-CTA kernel, else the N tile (128 | 256) of the 2-CTA kernel.  force_bn >= 1000 forces 2-CTA with bn = force_bn-1000.
// Default rule (measured, gpurun_out r2b gemm bench): the pair kernel wins where the 1-CTA kernel is bound by L2 -> smem operand
// traffic and the epilogue is light — wide bias-only forward products (QKV 12608x2304x768: 40 vs 43 us, LM head 8192x30522x768:
// 305 vs 340 us); with GELU / residual epilogues or few output tiles it loses to wave quantisation (256-row tiles on 74 pairs).
// VLM_GEMM_2CTA=1 applies the cost model below to every shape, =0 disables the kernel.
static int pick_2cta(int M, int N, int K, int batch, int force_bn, int bn1, bool light_forward) {
  if (batch != 1) return 0;
  if (force_bn >= 1000) return force_bn - 1000;
  if (force_bn != 0) return 0;
  static int enabled = -1;
  if (enabled < 0) {
    const char* env = getenv("VLM_GEMM_2CTA");
    enabled = env ? ((env[0] == '1') ? 1 : 0) : 2;
  }
  if (enabled == 2) return (light_forward && M >= 2048 && N >= 2048 && K <= 1536) ? 256 : 0;
  if (!enabled || M < 512 || N < 128) return 0;
  const int sms = num_sms();
  const long long m1 = (M + GEMM_BM - 1) / GEMM_BM, m2 = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const long long t1 = m1 * ((N + bn1 - 1) / bn1);
  const double cost1 = (double)((t1 + sms - 1) / sms) * (bn1 + 24.0);
  int best = 0;
  double best_cost = cost1;
  const int cands[2] = {256, 128};
  for (int i = 0; i < 2; ++i) {
    const int bn = cands[i];
    if (N <= bn / 2) continue;
    const long long t2 = m2 * ((N + bn - 1) / bn);
    const long long pairs = sms / 2;
    // a pair finishes a 256 x bn tile in the time a single CTA needs for 128 x bn, with 1.5x less L2->smem traffic
    const double cost2 = (double)((t2 + pairs - 1) / pairs) * (bn + 24.0) * 0.85;
    if (cost2 < best_cost) {
      best_cost = cost2;
      best = bn;
    }
  }
  return best;
}

}  // namespace vlm

using namespace vlm;

extern "C" int vlm_gemm_bf16(const void* a, long long lda, int a_mn_major, const void* b, long long ldb,
                             int b_mn_major, void* c, long long ldc, int c_is_fp32, int M, int N, int K,
                             const float* bias, const void* residual, long long ldr, int act, const void* aux_in,
                             void* aux_out, long long ld_aux, float alpha, const float* alpha_ptr, int accumulate,
                             int batch,
                             long long a_batch_stride, long long b_batch_stride, long long c_batch_stride,
                             long long aux_batch_stride, long long res_batch_stride, float p_drop,
                             unsigned long long seed, unsigned long long offset,
                             const unsigned long long* rng_offset_ptr, int force_bn, int max_ctas, void* stream) {
  VLM_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, "vlm_gemm_bf16: bad shape M=%d N=%d K=%d batch=%d", M, N, K, batch);
  VLM_REQUIRE(a && b && c, "vlm_gemm_bf16: null operand");
  VLM_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "vlm_gemm_bf16: lda/ldb must be multiples of 8 elements (TMA 16B stride)");
  VLM_REQUIRE(((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)c % 16 == 0),
              "vlm_gemm_bf16: operands must be 16B aligned");
  VLM_REQUIRE(ldc % (c_is_fp32 ? 4 : 8) == 0, "vlm_gemm_bf16: ldc must keep rows 16B aligned");
  VLM_REQUIRE(act >= 0 && act <= 2, "vlm_gemm_bf16: act must be 0|1|2");
  VLM_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || (N % 4 == 0 && batch == 1)),
              "vlm_gemm_bf16: dropout needs 0<=p<1, N %% 4 == 0, batch == 1");
  VLM_REQUIRE(act != 2 || aux_in, "vlm_gemm_bf16: act=2 needs aux_in");
  VLM_REQUIRE(!residual || ldr % (c_is_fp32 ? 4 : 8) == 0, "vlm_gemm_bf16: ldr alignment");
  VLM_REQUIRE(!(aux_in || aux_out) || ld_aux % 8 == 0, "vlm_gemm_bf16: ld_aux alignment");
  VLM_REQUIRE(!bias || ((uintptr_t)bias % 16 == 0), "vlm_gemm_bf16: bias must be 16B aligned");
  VLM_REQUIRE(batch == 1 || (a_batch_stride % 8 == 0 && b_batch_stride % 8 == 0), "vlm_gemm_bf16: batch strides % 8");

  // cuTensorMapEncodeTiled is a driver call: make sure this host thread (e.g. an autograd worker) has the primary
  // context bound before the first one, otherwise it fails with CUDA_ERROR_INVALID_CONTEXT.
  bind_context_for_driver_calls();
  // split-K candidates (weight gradients: fp32 C accumulated in place, K = number of tokens) prefer wide N tiles and
  // fill the machine along K instead of shrinking the tile.
  const bool splitk_ok = c_is_fp32 && accumulate && act == 0 && p_drop == 0.f && !residual && batch == 1 && K >= 1024;
  int bn = pick_bn(M, N, batch, force_bn >= 1000 ? 0 : force_bn);
  if (splitk_ok && force_bn == 0) {
    const int wide = N >= 192 ? 256 : (N >= 96 ? 128 : 64);
    const long long tiles_wide = (long long)((M + GEMM_BM - 1) / GEMM_BM) * ((N + wide - 1) / wide);
    if (tiles_wide < num_sms()) bn = wide;
  }
  const int bn2 = pick_2cta(M, N, K, batch, force_bn, bn, !a_mn_major && !b_mn_major && !c_is_fp32 && !accumulate && act == 0 &&
                                                          !residual && p_drop == 0.f);
  VLM_REQUIRE(bn2 == 0 || bn2 == 128 || bn2 == 256, "vlm_gemm_bf16: 2-CTA N tile must be 128 or 256");
  CUtensorMap ta, tb;
  // a zero batch stride broadcasts that operand: encode a single-batch map and pin the batch coordinate to 0
  const int a_bmul = (batch > 1 && a_batch_stride != 0) ? 1 : 0, b_bmul = (batch > 1 && b_batch_stride != 0) ? 1 : 0;
  const int a_nb = a_bmul ? batch : 1, b_nb = b_bmul ? batch : 1;
  // A: K-major -> tensor {K, M}, box {64, 128}; MN-major -> tensor {M, K}, box {64, 64}
  if (a_mn_major) {
    if (make_tmap_bf16(&ta, a, (uint64_t)M, (uint64_t)K, a_nb, lda, a_batch_stride, GEMM_BK)) return -1;
  } else {
    if (make_tmap_bf16(&ta, a, (uint64_t)K, (uint64_t)M, a_nb, lda, a_batch_stride, GEMM_BM)) return -1;
  }
  if (b_mn_major) {
    if (make_tmap_bf16(&tb, b, (uint64_t)N, (uint64_t)K, b_nb, ldb, b_batch_stride, GEMM_BK)) return -1;
  } else {
    if (make_tmap_bf16(&tb, b, (uint64_t)K, (uint64_t)N, b_nb, ldb, b_batch_stride, bn)) return -1;
  }
  GemmEpilogue e;
  e.c = c;
  e.ldc = ldc;
  e.c_fp32 = c_is_fp32;
  e.bias = bias;
  e.residual = residual;
  e.ldr = ldr;
  e.act = act;
  e.aux_in = reinterpret_cast<const bf16*>(aux_in);
  e.aux_out = reinterpret_cast<bf16*>(aux_out);
  e.ld_aux = ld_aux;
  e.alpha = alpha;
  e.alpha_ptr = alpha_ptr;
  e.accumulate = accumulate;
  e.p_drop = p_drop;
  e.seed = seed;
  e.offset = offset;
  e.offset_ptr = rng_offset_ptr;
  e.drop_ld = N;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // split-K for under-filled grids (weight gradients: few output tiles, K = number of tokens): fp32 C, accumulate
  // semantics, atomics in the epilogue.
  int split_k = 1;
  e.atomic = 0;
  if (bn2 == 0 && splitk_ok) {
    const long long tiles = (long long)((M + GEMM_BM - 1) / GEMM_BM) * ((N + bn - 1) / bn);
    const int k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
    const int sms = num_sms();
    if (tiles * 2 <= sms) {
      split_k = (int)(sms / tiles);
      if (split_k > k_blocks / 4) split_k = k_blocks / 4;
      if (split_k < 1) split_k = 1;
    }
    e.atomic = split_k > 1;
  }
  // C tensor map for the TMA-store epilogue: bf16 C, plain store (no accumulate / split-K atomics), 16-byte aligned pitch.
  CUtensorMap tc = ta, tx = ta;
  int tma_store = 0;
  static const bool tma_store_enabled = [] {   // debug knob: VLM_GEMM_TMA_STORE=0 keeps the direct-store epilogue
    const char* v = getenv("VLM_GEMM_TMA_STORE");
    return !(v && v[0] == '0');
  }();
  constexpr bool pair_staged = true;    // the CTA-pair kernel has the staged epilogue too
  if (tma_store_enabled && (bn2 == 0 || pair_staged) && !c_is_fp32 && !accumulate && !e.atomic && (ldc % 8) == 0 && (reinterpret_cast<uintptr_t>(c) & 15) == 0 &&
      (batch == 1 || (c_batch_stride % 8) == 0)) {
    const uint64_t dims[3] = {(uint64_t)N, (uint64_t)M, (uint64_t)batch};
    const uint64_t strides[2] = {(uint64_t)ldc, (uint64_t)(batch > 1 ? c_batch_stride : ldc * (long long)M)};
    const uint32_t box[3] = {32, 32, 1};
    if (make_tmap_bf16_nd(&tc, c, 3, dims, strides, box, 64)) return -1;
    tma_store = 1;
    if (act == 1 && aux_out) {               // GELU stash leaves through the same staging tile
      if ((ld_aux % 8) == 0 && (reinterpret_cast<uintptr_t>(aux_out) & 15) == 0 && (batch == 1 || (aux_batch_stride % 8) == 0)) {
        const uint64_t xstrides[2] = {(uint64_t)ld_aux, (uint64_t)(batch > 1 ? aux_batch_stride : ld_aux * (long long)M)};
        if (make_tmap_bf16_nd(&tx, aux_out, 3, dims, xstrides, box, 64)) return -1;
      } else {
        tma_store = 0;
      }
    }
    // the fast path reads the GELU' argument / residual rows with 16-byte loads
    if (act == 2 && ((ld_aux % 8) != 0 || (reinterpret_cast<uintptr_t>(aux_in) & 15) != 0)) tma_store = 0;
    if (residual && ((ldr % 8) != 0 || (reinterpret_cast<uintptr_t>(residual) & 15) != 0)) tma_store = 0;
    if (bias && (reinterpret_cast<uintptr_t>(bias) & 15) != 0) tma_store = 0;
    if (act == 2 && (residual || p_drop > 0.f)) tma_store = 0;          // combinations without a specialised kernel
    if (act == 1 && (residual || p_drop > 0.f)) tma_store = 0;
    if (tma_store) tma_store = act == 1 ? EPI_GELU : (act == 2 ? EPI_GELUGRAD : ((residual || p_drop > 0.f) ? EPI_RESID : EPI_BIAS));
  }
  if (bn2 != 0)   // CTA pair; the C / aux tensor maps + mode feed its staged epilogue
    return gemm2_dispatch(a, lda, a_mn_major, b, ldb, b_mn_major, M, N, K, bn2, e, tma_store ? &tc : nullptr, tma_store ? &tx : nullptr,
                          tma_store, s);

  switch (bn) {
    case 64:
      return gemm_launch_bn<64>(a_mn_major, b_mn_major, ta, tb, tc, tx, tma_store, M, N, K, batch, a_bmul, b_bmul, split_k, c_batch_stride,
                                aux_batch_stride, res_batch_stride, e, max_ctas, s);
    case 128:
      return gemm_launch_bn<128>(a_mn_major, b_mn_major, ta, tb, tc, tx, tma_store, M, N, K, batch, a_bmul, b_bmul, split_k, c_batch_stride,
                                 aux_batch_stride, res_batch_stride, e, max_ctas, s);
    case 192:
      return gemm_launch_bn<192>(a_mn_major, b_mn_major, ta, tb, tc, tx, tma_store, M, N, K, batch, a_bmul, b_bmul, split_k, c_batch_stride,
                                 aux_batch_stride, res_batch_stride, e, max_ctas, s);
    default:
      return gemm_launch_bn<256>(a_mn_major, b_mn_major, ta, tb, tc, tx, tma_store, M, N, K, batch, a_bmul, b_bmul, split_k, c_batch_stride,
                                 aux_batch_stride, res_batch_stride, e, max_ctas, s);
  }
}
