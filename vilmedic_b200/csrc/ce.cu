// Fused softmax cross-entropy (forward loss + gradient w.r.t. logits in one pass over the row).
// One CTA per row; the row is staged once in shared memory (V*4 bytes as fp32), so HBM sees exactly one read of the
// logits and one write of dlogits.  Supports: next-token labels derived in-kernel from input_ids (labels=input_ids
// shifted left, last position ignored — HF loss_utils.py:45-66 ForCausalLMLoss, called because
// vilmedic/blocks/huggingface/decoder/decoder_model.py:46 passes labels=input_ids), explicit labels with
// ignore_index=-100, and label smoothing (vilmedic/blocks/losses/mvqa/LabelSmoothingCrossEntropyLoss.py:38-48).
#include "common.cuh"
#include "vlm_b200.h"

namespace vlm {

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float r = (lane < nw) ? red[lane] : (is_max ? -INFINITY : 0.f);
  r = is_max ? warp_max(r) : warp_sum(r);
  return r;
}

// TL: logits type (bf16 or float).  dlogits has the same type and may alias logits.  The row cache keeps TL.
template <typename TL>
struct CeVec {
  static constexpr int N = 16 / sizeof(TL);
  static __device__ __forceinline__ void unpack(const uint4& u, float* f) {
    if constexpr (sizeof(TL) == 4) {
      f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    } else {
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
    }
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    uint4 u;
    if constexpr (sizeof(TL) == 4) {
      u.x = __float_as_uint(f[0]); u.y = __float_as_uint(f[1]); u.z = __float_as_uint(f[2]); u.w = __float_as_uint(f[3]);
    } else {
      u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    }
    return u;
  }
  static __device__ __forceinline__ float to_f(TL x) {
    if constexpr (sizeof(TL) == 4) return x; else return __bfloat162float(x);
  }
  static __device__ __forceinline__ TL from_f(float x) {
    if constexpr (sizeof(TL) == 4) return x; else return __float2bfloat16(x);
  }
};

template <typename TL>
__global__ void __launch_bounds__(512) softmax_ce_kernel(const TL* __restrict__ logits, long long ld, const long long* __restrict__ ids,
                                                         int shift_T, int V, float smoothing, float grad_scale,
                                                         TL* dlogits, long long ldd, float* __restrict__ loss_rows,
                                                         float* __restrict__ lse_rows, const float* __restrict__ row_weight) {
  using CV = CeVec<TL>;
  constexpr int VN = CV::N;
  extern __shared__ uint4 row4[];  // row cache, TL[V] (16B aligned)
  TL* row = reinterpret_cast<TL*>(row4);
  __shared__ float red[32];
  const int r = blockIdx.x;
  const TL* lr = logits + (size_t)r * ld;
  long long label;
  if (shift_T > 0) {
    const int t = r % shift_T;
    label = (t < shift_T - 1) ? ids[r + 1] : -100;
  } else {
    label = ids[r];
  }
  const bool ignore = (label < 0 || label >= V);
  const int nv = V / VN;
  if (row_weight) grad_scale *= __ldg(row_weight + r);     // per-row weight of the gradient (SCST: reward-weighted log-likelihood)

  float mx = -INFINITY;
  for (int v = threadIdx.x; v < nv; v += blockDim.x) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(lr) + v);
    row4[v] = u;
    float f[VN];
    CV::unpack(u, f);
#pragma unroll
    for (int j = 0; j < VN; ++j) mx = fmaxf(mx, f[j]);
  }
  for (int i = nv * VN + threadIdx.x; i < V; i += blockDim.x) {
    const TL x = lr[i];
    row[i] = x;
    mx = fmaxf(mx, CV::to_f(x));
  }
  mx = block_reduce(mx, red, true);  // includes the __syncthreads that publishes the row cache
  float se = 0.f, sx = 0.f;
  for (int v = threadIdx.x; v < nv; v += blockDim.x) {
    float f[VN];
    CV::unpack(row4[v], f);
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      se += __expf(f[j] - mx);
      sx += f[j];
    }
  }
  for (int i = nv * VN + threadIdx.x; i < V; i += blockDim.x) {
    const float x = CV::to_f(row[i]);
    se += __expf(x - mx);
    sx += x;
  }
  se = block_reduce(se, red, false);
  const float lse = mx + logf(se);
  float loss = 0.f;
  if (smoothing > 0.f) sx = block_reduce(sx, red, false);
  if (!ignore) {
    const float nll = lse - CV::to_f(row[label]);
    loss = nll;
    if (smoothing > 0.f) {
      const float neg_sum_logp = (float)V * lse - sx;  // -sum_j log p_j
      loss = (1.f - smoothing) * nll + smoothing / (float)V * neg_sum_logp;
    }
  }
  if (threadIdx.x == 0) {
    if (loss_rows) loss_rows[r] = loss;
    if (lse_rows) lse_rows[r] = lse;
  }
  if (dlogits) {
    TL* dr = dlogits + (size_t)r * ldd;
    const float uni = smoothing / (float)V;
    const float hot = 1.f - smoothing;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
      float f[VN];
      CV::unpack(row4[v], f);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        const int i = v * VN + j;
        f[j] = ignore ? 0.f : (__expf(f[j] - lse) - uni - ((i == label) ? hot : 0.f)) * grad_scale;
      }
      reinterpret_cast<uint4*>(dr)[v] = CV::pack(f);
    }
    for (int i = nv * VN + threadIdx.x; i < (int)ldd; i += blockDim.x) {
      float g = 0.f;
      if (!ignore && i < V) g = (__expf(CV::to_f(row[i]) - lse) - uni - ((i == label) ? hot : 0.f)) * grad_scale;
      dr[i] = CV::from_f(g);
    }
  }
}

}  // namespace vlm

using namespace vlm;

extern "C" int vlm_softmax_ce(const void* logits, int logits_fp32, long long ld, const long long* ids, int shift_T, int R,
                              int V, float smoothing, float grad_scale, void* dlogits, long long ldd, float* loss_rows,
                              float* lse_rows, const float* row_weight, void* stream) {
  VLM_REQUIRE(logits && ids && R > 0 && V > 0 && ld >= V, "vlm_softmax_ce: bad args (R=%d V=%d ld=%lld)", R, V, ld);
  VLM_REQUIRE(ld % (logits_fp32 ? 4 : 8) == 0 && (!dlogits || ldd % (logits_fp32 ? 4 : 8) == 0), "vlm_softmax_ce: rows must be 16B aligned");
  VLM_REQUIRE(!dlogits || ldd >= V, "vlm_softmax_ce: ldd < V");
  VLM_REQUIRE(smoothing >= 0.f && smoothing < 1.f, "vlm_softmax_ce: smoothing out of range");
  const size_t smem = (((size_t)V * (logits_fp32 ? 4 : 2)) + 15) / 16 * 16;
  VLM_REQUIRE(smem <= 200 * 1024, "vlm_softmax_ce: V=%d too large for the single-pass row cache", V);
  const int threads = V >= 8192 ? 512 : (V >= 1024 ? 256 : 128);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (logits_fp32) {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(softmax_ce_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set = true; }
    softmax_ce_kernel<float><<<R, threads, smem, s>>>((const float*)logits, ld, ids, shift_T, V, smoothing, grad_scale,
                                                      (float*)dlogits, ldd, loss_rows, lse_rows, row_weight);
  } else {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(softmax_ce_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set = true; }
    softmax_ce_kernel<bf16><<<R, threads, smem, s>>>((const bf16*)logits, ld, ids, shift_T, V, smoothing, grad_scale,
                                                     (bf16*)dlogits, ldd, loss_rows, lse_rows, row_weight);
  }
  return check_launch("softmax_ce");
}
