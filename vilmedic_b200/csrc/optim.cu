// Fused flat-buffer optimizer step (the step either side of the hot path,
// vilmedic/executors/trainor.py:119-124 unscale+clip+step+zero_grad; optimizer chosen by name in
// vilmedic/executors/utils.py:81-86).  One pass over the flat fp32 master parameters:
//   g' = g * grad_scale * clip_coef ; AdamW update of (p, m, v) ; bf16 mirror of p written in the same pass ;
//   optional zeroing of g.  HBM-bound: 16 B read + 14 B written per parameter.
// `step`, `lr_scale` and the squared grad-norm live in device memory so the launch is CUDA-graph replayable.
#include "common.cuh"
#include "vlm_b200.h"

namespace vlm {

__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

__global__ void adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             bf16* __restrict__ p_bf16, long long n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, const int* __restrict__ step_ptr, const float* __restrict__ lr_scale_ptr,
                             float grad_scale, const float* __restrict__ gnorm_sq_ptr, float max_norm, int zero_grad) {
  const int step = step_ptr ? *step_ptr : 1;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  float lr_eff = lr * (lr_scale_ptr ? *lr_scale_ptr : 1.f);
  float gs = grad_scale;
  if (gnorm_sq_ptr && max_norm > 0.f) {
    const float norm = sqrtf(*gnorm_sq_ptr) * grad_scale;
    const float coef = max_norm / (norm + 1e-6f);
    if (coef < 1.f) gs *= coef;
  }
  const float step_size = lr_eff / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gp[j] * gs;
      mp[j] = beta1 * mp[j] + (1.f - beta1) * gj;
      vp[j] = beta2 * vp[j] + (1.f - beta2) * gj * gj;
      const float denom = sqrtf(vp[j]) * inv_sqrt_bc2 + eps;
      pp[j] = pp[j] * (1.f - lr_eff * weight_decay) - step_size * mp[j] / denom;
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p_bf16) {
      uint2 u;
      u.x = pack_bf16x2(pv.x, pv.y);
      u.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(p_bf16)[i] = u;
    }
  }
}

__global__ void step_inc_kernel(int* step) { *step += 1; }

}  // namespace vlm

using namespace vlm;

extern "C" int vlm_sumsq_f32(const float* g, long long n, float* out, void* stream) {
  VLM_REQUIRE(g && out && n > 0, "vlm_sumsq_f32: bad args");
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, n, out);
  return check_launch("sumsq");
}

extern "C" int vlm_adamw_step(float* p, float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int* step_ptr, int increment_step,
                              const float* lr_scale_ptr, float grad_scale, const float* gnorm_sq_ptr, float max_norm,
                              int zero_grad, void* stream) {
  VLM_REQUIRE(p && g && m && v && n > 0 && n % 4 == 0, "vlm_adamw_step: flat buffers must be non-null with n %% 4 == 0");
  cudaStream_t s = (cudaStream_t)stream;
  if (step_ptr && increment_step) step_inc_kernel<<<1, 1, 0, s>>>(step_ptr);
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<(int)blocks, 256, 0, s>>>(p, g, m, v, (bf16*)p_bf16, n, lr, beta1, beta2, eps, weight_decay, step_ptr,
                                           lr_scale_ptr, grad_scale, gnorm_sq_ptr, max_norm, zero_grad);
  return check_launch("adamw_step");
}
