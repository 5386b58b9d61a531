// Image x text contrastive losses (ConVIRT, InfoNCE, GLoRIA-global) around the tcgen05 similarity GEMM.
//   reference: vilmedic/blocks/losses/selfsup/ConVIRTLoss.py:12-31, InfoNCELoss.py:11-19, GLoRIALoss.py:54-75.
// All three are "symmetric log-sum-exp" losses on S = scale * A B^T (N x N):
//     loss_row[i] = LSE_j(S_ij) - S_ii      loss_col[i] = LSE_j(S_ji) - S_ii
//   ConVIRT : rows/cols L2-normalised, scale 1/tau,  loss = mean(lambda*loss_col + (1-lambda)*loss_row)
//             (the reference forms exp(S/tau) without max subtraction; |S/tau| <= 10 so both are exact in fp32)
//   InfoNCE : raw dot products (tau unused in the reference), loss = mean((loss_row + loss_col)/2)
//   GLoRIA  : cosine * temp3, loss0 = mean(loss_row), loss1 = mean(loss_col)
// Pipeline: rownorm_split (fp32 -> L2-normalised, split into bf16 hi/lo so that the tensor-core product
// hi*hi + hi*lo + lo*hi carries ~16 mantissa bits) -> vlm_gemm_bf16 (K = 3D, fp32 out) -> sym_lse -> sym_lse_bwd ->
// two GEMMs -> rownorm_bwd.  The N x N matrix is <= 1 MB for N = 512 and stays L2 resident.
#include "common.cuh"
#include "vlm_b200.h"

namespace vlm {

// x fp32 [N,D] -> xa bf16 [N,3D] = [hi|hi|lo], xb bf16 [N,3D] = [hi|lo|hi], xh bf16 [N,D] = hi, inv_norm[N]
// normalize != 0: x <- x / max(||x||, eps).
__global__ void rownorm_split_kernel(const float* __restrict__ x, bf16* __restrict__ xa, bf16* __restrict__ xb,
                                     bf16* __restrict__ xh, float* __restrict__ inv_norm, int N, int D, int normalize, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const float* xr = x + (size_t)warp * D;
  float inv = 1.f;
  if (normalize) {
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s += xr[i] * xr[i];
    s = warp_sum(s);
    inv = 1.f / fmaxf(sqrtf(s), eps);
  }
  if (lane == 0 && inv_norm) inv_norm[warp] = inv;
  for (int i = lane; i < D; i += 32) {
    const float v = xr[i] * inv;
    const bf16 hi = __float2bfloat16(v);
    const bf16 lo = __float2bfloat16(v - __bfloat162float(hi));
    if (xa) { xa[(size_t)warp * 3 * D + i] = hi; xa[(size_t)warp * 3 * D + D + i] = hi; xa[(size_t)warp * 3 * D + 2 * D + i] = lo; }
    if (xb) { xb[(size_t)warp * 3 * D + i] = hi; xb[(size_t)warp * 3 * D + D + i] = lo; xb[(size_t)warp * 3 * D + 2 * D + i] = hi; }
    if (xh) xh[(size_t)warp * D + i] = hi;
  }
}

// dx = inv_norm * (dxh - xhat * (xhat . dxh)) when normalize, else dx = dxh.   xhat recomputed from x.
__global__ void rownorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ inv_norm, const float* __restrict__ dxh,
                                   float* __restrict__ dx, int N, int D, int normalize) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const float* xr = x + (size_t)warp * D;
  const float* gr = dxh + (size_t)warp * D;
  float* dr = dx + (size_t)warp * D;
  if (!normalize) {
    for (int i = lane; i < D; i += 32) dr[i] = gr[i];
    return;
  }
  const float inv = inv_norm[warp];
  float dot = 0.f;
  for (int i = lane; i < D; i += 32) dot += xr[i] * inv * gr[i];
  dot = warp_sum(dot);
  for (int i = lane; i < D; i += 32) dr[i] = inv * (gr[i] - xr[i] * inv * dot);
}

// One warp per index i: row LSE (contiguous) and column LSE (strided) of scale*S, plus the diagonal.
__global__ void sym_lse_kernel(const float* __restrict__ S, int N, long long ld, float scale, float* __restrict__ lse_row,
                               float* __restrict__ lse_col, float* __restrict__ loss_row, float* __restrict__ loss_col) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  float mr = -INFINITY, mc = -INFINITY;
  for (int j = lane; j < N; j += 32) {
    mr = fmaxf(mr, S[(size_t)i * ld + j] * scale);
    mc = fmaxf(mc, S[(size_t)j * ld + i] * scale);
  }
  mr = warp_max(mr);
  mc = warp_max(mc);
  float sr = 0.f, sc = 0.f;
  for (int j = lane; j < N; j += 32) {
    sr += expf(S[(size_t)i * ld + j] * scale - mr);
    sc += expf(S[(size_t)j * ld + i] * scale - mc);
  }
  sr = warp_sum(sr);
  sc = warp_sum(sc);
  if (lane == 0) {
    const float lr = mr + logf(sr), lc = mc + logf(sc), d = S[(size_t)i * ld + i] * scale;
    lse_row[i] = lr;
    lse_col[i] = lc;
    loss_row[i] = lr - d;
    loss_col[i] = lc - d;
  }
}

// dS_ij = g * scale * [ w_row * (exp(s_ij - lse_row[i]) - d_ij) + w_col * (exp(s_ij - lse_col[j]) - d_ij) ]   (bf16)
// w_row / w_col already include the 1/N of the mean.
__global__ void sym_lse_bwd_kernel(const float* __restrict__ S, int N, long long ld, float scale, const float* __restrict__ lse_row,
                                   const float* __restrict__ lse_col, float w_row, float w_col, const float* __restrict__ g_ptr,
                                   bf16* __restrict__ dS, long long ldd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= ldd) return;
  float v = 0.f;
  if (j < N) {
    const float s = S[(size_t)i * ld + j] * scale;
    const float d = (i == j) ? 1.f : 0.f;
    v = (g_ptr ? *g_ptr : 1.f) * scale * (w_row * (expf(s - lse_row[i]) - d) + w_col * (expf(s - lse_col[j]) - d));
  }
  dS[(size_t)i * ldd + j] = __float2bfloat16(v);
}

}  // namespace vlm

using namespace vlm;

extern "C" int vlm_rownorm_split(const float* x, void* xa, void* xb, void* xh, float* inv_norm, int N, int D, int normalize,
                                 float eps, void* stream) {
  VLM_REQUIRE(x && N > 0 && D > 0, "vlm_rownorm_split: bad args");
  rownorm_split_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, (bf16*)xa, (bf16*)xb, (bf16*)xh, inv_norm, N, D, normalize, eps);
  return check_launch("rownorm_split");
}

extern "C" int vlm_rownorm_bwd(const float* x, const float* inv_norm, const float* dxh, float* dx, int N, int D, int normalize,
                               void* stream) {
  VLM_REQUIRE(x && dxh && dx && N > 0 && D > 0 && (!normalize || inv_norm), "vlm_rownorm_bwd: bad args");
  rownorm_bwd_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, inv_norm, dxh, dx, N, D, normalize);
  return check_launch("rownorm_bwd");
}

extern "C" int vlm_sym_lse(const float* S, int N, long long ld, float scale, float* lse_row, float* lse_col, float* loss_row,
                           float* loss_col, void* stream) {
  VLM_REQUIRE(S && lse_row && lse_col && loss_row && loss_col && N > 0 && ld >= N, "vlm_sym_lse: bad args");
  sym_lse_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(S, N, ld, scale, lse_row, lse_col, loss_row, loss_col);
  return check_launch("sym_lse");
}

extern "C" int vlm_sym_lse_bwd(const float* S, int N, long long ld, float scale, const float* lse_row, const float* lse_col,
                               float w_row, float w_col, const float* g_ptr, void* dS, long long ldd, void* stream) {
  VLM_REQUIRE(S && lse_row && lse_col && dS && N > 0 && ld >= N && ldd >= N && ldd % 8 == 0, "vlm_sym_lse_bwd: bad args");
  dim3 grid((unsigned)((ldd + 255) / 256), N);
  sym_lse_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(S, N, ld, scale, lse_row, lse_col, w_row, w_col, g_ptr, (bf16*)dS, ldd);
  return check_launch("sym_lse_bwd");
}
