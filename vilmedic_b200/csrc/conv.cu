// CNN backbone pieces (torchvision ResNet-18/50 of vilmedic/blocks/vision/visual_encoder.py:71-83):
// convolutions run as GEMMs on the tcgen05 kernel (gemm_tcgen05.cu) over NHWC bf16 activations; this file holds the
// HBM-bound kernels around them — im2col / col2im gathers, weight (un)packing between torchvision's OIHW fp32 masters and the
// [Cout, kh*kw*Cin] bf16 GEMM operand, training-mode BatchNorm (column statistics, normalise + residual + ReLU, backward),
// 3x3/2 max-pool with recorded arg-max, global average pool.  Activations are [B*H*W, C] row-major (C % 8 == 0), so every
// access below is a 16-byte vector of 8 channels; algorithmic bytes = each tensor touched once per kernel.
#include "common.cuh"
#include "vlm_b200.h"

namespace vlm {

static inline int conv_grid(long long work, int block) {
  long long g = (work + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

// ---------------------------------------------------------------- weights: OIHW fp32 <-> [Cout, Kp] (kh, kw, ci) bf16 / fp32
__global__ void conv_weight_pack_kernel(const float* __restrict__ w, bf16* __restrict__ wm, int Cout, int Cin, int KH, int KW, int Kp) {
  const long long n = (long long)Cout * Kp;
  const int K = KH * KW * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / Kp), k = (int)(i % Kp);
    float v = 0.f;
    if (k < K) {
      const int ci = k % Cin, kw = (k / Cin) % KW, kh = k / (Cin * KW);
      v = w[(((long long)co * Cin + ci) * KH + kh) * KW + kw];
    }
    wm[i] = __float2bfloat16(v);
  }
}
__global__ void conv_wgrad_unpack_kernel(const float* __restrict__ dwm, float* __restrict__ gw, int Cout, int Cin, int KH, int KW, int Kp) {
  const long long n = (long long)Cout * Cin * KH * KW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int kw = (int)(i % KW), kh = (int)((i / KW) % KH), ci = (int)((i / ((long long)KW * KH)) % Cin);
    const int co = (int)(i / ((long long)KW * KH * Cin));
    gw[i] += dwm[(long long)co * Kp + (kh * KW + kw) * Cin + ci];
  }
}

// ---------------------------------------------------------------- im2col / col2im
// col[m, (kh*KW + kw)*C + c] = x[b, oy*s - p + kh, ox*s - p + kw, c]  (0 outside), m = (b*Ho + oy)*Wo + ox
__global__ void im2col_nhwc_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int B, int H, int W, int C, int KH, int KW,
                                   int stride, int pad, int Ho, int Wo) {
  const int C8 = C >> 3;
  const long long n = (long long)B * Ho * Wo * KH * KW * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long t = i / C8;
    const int kw = (int)(t % KW); t /= KW;
    const int kh = (int)(t % KH); t /= KH;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const int iy = oy * stride - pad + kh, ix = ox * stride - pad + kw;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = __ldg(reinterpret_cast<const uint4*>(x + (((long long)b * H + iy) * W + ix) * C) + c8);
    reinterpret_cast<uint4*>(col)[i] = v;       // i enumerates col in memory order: (m, kh, kw, c8)
  }
}
// stem: fp32 NCHW images -> col[m, (kh*KW + kw)*Cin + ci] bf16, row pitch Kp (zero padded)
__global__ void im2col_nchw_f32_kernel(const float* __restrict__ img, bf16* __restrict__ col, int B, int Cin, int H, int W, int KH, int KW,
                                       int stride, int pad, int Ho, int Wo, int Kp) {
  const long long n = (long long)B * Ho * Wo * Kp;
  const int K = KH * KW * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    long long m = i / Kp;
    float v = 0.f;
    if (k < K) {
      const int ci = k % Cin, kw = (k / Cin) % KW, kh = k / (Cin * KW);
      const int ox = (int)(m % Wo), oy = (int)((m / Wo) % Ho), b = (int)(m / ((long long)Wo * Ho));
      const int iy = oy * stride - pad + kh, ix = ox * stride - pad + kw;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(img + (((long long)b * Cin + ci) * H + iy) * W + ix);
    }
    col[i] = __float2bfloat16(v);
  }
}
// dx[b, y, x, c] = (add ? add : 0) + sum over (kh, kw) with (y + p - kh) % s == 0 ... of dcol[m(oy, ox), (kh*KW + kw)*C + c]
// (gather form of the transposed convolution: deterministic, no atomics)
__global__ void col2im_nhwc_kernel(const bf16* __restrict__ dcol, const bf16* __restrict__ add, bf16* __restrict__ dx, int B, int H,
                                   int W, int C, int KH, int KW, int stride, int pad, int Ho, int Wo) {
  const int C8 = C >> 3;
  const long long Kp = (long long)KH * KW * C;
  const long long n = (long long)B * H * W * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long t = i / C8;
    const int xq = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    float acc[8];
    if (add) unpack8(__ldg(reinterpret_cast<const uint4*>(add) + i), acc);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    }
    for (int kh = 0; kh < KH; ++kh) {
      const int ty = y + pad - kh;
      if (ty < 0 || ty % stride != 0) continue;
      const int oy = ty / stride;
      if (oy >= Ho) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int tx = xq + pad - kw;
        if (tx < 0 || tx % stride != 0) continue;
        const int ox = tx / stride;
        if (ox >= Wo) continue;
        const long long m = ((long long)b * Ho + oy) * Wo + ox;
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dcol + m * Kp + (long long)(kh * KW + kw) * C) + c8), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    }
    reinterpret_cast<uint4*>(dx)[i] = pack8(acc);
  }
}

// ---------------------------------------------------------------- BatchNorm (training mode), x [M, C] bf16
// column sums of x and x^2 (block (32, 8): a thread owns 8 adjacent channels, 8 row-lanes; one atomic per column per block)
__global__ void __launch_bounds__(256) bn_stats_kernel(const bf16* __restrict__ x, float* __restrict__ sum, float* __restrict__ sumsq,
                                                       int M, int C, int rows_per_block) {
  __shared__ float red[2][8][256 + 8];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * 256 + tx * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float a[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = q[j] = 0.f;
  if (col < C) {
    for (int r = r0 + ty; r < r1; r += 8) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + (size_t)r * C + col)), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] += v[j]; q[j] += v[j] * v[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[0][ty][tx * 8 + j] = a[j]; red[1][ty][tx * 8 + j] = q[j]; }
  __syncthreads();
  const int c = ty * 32 + tx, gc = blockIdx.x * 256 + c;
  if (gc < C) {
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += red[0][k][c]; s2 += red[1][k][c]; }
    atomicAdd(sum + gc, s);
    atomicAdd(sumsq + gc, s2);
  }
}
// per channel: batch mean / biased variance -> scale = gamma * rstd, shift = beta - mean * scale; running statistics as
// torch.nn.BatchNorm2d updates them (momentum, UNBIASED variance); num_batches_tracked += 1
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ num_batches, int M, int C, float eps,
                                   float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches) *num_batches += 1;
  if (c >= C) return;
  const float mu = sum[c] / (float)M;
  const float var = fmaxf(sumsq[c] / (float)M - mu * mu, 0.f);
  const float rs = rsqrtf(var + eps);
  mean[c] = mu;
  rstd[c] = rs;
  const float sc = gamma[c] * rs;
  scale[c] = sc;
  shift[c] = beta[c] - mu * sc;
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * ((float)M / (float)max(M - 1, 1));
  }
}
// evaluation mode: scale / shift from the running statistics
__global__ void bn_eval_coeffs_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ scale,
                                      float* __restrict__ shift, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(running_var[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - running_mean[c] * sc;
}
// y = relu?(x * scale[c] + shift[c] (+ res))
__global__ void bn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                const bf16* __restrict__ res, bf16* __restrict__ y, long long n8, int C8, int relu) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    float v[8], r[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), v);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c + 4));
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c)), h1 = __ldg(reinterpret_cast<const float4*>(shift + c + 4));
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    if (res) unpack8(__ldg(reinterpret_cast<const uint4*>(res) + i), r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float o = fmaf(v[j], sc[j], sh[j]);
      if (res) o += r[j];
      v[j] = relu ? fmaxf(o, 0.f) : o;
    }
    reinterpret_cast<uint4*>(y)[i] = pack8(v);
  }
}
// backward, pass 1: g = relu ? (y > 0 ? dy : 0) : dy;  sum_g[c] += g, sum_gx[c] += g * xhat, xhat = (x - mean) * rstd
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                                            const bf16* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ sum_g,
                                                            float* __restrict__ sum_gx, int M, int C, int rows_per_block, int relu) {
  __shared__ float red[2][8][256 + 8];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * 256 + tx * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float a[8], q[8], mu[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = q[j] = 0.f;
  if (col < C) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { mu[j] = mean[col + j]; rs[j] = rstd[col + j]; }
    for (int r = r0 + ty; r < r1; r += 8) {
      float g[8], xv[8], yv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (size_t)r * C + col)), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + (size_t)r * C + col)), xv);
      if (relu) unpack8(__ldg(reinterpret_cast<const uint4*>(y + (size_t)r * C + col)), yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gj = (relu && !(yv[j] > 0.f)) ? 0.f : g[j];
        a[j] += gj;
        q[j] += gj * (xv[j] - mu[j]) * rs[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[0][ty][tx * 8 + j] = a[j]; red[1][ty][tx * 8 + j] = q[j]; }
  __syncthreads();
  const int c = ty * 32 + tx, gc = blockIdx.x * 256 + c;
  if (gc < C) {
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += red[0][k][c]; s2 += red[1][k][c]; }
    atomicAdd(sum_g + gc, s);
    atomicAdd(sum_gx + gc, s2);
  }
}
// backward, pass 2: dx = gamma * rstd * (g - sum_g / M - xhat * sum_gx / M);  dres = g (gradient of the residual branch);
// dgamma += sum_gx, dbeta += sum_g are accumulated by bn_bwd_param_kernel
__global__ void bn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y, const bf16* __restrict__ x,
                                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                    const float* __restrict__ sum_g, const float* __restrict__ sum_gx, bf16* __restrict__ dx,
                                    bf16* __restrict__ dres, long long n8, int C8, float inv_m, int relu) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    float g[8], xv[8], yv[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + i), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), xv);
    if (relu) unpack8(__ldg(reinterpret_cast<const uint4*>(y) + i), yv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gj = (relu && !(yv[j] > 0.f)) ? 0.f : g[j];
      const float rs = rstd[c + j];
      const float xh = (xv[j] - mean[c + j]) * rs;
      o[j] = gamma[c + j] * rs * (gj - sum_g[c + j] * inv_m - xh * sum_gx[c + j] * inv_m);
      g[j] = gj;
    }
    reinterpret_cast<uint4*>(dx)[i] = pack8(o);
    if (dres) reinterpret_cast<uint4*>(dres)[i] = pack8(g);
  }
}
__global__ void bn_bwd_param_kernel(const float* __restrict__ sum_g, const float* __restrict__ sum_gx, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dgamma[c] += sum_gx[c];
  dbeta[c] += sum_g[c];
}

// ---------------------------------------------------------------- max-pool 3x3 / stride 2 / pad 1 (torchvision ResNet stem)
// first maximum in (kh, kw) scan order wins (strict >, as ATen), its window position 0..8 is recorded for the backward
__global__ void maxpool3x3s2_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, uint8_t* __restrict__ idx, int B, int H, int W,
                                        int C, int Ho, int Wo) {
  const int C8 = C >> 3;
  const long long n = (long long)B * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long t = i / C8;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float best[8];
    int pos[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; pos[j] = 0; }
    bool first = true;
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = oy * 2 - 1 + kh;
      if (iy < 0 || iy >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = ox * 2 - 1 + kw;
        if (ix < 0 || ix >= W) continue;
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long long)b * H + iy) * W + ix) * C) + c8), v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (first || v[j] > best[j]) { best[j] = v[j]; pos[j] = kh * 3 + kw; }
        first = false;
      }
    }
    reinterpret_cast<uint4*>(y)[i] = pack8(best);
    uint2 p;
    p.x = (uint32_t)pos[0] | ((uint32_t)pos[1] << 8) | ((uint32_t)pos[2] << 16) | ((uint32_t)pos[3] << 24);
    p.y = (uint32_t)pos[4] | ((uint32_t)pos[5] << 8) | ((uint32_t)pos[6] << 16) | ((uint32_t)pos[7] << 24);
    reinterpret_cast<uint2*>(idx)[i] = p;
  }
}
__global__ void maxpool3x3s2_bwd_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ idx, bf16* __restrict__ dx, int B, int H,
                                        int W, int C, int Ho, int Wo) {
  const int C8 = C >> 3;
  const long long n = (long long)B * H * W * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long t = i / C8;
    const int xq = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int a = 0; a < 2; ++a) {                      // windows with oy*2 - 1 <= y <= oy*2 + 1: oy in {(y+1)/2, (y+1)/2 - 1}
      const int oy = (y + 1) / 2 - a;
      const int kh = y - (oy * 2 - 1);
      if (oy < 0 || oy >= Ho || kh < 0 || kh > 2) continue;
      for (int c = 0; c < 2; ++c) {
        const int ox = (xq + 1) / 2 - c;
        const int kw = xq - (ox * 2 - 1);
        if (ox < 0 || ox >= Wo || kw < 0 || kw > 2) continue;
        const long long m = (((long long)b * Ho + oy) * Wo + ox) * C8 + c8;
        const uint2 p = __ldg(reinterpret_cast<const uint2*>(idx) + m);
        float g[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + m), g);
        const uint32_t want = (uint32_t)(kh * 3 + kw);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t pj = ((j < 4 ? p.x : p.y) >> (8 * (j & 3))) & 0xffu;
          if (pj == want) acc[j] += g[j];
        }
      }
    }
    reinterpret_cast<uint4*>(dx)[i] = pack8(acc);
  }
}

// ---------------------------------------------------------------- global average pool: y[b, c] = mean_hw x[b, hw, c]
__global__ void avgpool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int HW, int C) {
  const int C8 = C >> 3;
  const long long n = (long long)B * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8), b = (int)(i / C8);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int p = 0; p < HW; ++p) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + ((long long)b * HW + p) * C) + c8), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    const float inv = 1.f / (float)HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    reinterpret_cast<uint4*>(y)[i] = pack8(acc);
  }
}
__global__ void avgpool_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int B, int HW, int C) {
  const int C8 = C >> 3;
  const long long n = (long long)B * HW * C8;
  const float inv = 1.f / (float)HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int b = (int)(i / ((long long)HW * C8));
    float g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + (long long)b * C8 + c8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= inv;
    reinterpret_cast<uint4*>(dx)[i] = pack8(g);
  }
}

}  // namespace vlm

using namespace vlm;

#define VLM_S(stream) reinterpret_cast<cudaStream_t>(stream)

extern "C" int vlm_conv_weight_pack(const float* w, void* wm, int Cout, int Cin, int KH, int KW, int Kp, void* stream) {
  VLM_REQUIRE(w && wm && Cout > 0 && Cin > 0 && KH > 0 && KW > 0 && Kp >= KH * KW * Cin && Kp % 8 == 0,
              "vlm_conv_weight_pack: bad args (Cout=%d Cin=%d k=%dx%d Kp=%d)", Cout, Cin, KH, KW, Kp);
  conv_weight_pack_kernel<<<conv_grid((long long)Cout * Kp, 256), 256, 0, VLM_S(stream)>>>(w, (bf16*)wm, Cout, Cin, KH, KW, Kp);
  return check_launch("conv_weight_pack");
}

extern "C" int vlm_conv_wgrad_unpack(const float* dwm, float* gw, int Cout, int Cin, int KH, int KW, int Kp, void* stream) {
  VLM_REQUIRE(dwm && gw && Cout > 0 && Cin > 0 && KH > 0 && KW > 0 && Kp >= KH * KW * Cin, "vlm_conv_wgrad_unpack: bad args");
  conv_wgrad_unpack_kernel<<<conv_grid((long long)Cout * Cin * KH * KW, 256), 256, 0, VLM_S(stream)>>>(dwm, gw, Cout, Cin, KH, KW, Kp);
  return check_launch("conv_wgrad_unpack");
}

static int conv_out(int H, int k, int s, int p) { return (H + 2 * p - k) / s + 1; }

extern "C" int vlm_im2col_nhwc(const void* x, void* col, int B, int H, int W, int C, int KH, int KW, int stride, int pad, void* stream) {
  VLM_REQUIRE(x && col && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0,
              "vlm_im2col_nhwc: bad args (C=%d must be a multiple of 8)", C);
  const int Ho = conv_out(H, KH, stride, pad), Wo = conv_out(W, KW, stride, pad);
  VLM_REQUIRE(Ho > 0 && Wo > 0, "vlm_im2col_nhwc: empty output");
  const long long n = (long long)B * Ho * Wo * KH * KW * (C / 8);
  im2col_nhwc_kernel<<<conv_grid(n, 256), 256, 0, VLM_S(stream)>>>((const bf16*)x, (bf16*)col, B, H, W, C, KH, KW, stride, pad, Ho, Wo);
  return check_launch("im2col_nhwc");
}

extern "C" int vlm_im2col_nchw_f32(const float* img, void* col, int B, int Cin, int H, int W, int KH, int KW, int stride, int pad, int Kp,
                                   void* stream) {
  VLM_REQUIRE(img && col && B > 0 && Cin > 0 && H > 0 && W > 0 && Kp >= KH * KW * Cin && Kp % 8 == 0, "vlm_im2col_nchw_f32: bad args");
  const int Ho = conv_out(H, KH, stride, pad), Wo = conv_out(W, KW, stride, pad);
  VLM_REQUIRE(Ho > 0 && Wo > 0, "vlm_im2col_nchw_f32: empty output");
  im2col_nchw_f32_kernel<<<conv_grid((long long)B * Ho * Wo * Kp, 256), 256, 0, VLM_S(stream)>>>(img, (bf16*)col, B, Cin, H, W, KH, KW, stride,
                                                                                              pad, Ho, Wo, Kp);
  return check_launch("im2col_nchw_f32");
}

extern "C" int vlm_col2im_nhwc(const void* dcol, const void* add, void* dx, int B, int H, int W, int C, int KH, int KW, int stride, int pad,
                               void* stream) {
  VLM_REQUIRE(dcol && dx && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0,
              "vlm_col2im_nhwc: bad args");
  const int Ho = conv_out(H, KH, stride, pad), Wo = conv_out(W, KW, stride, pad);
  col2im_nhwc_kernel<<<conv_grid((long long)B * H * W * (C / 8), 256), 256, 0, VLM_S(stream)>>>((const bf16*)dcol, (const bf16*)add, (bf16*)dx,
                                                                                             B, H, W, C, KH, KW, stride, pad, Ho, Wo);
  return check_launch("col2im_nhwc");
}

static void bn_reduce_grid(int M, int C, dim3* grid, int* rows_per_block) {
  const int col_blocks = (C + 255) / 256;
  int row_blocks = (num_sms() * 4 + col_blocks - 1) / col_blocks;
  if (row_blocks > (M + 63) / 64) row_blocks = (M + 63) / 64;
  if (row_blocks < 1) row_blocks = 1;
  *rows_per_block = (M + row_blocks - 1) / row_blocks;
  *grid = dim3(col_blocks, row_blocks);
}

extern "C" int vlm_bn_train_fwd(const void* x, const void* res, void* y, const float* gamma, const float* beta, float* mean, float* rstd,
                                float* scale, float* shift, float* sum_ws, float* running_mean, float* running_var,
                                long long* num_batches, int M, int C, float eps, float momentum, int relu, void* stream) {
  VLM_REQUIRE(x && y && gamma && beta && mean && rstd && scale && shift && sum_ws && M > 0 && C > 0 && C % 8 == 0,
              "vlm_bn_train_fwd: bad args (M=%d C=%d)", M, C);
  cudaStream_t s = VLM_S(stream);
  cudaMemsetAsync(sum_ws, 0, sizeof(float) * 2 * C, s);
  dim3 grid;
  int rpb;
  bn_reduce_grid(M, C, &grid, &rpb);
  bn_stats_kernel<<<grid, dim3(32, 8), 0, s>>>((const bf16*)x, sum_ws, sum_ws + C, M, C, rpb);
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(sum_ws, sum_ws + C, gamma, beta, mean, rstd, scale, shift, running_mean, running_var,
                                                     num_batches, M, C, eps, momentum);
  const long long n8 = (long long)M * (C / 8);
  bn_apply_kernel<<<conv_grid(n8, 256), 256, 0, s>>>((const bf16*)x, scale, shift, (const bf16*)res, (bf16*)y, n8, C / 8, relu);
  return check_launch("bn_train_fwd");
}

extern "C" int vlm_bn_eval_fwd(const void* x, const void* res, void* y, const float* gamma, const float* beta, const float* running_mean,
                               const float* running_var, float* scale, float* shift, int M, int C, float eps, int relu, void* stream) {
  VLM_REQUIRE(x && y && gamma && beta && running_mean && running_var && scale && shift && M > 0 && C > 0 && C % 8 == 0,
              "vlm_bn_eval_fwd: bad args");
  cudaStream_t s = VLM_S(stream);
  bn_eval_coeffs_kernel<<<(C + 127) / 128, 128, 0, s>>>(running_mean, running_var, gamma, beta, scale, shift, C, eps);
  const long long n8 = (long long)M * (C / 8);
  bn_apply_kernel<<<conv_grid(n8, 256), 256, 0, s>>>((const bf16*)x, scale, shift, (const bf16*)res, (bf16*)y, n8, C / 8, relu);
  return check_launch("bn_eval_fwd");
}

extern "C" int vlm_bn_train_bwd(const void* dy, const void* y, const void* x, const float* mean, const float* rstd, const float* gamma,
                                float* dgamma, float* dbeta, float* sum_ws, void* dx, void* dres, int M, int C, int relu, void* stream) {
  VLM_REQUIRE(dy && x && mean && rstd && gamma && dgamma && dbeta && sum_ws && dx && (!relu || y) && M > 0 && C > 0 && C % 8 == 0,
              "vlm_bn_train_bwd: bad args");
  cudaStream_t s = VLM_S(stream);
  cudaMemsetAsync(sum_ws, 0, sizeof(float) * 2 * C, s);
  dim3 grid;
  int rpb;
  bn_reduce_grid(M, C, &grid, &rpb);
  bn_bwd_reduce_kernel<<<grid, dim3(32, 8), 0, s>>>((const bf16*)dy, (const bf16*)y, (const bf16*)x, mean, rstd, sum_ws, sum_ws + C, M, C, rpb,
                                                   relu);
  bn_bwd_param_kernel<<<(C + 127) / 128, 128, 0, s>>>(sum_ws, sum_ws + C, dgamma, dbeta, C);
  const long long n8 = (long long)M * (C / 8);
  bn_bwd_apply_kernel<<<conv_grid(n8, 256), 256, 0, s>>>((const bf16*)dy, (const bf16*)y, (const bf16*)x, mean, rstd, gamma, sum_ws, sum_ws + C,
                                                        (bf16*)dx, (bf16*)dres, n8, C / 8, 1.f / (float)M, relu);
  return check_launch("bn_train_bwd");
}

extern "C" int vlm_maxpool3x3s2_fwd(const void* x, void* y, uint8_t* idx, int B, int H, int W, int C, void* stream) {
  VLM_REQUIRE(x && y && idx && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "vlm_maxpool3x3s2_fwd: bad args");
  const int Ho = conv_out(H, 3, 2, 1), Wo = conv_out(W, 3, 2, 1);
  maxpool3x3s2_fwd_kernel<<<conv_grid((long long)B * Ho * Wo * (C / 8), 256), 256, 0, VLM_S(stream)>>>((const bf16*)x, (bf16*)y, idx, B, H, W, C,
                                                                                                    Ho, Wo);
  return check_launch("maxpool3x3s2_fwd");
}

extern "C" int vlm_maxpool3x3s2_bwd(const void* dy, const uint8_t* idx, void* dx, int B, int H, int W, int C, void* stream) {
  VLM_REQUIRE(dy && idx && dx && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "vlm_maxpool3x3s2_bwd: bad args");
  const int Ho = conv_out(H, 3, 2, 1), Wo = conv_out(W, 3, 2, 1);
  maxpool3x3s2_bwd_kernel<<<conv_grid((long long)B * H * W * (C / 8), 256), 256, 0, VLM_S(stream)>>>((const bf16*)dy, idx, (bf16*)dx, B, H, W, C,
                                                                                                  Ho, Wo);
  return check_launch("maxpool3x3s2_bwd");
}

extern "C" int vlm_avgpool_fwd(const void* x, void* y, int B, int HW, int C, void* stream) {
  VLM_REQUIRE(x && y && B > 0 && HW > 0 && C > 0 && C % 8 == 0, "vlm_avgpool_fwd: bad args");
  avgpool_fwd_kernel<<<conv_grid((long long)B * (C / 8), 128), 128, 0, VLM_S(stream)>>>((const bf16*)x, (bf16*)y, B, HW, C);
  return check_launch("avgpool_fwd");
}

extern "C" int vlm_avgpool_bwd(const void* dy, void* dx, int B, int HW, int C, void* stream) {
  VLM_REQUIRE(dy && dx && B > 0 && HW > 0 && C > 0 && C % 8 == 0, "vlm_avgpool_bwd: bad args");
  avgpool_bwd_kernel<<<conv_grid((long long)B * HW * (C / 8), 256), 256, 0, VLM_S(stream)>>>((const bf16*)dy, (bf16*)dx, B, HW, C);
  return check_launch("avgpool_bwd");
}
