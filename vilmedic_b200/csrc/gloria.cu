// GLoRIA local (word x region) contrastive loss, batched over ALL (image, caption) pairs at once.
//   reference: vilmedic/blocks/losses/selfsup/GLoRIALoss.py:13-51 (gloria_attention_fn), :78-129 (local_loss),
//              :5-10 (cosine_similarity).  The reference loops over captions in Python and launches 2 bmm + 2 softmax per
//              caption; here every (image i, caption j, word w, region s) term is one element of a few large tensors.
//
// Data layout (B images, B captions, S = ih*iw regions, L = padded words per caption, NL = B*L, D features):
//   Xc  bf16 [B, S, D]     image regions, K-major rows (transposed from the reference's [B, D, ih, iw])
//   Ww  bf16 [NL, D]       words of all captions, rows (j, w); rows w >= cap_len[j] are zero
//   A   fp32 [B*S, NL]     A[(i,s),(j,w)] = c_is . q_jw                (tcgen05 GEMM, vlm_gemm_bf16)
//   P1  fp32 [B*S, NL]     softmax over the words of caption j         (:32-33)           gloria_word_softmax
//   P2  fp32 + bf16 [B, S, NL]  softmax over regions of temp1 * P1     (:38-43)           gloria_region_softmax
//   WC  fp32 [B, NL, D]    weighted context  WC_i = P2_i^T Xc_i        (:49; batched tcgen05 GEMM, both operands MN-major)
//   cos fp32 [B, NL]       cosine(q_jw, WC_i,jw) with the (|q||wc|).clamp(eps) of :5-10                  gloria_cos
//   sims fp32 [B, B]       temp3 * log sum_w exp(temp2 * cos)          (:112-122)         gloria_sims
// followed by vlm_sym_lse for the two cross-entropies (:127-128).  The backward mirrors this with four more GEMMs.
// Every kernel here is HBM-bound: one read (+ one write) of its [B*S, NL] or [B, NL, D] operand; no atomics, deterministic.
#include "common.cuh"
#include "vlm_b200.h"

namespace vlm {

// ---------------------------------------------------------------------------------------------- batched transpose + cast
// out[b][c][r] = in[b][r][c]  for r < R, c < C;  rows c in [C, Cout) and rows c >= row_limit[b] are zero-filled.
// bf16 output optionally comes as a hi/lo pair (x = hi + lo to ~16 mantissa bits) for the split tensor-core products.
template <typename OutT>
__global__ void transpose_cast_kernel(const float* __restrict__ in, OutT* __restrict__ out, bf16* __restrict__ out_lo, int R, int C,
                                      long long in_ld, long long in_bs, long long out_ld, long long out_bs, int Cout,
                                      const int* __restrict__ row_limit) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;                       // 32 x 8
  const float* ib = in + (size_t)b * in_bs;
  OutT* ob = out + (size_t)b * out_bs;
  const int lim = row_limit ? min(row_limit[b], C) : C;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    tile[ty + 8 * k][tx] = (r < R && c < lim) ? ib[(size_t)r * in_ld + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (c < Cout && r < R) {
      const float v = tile[tx][ty + 8 * k];
      if constexpr (sizeof(OutT) == 2) {
        const bf16 hi = __float2bfloat16(v);
        ob[(size_t)c * out_ld + r] = hi;
        if (out_lo) out_lo[(size_t)b * out_bs + (size_t)c * out_ld + r] = __float2bfloat16(v - __bfloat162float(hi));
      } else {
        ob[(size_t)c * out_ld + r] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- forward softmaxes
// P1[(i,s),(j,w)] = softmax_w(A[(i,s),(j,:cap_len[j])]); zero for w >= cap_len[j].   One CTA per row, one warp per caption.
__global__ void gloria_word_softmax_kernel(const float* __restrict__ A, float* __restrict__ P1, const int* __restrict__ cap_lens,
                                           int NB, int L, long long ld) {
  const size_t row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* a = A + row * ld;
  float* p = P1 + row * ld;
  for (int j = warp; j < NB; j += nwarps) {
    const int n = min(cap_lens[j], L);
    float m = -INFINITY;
    for (int w = lane; w < n; w += 32) m = fmaxf(m, a[j * L + w]);
    m = warp_max(m);
    float s = 0.f;
    for (int w = lane; w < n; w += 32) s += __expf(a[j * L + w] - m);
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int w = lane; w < L; w += 32) p[j * L + w] = w < n ? __expf(a[j * L + w] - m) * inv : 0.f;
  }
}

// P2[i,s,c] = softmax_s(temp1 * P1[(i,s),c]) per column c = (j,w); 0 <= temp1*P1 <= temp1, so no max subtraction is needed.
// One thread per column (adjacent threads -> adjacent columns, coalesced), grid (NL/128, B).
__global__ void gloria_region_softmax_kernel(const float* __restrict__ P1, float* __restrict__ P2, bf16* __restrict__ P2h,
                                             bf16* __restrict__ P2l, const int* __restrict__ cap_lens, int S, int L, long long ld, float temp1) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (c >= ld) return;
  const int j = c / L, w = c - j * L;
  const bool valid = w < cap_lens[j];
  const float* p = P1 + (size_t)i * S * ld + c;
  float* o = P2 + (size_t)i * S * ld + c;
  bf16* oh = P2h + (size_t)i * S * ld + c;
  bf16* ol = P2l + (size_t)i * S * ld + c;
  if (!valid) {
    for (int s = 0; s < S; ++s) {
      o[(size_t)s * ld] = 0.f;
      oh[(size_t)s * ld] = __float2bfloat16(0.f);
      ol[(size_t)s * ld] = __float2bfloat16(0.f);
    }
    return;
  }
  float sum = 0.f;
  for (int s = 0; s < S; ++s) sum += __expf(temp1 * p[(size_t)s * ld]);
  const float inv = 1.f / sum;
  for (int s = 0; s < S; ++s) {
    const float v = __expf(temp1 * p[(size_t)s * ld]) * inv;
    const bf16 hi = __float2bfloat16(v);
    o[(size_t)s * ld] = v;
    oh[(size_t)s * ld] = hi;
    ol[(size_t)s * ld] = __float2bfloat16(v - __bfloat162float(hi));
  }
}

// ---------------------------------------------------------------------------------------------- cosine + similarity
// One CTA per column c = (j,w), one warp per image i (strided): cos[i,c] = q.wc / max(|q||wc|, eps); also |wc| and |q|.
__global__ void gloria_cos_kernel(const float* __restrict__ WC, const float* __restrict__ Q, const int* __restrict__ cap_lens,
                                  float* __restrict__ cosv, float* __restrict__ wnorm, float* __restrict__ qnorm, int NB, int L,
                                  int NL, int D, float eps) {
  const int c = blockIdx.x, j = c / L, w = c - j * L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const bool valid = w < cap_lens[j];
  const float* q = Q + (size_t)c * D;
  float a2 = 0.f;
  for (int d = lane; d < D; d += 32) a2 += q[d] * q[d];
  const float a = sqrtf(warp_sum(a2));
  if (warp == 0 && lane == 0) qnorm[c] = a;
  for (int i = warp; i < NB; i += nwarps) {
    if (!valid) {
      if (lane == 0) { cosv[(size_t)i * NL + c] = 0.f; wnorm[(size_t)i * NL + c] = 0.f; }
      continue;
    }
    const float4* wc4 = reinterpret_cast<const float4*>(WC + ((size_t)i * NL + c) * D);
    const float4* q4 = reinterpret_cast<const float4*>(q);
    float n = 0.f, b2 = 0.f;
    for (int d = lane; d < D / 4; d += 32) {
      const float4 x = wc4[d], y = q4[d];
      n += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
      b2 += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    n = warp_sum(n);
    b2 = warp_sum(b2);
    if (lane == 0) {
      const float b = sqrtf(b2);
      cosv[(size_t)i * NL + c] = n / fmaxf(a * b, eps);
      wnorm[(size_t)i * NL + c] = b;
    }
  }
}

// sims[i,j] = temp3 * log(sum_{w < cap_len[j]} exp(temp2 * cos[i,(j,w)]))      (agg = "sum", GLoRIALoss.py:112-122)
__global__ void gloria_sims_kernel(const float* __restrict__ cosv, const int* __restrict__ cap_lens, float* __restrict__ sims,
                                   int NB, int L, int NL, float temp2, float temp3) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NB * NB) return;
  const int i = t / NB, j = t - i * NB;
  const int n = min(cap_lens[j], L);
  const float* cr = cosv + (size_t)i * NL + (size_t)j * L;
  float s = 0.f;
  for (int w = 0; w < n; ++w) s += __expf(temp2 * cr[w]);       // |temp2 * cos| <= temp2: no overflow
  sims[t] = temp3 * __logf(s);
}

// ---------------------------------------------------------------------------------------------- backward
// One CTA per column c = (j,w), one warp per image (strided).  With g0/g1 the upstream gradients of loss0/loss1 (means of the
// row / column cross-entropies of sims):
//   dsims[i,j] = (g0 (exp(sims_ij - lse_row_i) - [i==j]) + g1 (exp(sims_ij - lse_col_j) - [i==j])) / B
//   dcos       = dsims * temp3 * temp2 * exp(temp2 cos - sims/temp3)
//   dWC[i,c,:] = dcos (q / (ab) - cos wc / b^2)         (bf16, operand of the two following GEMMs)
//   dQ[c,:]    = sum_i dcos (wc / (ab) - cos q / a^2)   (fp32, summed across the CTA's warps through shared memory)
// If ab < eps the clamp is active and cos = n / eps is bilinear: dWC = dcos q / eps, dQ = dcos wc / eps.
__global__ void gloria_cos_bwd_kernel(const float* __restrict__ WC, const float* __restrict__ Q, const int* __restrict__ cap_lens,
                                      const float* __restrict__ cosv, const float* __restrict__ wnorm,
                                      const float* __restrict__ qnorm, const float* __restrict__ sims,
                                      const float* __restrict__ lse_row, const float* __restrict__ lse_col,
                                      const float* __restrict__ g0p, const float* __restrict__ g1p, bf16* __restrict__ dWC,
                                      float* __restrict__ dQ, int NB, int L, int NL, int D, float temp2, float temp3, float eps) {
  extern __shared__ float sdq[];                                // [nwarps][D]
  const int c = blockIdx.x, j = c / L, w = c - j * L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const bool valid = w < cap_lens[j];
  const float* q = Q + (size_t)c * D;
  if (!valid) {
    for (int i = warp; i < NB; i += nwarps) {
      uint2* o = reinterpret_cast<uint2*>(dWC + ((size_t)i * NL + c) * D);
      for (int d = lane; d < D / 4; d += 32) o[d] = make_uint2(0u, 0u);
    }
    for (int d = threadIdx.x; d < D; d += blockDim.x) dQ[(size_t)c * D + d] = 0.f;
    return;
  }
  const float a = qnorm[c];
  const float g0 = g0p ? *g0p : 1.f, g1 = g1p ? *g1p : 1.f;
  float* mine = sdq + (size_t)warp * D;
  for (int d = lane; d < D; d += 32) mine[d] = 0.f;
  const float4* q4 = reinterpret_cast<const float4*>(q);
  for (int i = warp; i < NB; i += nwarps) {
    const float sij = sims[(size_t)i * NB + j];
    const float diag = i == j ? 1.f : 0.f;
    const float dsim = (g0 * (__expf(sij - lse_row[i]) - diag) + g1 * (__expf(sij - lse_col[j]) - diag)) / (float)NB;
    const float cs = cosv[(size_t)i * NL + c], b = wnorm[(size_t)i * NL + c];
    const float dcos = dsim * temp3 * temp2 * __expf(temp2 * cs - sij / temp3);
    float kq, kw_wc, kw_q, kq_q;       // dWC = kq*q + kw_wc*wc ; dQ += kw_q*wc + kq_q*q
    if (a * b >= eps) {
      const float iab = 1.f / (a * b);
      kq = dcos * iab; kw_wc = -dcos * cs / (b * b);
      kw_q = dcos * iab; kq_q = -dcos * cs / (a * a);
    } else {
      kq = dcos / eps; kw_wc = 0.f; kw_q = dcos / eps; kq_q = 0.f;
    }
    const float4* wc4 = reinterpret_cast<const float4*>(WC + ((size_t)i * NL + c) * D);
    uint2* o = reinterpret_cast<uint2*>(dWC + ((size_t)i * NL + c) * D);
    for (int d = lane; d < D / 4; d += 32) {
      const float4 x = wc4[d], y = q4[d];
      o[d] = make_uint2(pack_bf16x2(kq * y.x + kw_wc * x.x, kq * y.y + kw_wc * x.y),
                        pack_bf16x2(kq * y.z + kw_wc * x.z, kq * y.w + kw_wc * x.w));
      float* m = mine + 4 * d;
      m[0] += kw_q * x.x + kq_q * y.x; m[1] += kw_q * x.y + kq_q * y.y;
      m[2] += kw_q * x.z + kq_q * y.z; m[3] += kw_q * x.w + kq_q * y.w;
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nwarps; ++k) s += sdq[(size_t)k * D + d];
    dQ[(size_t)c * D + d] = s;
  }
}

// column softmax backward, in place: G[i,s,c] <- temp1 * P2 (G - sum_s P2 G)   (= dL/dP1)
__global__ void gloria_region_softmax_bwd_kernel(const float* __restrict__ P2, float* __restrict__ G, int S, long long ld,
                                                 float temp1) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (c >= ld) return;
  const float* p = P2 + (size_t)i * S * ld + c;
  float* g = G + (size_t)i * S * ld + c;
  float dot = 0.f;
  for (int s = 0; s < S; ++s) dot += p[(size_t)s * ld] * g[(size_t)s * ld];
  for (int s = 0; s < S; ++s) g[(size_t)s * ld] = temp1 * p[(size_t)s * ld] * (g[(size_t)s * ld] - dot);
}

// word softmax backward: dA[(i,s),(j,w)] = P1 (G - sum_w P1 G)  -> bf16 (operand of the dXc / dWw GEMMs)
__global__ void gloria_word_softmax_bwd_kernel(const float* __restrict__ P1, const float* __restrict__ G, bf16* __restrict__ dA,
                                               const int* __restrict__ cap_lens, int NB, int L, long long ld) {
  const size_t row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* p = P1 + row * ld;
  const float* g = G + row * ld;
  bf16* o = dA + row * ld;
  for (int j = warp; j < NB; j += nwarps) {
    const int n = min(cap_lens[j], L);
    float dot = 0.f;
    for (int w = lane; w < n; w += 32) dot += p[j * L + w] * g[j * L + w];
    dot = warp_sum(dot);
    for (int w = lane; w < L; w += 32)
      o[j * L + w] = __float2bfloat16(w < n ? p[j * L + w] * (g[j * L + w] - dot) : 0.f);
  }
}

}  // namespace vlm

using namespace vlm;

extern "C" int vlm_transpose_cast(const float* in, void* out, void* out_lo, int out_bf16, int batch, int R, int C, long long in_ld,
                                  long long in_bs, long long out_ld, long long out_bs, int Cout, const int* row_limit,
                                  void* stream) {
  VLM_REQUIRE(in && out && batch > 0 && R > 0 && C > 0 && Cout >= C && in_ld >= C && out_ld >= R, "vlm_transpose_cast: bad args");
  VLM_REQUIRE(batch <= 65535 && (Cout + 31) / 32 <= 65535, "vlm_transpose_cast: grid too large");
  dim3 grid((R + 31) / 32, (Cout + 31) / 32, batch), block(32, 8);
  if (out_bf16)
    transpose_cast_kernel<bf16><<<grid, block, 0, (cudaStream_t)stream>>>(in, (bf16*)out, (bf16*)out_lo, R, C, in_ld, in_bs, out_ld, out_bs, Cout, row_limit);
  else
    transpose_cast_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(in, (float*)out, nullptr, R, C, in_ld, in_bs, out_ld, out_bs, Cout, row_limit);
  return check_launch("transpose_cast");
}

extern "C" int vlm_gloria_word_softmax(const float* A, float* P1, const int* cap_lens, int rows, int NB, int L, long long ld,
                                       void* stream) {
  VLM_REQUIRE(A && P1 && cap_lens && rows > 0 && NB > 0 && L > 0 && ld >= (long long)NB * L, "vlm_gloria_word_softmax: bad args");
  gloria_word_softmax_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(A, P1, cap_lens, NB, L, ld);
  return check_launch("gloria_word_softmax");
}

extern "C" int vlm_gloria_region_softmax(const float* P1, float* P2, void* P2h, void* P2l, const int* cap_lens, int NI, int S, int NB, int L,
                                         float temp1, void* stream) {
  VLM_REQUIRE(P1 && P2 && P2h && P2l && cap_lens && NI > 0 && S > 0 && NB > 0 && L > 0 && NI <= 65535, "vlm_gloria_region_softmax: bad args");
  const long long ld = (long long)NB * L;
  dim3 grid((unsigned)((ld + 127) / 128), NI);
  gloria_region_softmax_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P1, P2, (bf16*)P2h, (bf16*)P2l, cap_lens, S, L, ld, temp1);
  return check_launch("gloria_region_softmax");
}

extern "C" int vlm_gloria_cos(const float* WC, const float* Q, const int* cap_lens, float* cosv, float* wnorm, float* qnorm, int NI,
                              int NB, int L, int D, float eps, void* stream) {
  VLM_REQUIRE(WC && Q && cap_lens && cosv && wnorm && qnorm && NI > 0 && NB > 0 && L > 0 && D > 0 && D % 4 == 0,
              "vlm_gloria_cos: bad args (D must be a multiple of 4)");
  gloria_cos_kernel<<<NB * L, 256, 0, (cudaStream_t)stream>>>(WC, Q, cap_lens, cosv, wnorm, qnorm, NI, L, NB * L, D, eps);
  return check_launch("gloria_cos");
}

extern "C" int vlm_gloria_sims(const float* cosv, const int* cap_lens, float* sims, int NB, int L, float temp2, float temp3,
                               void* stream) {
  VLM_REQUIRE(cosv && cap_lens && sims && NB > 0 && L > 0, "vlm_gloria_sims: bad args");
  gloria_sims_kernel<<<(NB * NB + 127) / 128, 128, 0, (cudaStream_t)stream>>>(cosv, cap_lens, sims, NB, L, NB * L, temp2, temp3);
  return check_launch("gloria_sims");
}

extern "C" int vlm_gloria_cos_bwd(const float* WC, const float* Q, const int* cap_lens, const float* cosv, const float* wnorm,
                                  const float* qnorm, const float* sims, const float* lse_row, const float* lse_col,
                                  const float* g0, const float* g1, void* dWC, float* dQ, int NB, int L, int D, float temp2,
                                  float temp3, float eps, void* stream) {
  VLM_REQUIRE(WC && Q && cap_lens && cosv && wnorm && qnorm && sims && lse_row && lse_col && dWC && dQ, "vlm_gloria_cos_bwd: null");
  VLM_REQUIRE(NB > 0 && L > 0 && D > 0 && D % 4 == 0 && D <= 6144, "vlm_gloria_cos_bwd: D must be a multiple of 4, <= 6144");
  const size_t smem = (size_t)8 * D * sizeof(float);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(gloria_cos_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("vlm_gloria_cos_bwd: cannot reserve %zu B of shared memory", smem);
    return -2;
  }
  gloria_cos_bwd_kernel<<<NB * L, 256, smem, (cudaStream_t)stream>>>(WC, Q, cap_lens, cosv, wnorm, qnorm, sims, lse_row, lse_col, g0,
                                                                    g1, (bf16*)dWC, dQ, NB, L, NB * L, D, temp2, temp3, eps);
  return check_launch("gloria_cos_bwd");
}

extern "C" int vlm_gloria_region_softmax_bwd(const float* P2, float* G, int NI, int S, long long ld, float temp1, void* stream) {
  VLM_REQUIRE(P2 && G && NI > 0 && NI <= 65535 && S > 0 && ld > 0, "vlm_gloria_region_softmax_bwd: bad args");
  dim3 grid((unsigned)((ld + 127) / 128), NI);
  gloria_region_softmax_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P2, G, S, ld, temp1);
  return check_launch("gloria_region_softmax_bwd");
}

extern "C" int vlm_gloria_word_softmax_bwd(const float* P1, const float* G, void* dA, const int* cap_lens, int rows, int NB, int L,
                                           long long ld, void* stream) {
  VLM_REQUIRE(P1 && G && dA && cap_lens && rows > 0 && NB > 0 && L > 0 && ld >= (long long)NB * L, "vlm_gloria_word_softmax_bwd: bad args");
  gloria_word_softmax_bwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(P1, G, (bf16*)dA, cap_lens, NB, L, ld);
  return check_launch("gloria_word_softmax_bwd");
}
