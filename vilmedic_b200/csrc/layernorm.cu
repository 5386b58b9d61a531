// LayerNorm forward / backward, one warp per row, 128-bit vector access, row cached in registers (exact two-pass
// statistics).  HBM-bound: algorithmic bytes = read x + write y (fwd), read x,dy + write dx (bwd).
// Replaces nn.LayerNorm in HF ViT (modeling_vit.py:333,340,455 layernorm_before/after/final) and BertGeneration
// (modeling_bert_generation.py:52-56 SelfOutput.LayerNorm, 288-292 Output.LayerNorm, 410-429 embeddings.LayerNorm),
// reached from vilmedic/blocks/vision/visual_encoder.py:180-186 and vilmedic/blocks/huggingface/decoder/decoder_model.py:42-47.
#include "common.cuh"
#include <cstdlib>
#include "vlm_b200.h"

namespace vlm {

template <typename T>
struct Vec8;
template <>
struct Vec8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  static __device__ __forceinline__ void store(bf16* p, const float* v) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

// NV = number of 8-element vectors per lane (D <= NV*256)
template <typename TIn, int NV>
__global__ void __launch_bounds__(128) layernorm_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, bf16* __restrict__ y,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int M, int D, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= M) return;
  const int nvec = D >> 3;
  const TIn* xr = x + (size_t)row * D;
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      Vec8<TIn>::load(xr + vi * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  bf16* yr = y + (size_t)row * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      float g[8], b[8], o[8];
      Vec8<float>::load(gamma + vi * 8, g);
      Vec8<float>::load(beta + vi * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
      Vec8<bf16>::store(yr + vi * 8, o);
    }
  }
}

// Backward.  Persistent over rows so that the per-column dgamma/dbeta partials live in registers; one atomicAdd per
// column per CTA at the end.  dx (+= dres) has the dtype of x.
template <typename TIn, int NV>
__global__ void __launch_bounds__(128) layernorm_bwd_kernel(const bf16* __restrict__ dy, const TIn* __restrict__ x,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const TIn* __restrict__ dres,
                                                            TIn* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int M, int D, TIn* __restrict__ dx_drop,
                                                            float p_drop, unsigned long long seed, unsigned long long offset,
                                                            const unsigned long long* __restrict__ offset_ptr,
                                                            float* __restrict__ colsum) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  float dg[NV][8], db[NV][8], cs[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[i][j] = db[i][j] = cs[i][j] = 0.f;
  const Philox rng(seed);
  if (dx_drop && offset_ptr) offset += __ldg(offset_ptr);
  const uint32_t thr = (uint32_t)(p_drop * 4294967296.0f);
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;

  // Rows are software-pipelined: the raw x / dy vectors (and mean / rstd) of this warp's NEXT row are requested before the
  // current row is reduced, so two rows of loads are in flight per warp (the kernel is latency-bound at 12 warps / SM).
  constexpr int XW = sizeof(TIn) == 2 ? 1 : 2;   // uint4 words per 8-element vector of x
  // (round 2: the residual gradient `dres` travels with them — it used to be loaded where it is consumed, one exposed DRAM
  //  round trip per row and warp; profiles/launches_r2 LayerNorm backward 26 us -> see ROUND_NOTES)
  uint4 nx[NV][XW], ndy[NV], nres[NV][XW];
  float nmu = 0.f, nrs = 0.f;
  auto prefetch = [&](int prow) {
    if (prow < M) {
      const uint4* xr4 = reinterpret_cast<const uint4*>(x + (size_t)prow * D);
      const uint4* dy4 = reinterpret_cast<const uint4*>(dy + (size_t)prow * D);
      const uint4* rs4 = dres ? reinterpret_cast<const uint4*>(dres + (size_t)prow * D) : nullptr;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
#pragma unroll
          for (int w = 0; w < XW; ++w) nx[i][w] = xr4[vi * XW + w];
          ndy[i] = dy4[vi];
          if (rs4) {
#pragma unroll
            for (int w = 0; w < XW; ++w) nres[i][w] = rs4[vi * XW + w];
          }
        }
      }
      nmu = mean[prow];
      nrs = rstd[prow];
    }
  };
  prefetch(blockIdx.x * 4 + warp);
  for (int row = blockIdx.x * 4 + warp; row < M; row += gridDim.x * 4) {
    const float mu = nmu, rs = nrs;
    uint4 cx[NV][XW], cdy[NV], cres[NV][XW];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
      for (int w = 0; w < XW; ++w) {
        cx[i][w] = nx[i][w];
        cres[i][w] = nres[i][w];
      }
      cdy[i] = ndy[i];
    }
    prefetch(row + gridDim.x * 4);
    // pass 1 over the raw registers: row statistics + column partials; pass 2 re-derives xhat / dy*gamma from the same raw
    // registers (cheaper in registers than keeping 2 x 8 x NV floats alive across the warp reductions)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float xv[8], dv[8], gm[8];
        Vec8<TIn>::load(reinterpret_cast<const TIn*>(&cx[i][0]), xv);
        Vec8<bf16>::load(reinterpret_cast<const bf16*>(&cdy[i]), dv);
        Vec8<float>::load(gamma + vi * 8, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mu) * rs;
          const float g = dv[j] * gm[j];
          s1 += g;
          s2 += g * xh;
          dg[i][j] += dv[j] * xh;
          db[i][j] += dv[j];
        }
      }
    }
    const float c1 = warp_sum(s1) / (float)D, c2 = warp_sum(s2) / (float)D;
    TIn* dxr = dx + (size_t)row * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float o[8];
        {
          float xv[8], dv[8], gm[8];
          Vec8<TIn>::load(reinterpret_cast<const TIn*>(&cx[i][0]), xv);
          Vec8<bf16>::load(reinterpret_cast<const bf16*>(&cdy[i]), dv);
          Vec8<float>::load(gamma + vi * 8, gm);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rs * (dv[j] * gm[j] - c1 - (xv[j] - mu) * rs * c2);
        }
        if (dres) {
          float r[8];
          Vec8<TIn>::load(reinterpret_cast<const TIn*>(&cres[i][0]), r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        Vec8<TIn>::store(dxr + vi * 8, o);
        if (dx_drop) {
          // same Philox stream as the GEMM epilogue / vlm_dropout_bf16 on the flat [M, D] tensor
          const unsigned long long base = ((unsigned long long)row * (unsigned long long)D + (unsigned long long)vi * 8ull) >> 2;
          const uint4 r0 = rng(base, offset), r1 = rng(base + 1, offset);
          o[0] = r0.x >= thr ? o[0] * inv_keep : 0.f; o[1] = r0.y >= thr ? o[1] * inv_keep : 0.f;
          o[2] = r0.z >= thr ? o[2] * inv_keep : 0.f; o[3] = r0.w >= thr ? o[3] * inv_keep : 0.f;
          o[4] = r1.x >= thr ? o[4] * inv_keep : 0.f; o[5] = r1.y >= thr ? o[5] * inv_keep : 0.f;
          o[6] = r1.z >= thr ? o[6] * inv_keep : 0.f; o[7] = r1.w >= thr ? o[7] * inv_keep : 0.f;
          Vec8<TIn>::store(dx_drop + (size_t)row * D + vi * 8, o);
        }
        if (colsum) {
#pragma unroll
          for (int j = 0; j < 8; ++j) cs[i][j] += o[j];
        }
      }
    }
  }

  // cross-warp reduction of the column partials, then one atomic per column per CTA
  __shared__ float red[4][256];
  const int npass = colsum ? 3 : 2;
#pragma unroll 1
  for (int pass = 0; pass < npass; ++pass) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = pass == 0 ? dg[i][j] : (pass == 1 ? db[i][j] : cs[i][j]);
      __syncthreads();
      // 256 columns of this vector slot: thread t sums columns t and t+128
      for (int c = threadIdx.x; c < 256; c += 128) {
        const int vi = (c >> 3) + 32 * i;  // vector index within the row
        if (vi < nvec) {
          const float tot = red[0][c] + red[1][c] + red[2][c] + red[3][c];
          float* dst = (pass == 0 ? dgamma : (pass == 1 ? dbeta : colsum));
          if (dst) atomicAdd(dst + vi * 8 + (c & 7), tot);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// Round-2 kernels for bf16 rows with D = WPR * 256 (768 -> WPR = 3): a row is shared by WPR warps, every lane owns ONE 16-byte
// vector (8 columns).  Compared with the warp-per-row kernels above this cuts the per-thread state of the backward from 226
// registers (8 warps / SM, issue-latency bound: ncu r2l_ln_bwd — 36 instructions per element, 48 % issue-active) to < 85
// (24 warps / SM), keeps gamma in registers for the whole persistent loop, does the arithmetic on the packed fp32x2 pipe and
// derives xhat / dy*gamma once.  Row statistics cross the WPR warps through a double-buffered shared-memory slot and one named
// barrier per row; the column partials (dgamma, dbeta, bias-gradient column sums) leave with 16-byte vector atomics.
static constexpr int LN2_G = 4;   // row groups per CTA

__device__ __forceinline__ void unpack8(const uint4& u, float2 (&v)[4]) {
  v[0] = unpack_bf16x2(u.x); v[1] = unpack_bf16x2(u.y); v[2] = unpack_bf16x2(u.z); v[3] = unpack_bf16x2(u.w);
}
__device__ __forceinline__ uint4 pack8(const float2 (&v)[4]) {
  return make_uint4(pack_bf16x2(v[0].x, v[0].y), pack_bf16x2(v[1].x, v[1].y), pack_bf16x2(v[2].x, v[2].y), pack_bf16x2(v[3].x, v[3].y));
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {      // read-once data: do not keep it in L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <int WPR>
__global__ void __launch_bounds__(WPR * 32 * LN2_G, 2)
layernorm_fwd_v2_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                        bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, float eps) {
  constexpr int D = WPR * 256;
  __shared__ float xch[2][2][LN2_G][WPR];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp / WPR, wr = warp - grp * WPR;
  const int col = (wr * 32 + lane) * 8;
  float2 gm[4], bt[4];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + col), g1 = *reinterpret_cast<const float4*>(gamma + col + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + col), b1 = *reinterpret_cast<const float4*>(beta + col + 4);
    gm[0] = make_float2(g0.x, g0.y); gm[1] = make_float2(g0.z, g0.w); gm[2] = make_float2(g1.x, g1.y); gm[3] = make_float2(g1.z, g1.w);
    bt[0] = make_float2(b0.x, b0.y); bt[1] = make_float2(b0.z, b0.w); bt[2] = make_float2(b1.x, b1.y); bt[3] = make_float2(b1.z, b1.w);
  }
  const int stride = gridDim.x * LN2_G;
  int row = blockIdx.x * LN2_G + grp;
  uint4 nx = make_uint4(0u, 0u, 0u, 0u);
  if (row < M) nx = ldg_stream(reinterpret_cast<const uint4*>(x + (size_t)row * D + col));
  for (int it = 0; row < M; row += stride, ++it) {
    const uint4 cx = nx;
    if (row + stride < M) nx = ldg_stream(reinterpret_cast<const uint4*>(x + (size_t)(row + stride) * D + col));
    float2 v[4];
    unpack8(cx, v);
    float2 s2 = __fadd2_rn(__fadd2_rn(v[0], v[1]), __fadd2_rn(v[2], v[3]));
    float s = warp_sum(s2.x + s2.y);
    if (WPR > 1) {
      if (lane == 0) xch[it & 1][0][grp][wr] = s;
      named_bar_sync(1 + grp, WPR * 32);
      s = 0.f;
#pragma unroll
      for (int w = 0; w < WPR; ++w) s += xch[it & 1][0][grp][w];
    }
    const float mean = s * (1.f / (float)D);
    const float2 nm = make_float2(-mean, -mean);
    float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = __fadd2_rn(v[j], nm);
      q2 = __ffma2_rn(v[j], v[j], q2);
    }
    float q = warp_sum(q2.x + q2.y);
    if (WPR > 1) {
      if (lane == 0) xch[it & 1][1][grp][wr] = q;
      named_bar_sync(1 + grp, WPR * 32);
      q = 0.f;
#pragma unroll
      for (int w = 0; w < WPR; ++w) q += xch[it & 1][1][grp][w];
    }
    const float rstd = rsqrtf(q * (1.f / (float)D) + eps);
    if (wr == 0 && lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    const float2 r2 = make_float2(rstd, rstd);
    float2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __ffma2_rn(__fmul2_rn(v[j], r2), gm[j], bt[j]);
    *reinterpret_cast<uint4*>(y + (size_t)row * D + col) = pack8(o);
  }
}

template <int WPR, bool DROP, bool COLSUM>
__global__ void __launch_bounds__(WPR * 32 * LN2_G, WPR <= 3 ? 2 : 1)
layernorm_bwd_v2_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ gamma, const bf16* __restrict__ dres,
                        bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int M,
                        bf16* __restrict__ dx_drop, float p_drop, unsigned long long seed, unsigned long long offset,
                        const unsigned long long* __restrict__ offset_ptr, float* __restrict__ colsum) {
  constexpr int D = WPR * 256;
  __shared__ float2 xch[2][LN2_G][WPR];
  __shared__ __align__(16) float red[LN2_G][D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp / WPR, wr = warp - grp * WPR;
  const int col = (wr * 32 + lane) * 8;
  float2 gm[4];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + col), g1 = *reinterpret_cast<const float4*>(gamma + col + 4);
    gm[0] = make_float2(g0.x, g0.y); gm[1] = make_float2(g0.z, g0.w); gm[2] = make_float2(g1.x, g1.y); gm[3] = make_float2(g1.z, g1.w);
  }
  float2 dg[4], db[4], cs[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) dg[j] = db[j] = cs[j] = make_float2(0.f, 0.f);
  const Philox rng(seed);
  if (DROP && offset_ptr) offset += __ldg(offset_ptr);
  const uint32_t thr = (uint32_t)(p_drop * 4294967296.0f);
  const float inv_keep = DROP ? 1.f / (1.f - p_drop) : 1.f;

  const int stride = gridDim.x * LN2_G;
  int row = blockIdx.x * LN2_G + grp;
  uint4 nx, ndy, nres = make_uint4(0u, 0u, 0u, 0u);
  float nmu = 0.f, nrs = 0.f;
  auto prefetch = [&](int prow) {
    if (prow < M) {
      const size_t o = (size_t)prow * D + col;
      nx = ldg_stream(reinterpret_cast<const uint4*>(x + o));
      ndy = ldg_stream(reinterpret_cast<const uint4*>(dy + o));
      if (dres) nres = ldg_stream(reinterpret_cast<const uint4*>(dres + o));
      nmu = __ldg(mean + prow);
      nrs = __ldg(rstd + prow);
    }
  };
  nx = ndy = nres;
  prefetch(row);
  for (int it = 0; row < M; row += stride, ++it) {
    const uint4 cx = nx, cdy = ndy, cres = nres;
    const float mu = nmu, rs = nrs;
    prefetch(row + stride);
    float2 xh[4], g[4];
    {
      float2 xv[4], dv[4];
      unpack8(cx, xv);
      unpack8(cdy, dv);
      const float2 rs2 = make_float2(rs, rs), nmr = make_float2(-mu * rs, -mu * rs);
      float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xh[j] = __ffma2_rn(xv[j], rs2, nmr);
        g[j] = __fmul2_rn(dv[j], gm[j]);
        s1 = __fadd2_rn(s1, g[j]);
        s2 = __ffma2_rn(g[j], xh[j], s2);
        dg[j] = __ffma2_rn(dv[j], xh[j], dg[j]);
        db[j] = __fadd2_rn(db[j], dv[j]);
      }
      float a = warp_sum(s1.x + s1.y), b = warp_sum(s2.x + s2.y);
      if (WPR > 1) {
        if (lane == 0) xch[it & 1][grp][wr] = make_float2(a, b);
        named_bar_sync(1 + grp, WPR * 32);
        a = b = 0.f;
#pragma unroll
        for (int w = 0; w < WPR; ++w) {
          const float2 t = xch[it & 1][grp][w];
          a += t.x;
          b += t.y;
        }
      }
      // dx = rs * (g - c1 - xhat * c2)  (+ dres)
      const float c1 = a * (1.f / (float)D), c2 = b * (1.f / (float)D);
      const float2 nc1 = make_float2(-c1 * rs, -c1 * rs), nc2 = make_float2(-c2 * rs, -c2 * rs);
      float2 rv[4];
      unpack8(cres, rv);                      // zeros when there is no residual gradient
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 t = __ffma2_rn(g[j], rs2, nc1);
        t = __ffma2_rn(xh[j], nc2, t);
        g[j] = __fadd2_rn(t, rv[j]);
      }
    }
    const size_t o = (size_t)row * D + col;
    *reinterpret_cast<uint4*>(dx + o) = pack8(g);
    if (DROP) {
      // same Philox stream as the GEMM epilogue / vlm_dropout_bf16 on the flat [M, D] tensor
      const unsigned long long base = (unsigned long long)o >> 2;
      const uint4 r0 = rng(base, offset), r1 = rng(base + 1, offset);
      g[0].x = r0.x >= thr ? g[0].x * inv_keep : 0.f; g[0].y = r0.y >= thr ? g[0].y * inv_keep : 0.f;
      g[1].x = r0.z >= thr ? g[1].x * inv_keep : 0.f; g[1].y = r0.w >= thr ? g[1].y * inv_keep : 0.f;
      g[2].x = r1.x >= thr ? g[2].x * inv_keep : 0.f; g[2].y = r1.y >= thr ? g[2].y * inv_keep : 0.f;
      g[3].x = r1.z >= thr ? g[3].x * inv_keep : 0.f; g[3].y = r1.w >= thr ? g[3].y * inv_keep : 0.f;
      *reinterpret_cast<uint4*>(dx_drop + o) = pack8(g);
    }
    if (COLSUM) {
#pragma unroll
      for (int j = 0; j < 4; ++j) cs[j] = __fadd2_rn(cs[j], g[j]);
    }
  }

  // column partials: across the row groups of the CTA through shared memory, then one 16-byte vector atomic per 4 columns
  constexpr int NPASS = COLSUM ? 3 : 2;
#pragma unroll
  for (int pass = 0; pass < NPASS; ++pass) {
    const float2* src = pass == 0 ? dg : (pass == 1 ? db : cs);
    float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : colsum);
    __syncthreads();
    *reinterpret_cast<float4*>(&red[grp][col]) = make_float4(src[0].x, src[0].y, src[1].x, src[1].y);
    *reinterpret_cast<float4*>(&red[grp][col + 4]) = make_float4(src[2].x, src[2].y, src[3].x, src[3].y);
    __syncthreads();
    if (dst) {
      for (int c = threadIdx.x * 4; c < D; c += WPR * 32 * LN2_G * 4) {
        float4 t = *reinterpret_cast<const float4*>(&red[0][c]);
#pragma unroll
        for (int gi = 1; gi < LN2_G; ++gi) {
          const float4 u = *reinterpret_cast<const float4*>(&red[gi][c]);
          t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        atomicAdd(reinterpret_cast<float4*>(dst + c), t);
      }
    }
  }
}

static int ln_variant() {          // VLM_LN_V2=0 selects the warp-per-row kernels (kept for fp32 rows and other widths)
  const char* v = getenv("VLM_LN_V2");
  return (v && v[0] == '0') ? 0 : 1;
}

// grid of the persistent v2 kernels: every row group gets the same number of rows (no partial last pass)
static int ln2_grid(int M, int max_ctas) {
  const int groups = max_ctas * LN2_G;
  const int rpg = (M + groups - 1) / groups;
  return (M + LN2_G * rpg - 1) / (LN2_G * rpg);
}

template <typename TIn>
static int ln_fwd_dispatch(const TIn* x, const float* gamma, const float* beta, bf16* y, float* mean, float* rstd, int M,
                           int D, float eps, cudaStream_t s) {
  if constexpr (sizeof(TIn) == 2) {
    if (D % 256 == 0 && D >= 512 && D <= 1024 && ln_variant() == 1 && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(beta) & 15) == 0) {
#define LN_FWD2(W_)                                                                                                     \
  {                                                                                                                     \
    static int occ = 0;                                                                                                 \
    if (occ == 0) {                                                                                                     \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layernorm_fwd_v2_kernel<W_>, W_ * 32 * LN2_G, 0) != cudaSuccess || occ < 1) occ = 2; \
    }                                                                                                                   \
    layernorm_fwd_v2_kernel<W_><<<ln2_grid(M, num_sms() * occ), W_ * 32 * LN2_G, 0, s>>>(x, gamma, beta, y, mean, rstd, M, eps); \
  }
      if (D == 512) LN_FWD2(2)
      else if (D == 768) LN_FWD2(3)
      else LN_FWD2(4)
#undef LN_FWD2
      return check_launch("layernorm_fwd_v2");
    }
  }
  const int nv = (D / 8 + 31) / 32;
  const int grid = (M + 3) / 4;                // (two rows per warp was tried in round 2: 0.78 vs 0.66 ms per step — worse)
#define LN_FWD(NV_) layernorm_fwd_kernel<TIn, NV_><<<grid, 128, 0, s>>>(x, gamma, beta, y, mean, rstd, M, D, eps)
  switch (nv) {
    case 1: LN_FWD(1); break;
    case 2: LN_FWD(2); break;
    case 3: LN_FWD(3); break;
    case 4: LN_FWD(4); break;
    case 5: case 6: LN_FWD(6); break;
    case 7: case 8: LN_FWD(8); break;
    default: set_error("layernorm: D=%d unsupported (max 2048)", D); return -1;
  }
#undef LN_FWD
  return check_launch("layernorm_fwd");
}

template <typename TIn>
static int ln_bwd_dispatch(const bf16* dy, const TIn* x, const float* mean, const float* rstd, const float* gamma,
                           const TIn* dres, TIn* dx, float* dgamma, float* dbeta, int M, int D, TIn* dx_drop, float p_drop,
                           unsigned long long seed, unsigned long long offset, const unsigned long long* offset_ptr,
                           float* colsum, cudaStream_t s) {
  if constexpr (sizeof(TIn) == 2) {
    const bool al16 = ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(dgamma) | reinterpret_cast<uintptr_t>(dbeta) |
                        reinterpret_cast<uintptr_t>(colsum)) & 15) == 0;
    if (D % 256 == 0 && D >= 512 && D <= 1024 && ln_variant() == 1 && al16) {
#define LN_BWD2K(W_, DR_, CS_)                                                                                           \
  {                                                                                                                     \
    static int occ = 0;                                                                                                 \
    if (occ == 0) {                                                                                                     \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layernorm_bwd_v2_kernel<W_, DR_, CS_>, W_ * 32 * LN2_G, 0) != cudaSuccess || occ < 1) occ = 1; \
    }                                                                                                                   \
    layernorm_bwd_v2_kernel<W_, DR_, CS_><<<ln2_grid(M, num_sms() * occ), W_ * 32 * LN2_G, 0, s>>>(                      \
        dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, M, dx_drop, p_drop, seed, offset, offset_ptr, colsum);       \
  }
#define LN_BWD2(W_)                                                                                                     \
  {                                                                                                                     \
    if (dx_drop) { if (colsum) LN_BWD2K(W_, true, true) else LN_BWD2K(W_, true, false) }                                \
    else { if (colsum) LN_BWD2K(W_, false, true) else LN_BWD2K(W_, false, false) }                                      \
  }
      if (D == 512) LN_BWD2(2)
      else if (D == 768) LN_BWD2(3)
      else LN_BWD2(4)
#undef LN_BWD2
#undef LN_BWD2K
      return check_launch("layernorm_bwd_v2");
    }
  }
  const int nv = (D / 8 + 31) / 32;
  // persistent grid = exactly one resident wave (SMs x occupancy): a partial second wave would cost a full pass
#define LN_BWD(NV_)                                                                                                     \
  {                                                                                                                     \
    static int occ = 0;                                                                                                 \
    if (occ == 0) {                                                                                                     \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layernorm_bwd_kernel<TIn, NV_>, 128, 0) != cudaSuccess || occ < 1) occ = 2; \
    }                                                                                                                   \
    int grid = num_sms() * occ;                                                                                         \
    if (grid > (M + 3) / 4) grid = (M + 3) / 4;                                                                         \
    layernorm_bwd_kernel<TIn, NV_><<<grid, 128, 0, s>>>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, M, D, dx_drop, p_drop, seed, offset, offset_ptr, colsum); \
  }
  switch (nv) {
    case 1: LN_BWD(1); break;
    case 2: LN_BWD(2); break;
    case 3: LN_BWD(3); break;
    case 4: LN_BWD(4); break;
    case 5: case 6: LN_BWD(6); break;
    case 7: case 8: LN_BWD(8); break;
    default: set_error("layernorm: D=%d unsupported (max 2048)", D); return -1;
  }
#undef LN_BWD
  return check_launch("layernorm_bwd");
}

}  // namespace vlm

using namespace vlm;

extern "C" int vlm_layernorm_fwd(const void* x, int x_is_fp32, const float* gamma, const float* beta, void* y,
                                 float* mean, float* rstd, int M, int D, float eps, void* stream) {
  VLM_REQUIRE(M > 0 && D > 0 && D % 8 == 0, "vlm_layernorm_fwd: need D %% 8 == 0 (M=%d D=%d)", M, D);
  VLM_REQUIRE(x && gamma && beta && y, "vlm_layernorm_fwd: null pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (x_is_fp32)
    return ln_fwd_dispatch<float>(reinterpret_cast<const float*>(x), gamma, beta, reinterpret_cast<bf16*>(y), mean, rstd,
                                  M, D, eps, s);
  return ln_fwd_dispatch<bf16>(reinterpret_cast<const bf16*>(x), gamma, beta, reinterpret_cast<bf16*>(y), mean, rstd, M, D,
                               eps, s);
}

extern "C" int vlm_layernorm_bwd(const void* dy, const void* x, int x_is_fp32, const float* mean, const float* rstd,
                                 const float* gamma, const void* dres, void* dx, float* dgamma, float* dbeta, int M,
                                 int D, void* dx_drop, float p_drop, unsigned long long seed, unsigned long long offset,
                                 const unsigned long long* rng_offset_ptr, float* colsum, void* stream) {
  VLM_REQUIRE(M > 0 && D > 0 && D % 8 == 0, "vlm_layernorm_bwd: need D %% 8 == 0 (M=%d D=%d)", M, D);
  VLM_REQUIRE(dy && x && mean && rstd && gamma && dx, "vlm_layernorm_bwd: null pointer");
  VLM_REQUIRE(!dx_drop || (p_drop > 0.f && p_drop < 1.f && D % 4 == 0), "vlm_layernorm_bwd: dx_drop needs 0 < p_drop < 1");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (x_is_fp32)
    return ln_bwd_dispatch<float>(reinterpret_cast<const bf16*>(dy), reinterpret_cast<const float*>(x), mean, rstd, gamma,
                                  reinterpret_cast<const float*>(dres), reinterpret_cast<float*>(dx), dgamma, dbeta, M, D,
                                  reinterpret_cast<float*>(dx_drop), p_drop, seed, offset, rng_offset_ptr, colsum, s);
  return ln_bwd_dispatch<bf16>(reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(x), mean, rstd, gamma,
                               reinterpret_cast<const bf16*>(dres), reinterpret_cast<bf16*>(dx), dgamma, dbeta, M, D,
                               reinterpret_cast<bf16*>(dx_drop), p_drop, seed, offset, rng_offset_ptr, colsum, s);
}
