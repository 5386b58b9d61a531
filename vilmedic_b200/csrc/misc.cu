// HBM-bound helper kernels around the GEMM / attention core: casts, ViT patch extraction (im2col for k = s = P is a
// pure permutation), CLS/position assembly and its backward reductions, bias-gradient column sums, the
// VisualEncoder feature mask, token+position embedding gather / scatter, and counter-based dropout.
#include "common.cuh"
#include "vlm_b200.h"

namespace vlm {

// ---------------------------------------------------------------- cast fp32 -> bf16 (8 per thread)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n8, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n8; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = u;
  }
  // tail
  i = n8 * 8 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(src[i]);
}

// ---------------------------------------------------------------- ViT patchify
// images fp32 [B,C,H,W] -> patches bf16 [B, 1 + (H/P)*(W/P), C*P*P]; row 0 of each image is zero (CLS slot), so the
// patch-embedding GEMM and its wgrad run on the same [B*S] row space as the rest of the encoder.
// Column order (c, ky, kx) matches Conv2d.weight[out, c, ky, kx].flatten(1)  (HF modeling_vit.py:151,166).
__global__ void patchify_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int C, int H, int W, int P, int n_prefix) {
  const int gw = W / P, gh = H / P;
  const int S = n_prefix + gh * gw;
  const int Kp = C * P * P;
  const int vec_per_row = Kp / 8;
  const long long total = (long long)B * S * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec_per_row);
    const long long rs = i / vec_per_row;
    const int s = (int)(rs % S), b = (int)(rs / S);
    uint4 u = make_uint4(0, 0, 0, 0);
    if (s >= n_prefix) {
      const int p = s - n_prefix, py = p / gw, px = p % gw;
      const int col = v * 8;  // 8 consecutive kx within one (c, ky) row since P % 8 == 0
      const int c = col / (P * P), rem = col % (P * P), ky = rem / P, kx = rem % P;
      const float* src = img + (((long long)b * C + c) * H + (py * P + ky)) * W + px * P + kx;
      const float4 a = __ldg(reinterpret_cast<const float4*>(src));
      const float4 d = __ldg(reinterpret_cast<const float4*>(src) + 1);
      u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(d.x, d.y); u.w = pack_bf16x2(d.z, d.w);
    }
    reinterpret_cast<uint4*>(out)[i] = u;
  }
}

// x[b,row,:] = tok + pos[row]   (HF modeling_vit.py:117-124: cat(cls, patches) + position_embeddings; DeiT: row 1 = distillation token)
template <typename T>
__global__ void vit_cls_kernel(T* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos, int B, int S, int D, int row) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i % D;
  const float v = cls[d] + pos[(size_t)row * D + d];
  if constexpr (sizeof(T) == 4) x[((size_t)b * S + row) * D + d] = v;
  else x[((size_t)b * S + row) * D + d] = __float2bfloat16(v);
}

// dpos[s,d] += sum_b dx[b,s,d];  dcls[d] += sum_b dx[b,0,d];  dbias[d] += sum_{b,s>=1} dx[b,s,d]
template <typename T>
__global__ void vit_embed_bwd_kernel(const T* __restrict__ dx, float* __restrict__ dpos, float* __restrict__ dcls,
                                     float* __restrict__ ddist, float* __restrict__ dbias, int B, int S, int D, int n_prefix) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (d >= D) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const T v = dx[((size_t)b * S + s) * D + d];
    if constexpr (sizeof(T) == 4) acc += v; else acc += __bfloat162float(v);
  }
  if (dpos) dpos[(size_t)s * D + d] += acc;
  if (s == 0) { if (dcls) dcls[d] += acc; }
  else if (s < n_prefix) { if (ddist) ddist[d] += acc; }
  else if (dbias) atomicAdd(dbias + d, acc);
}

// ---------------------------------------------------------------- column sums (bias gradients)
// out[n] (+)= sum_m x[m,n]; block (32,8): each thread owns 2 adjacent columns, 8 row-lanes, grid.y row chunks.
__global__ void colsum_kernel(const bf16* __restrict__ x, long long ld, float* __restrict__ out, int M, int N, int rows_per_block,
                              const float* __restrict__ scale_ptr) {
  __shared__ float red[8][64];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * 64 + tx * 2;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f;
  if (col < N) {
    for (int r = r0 + ty; r < r1; r += 8) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + (size_t)r * ld + col));
      a0 += v.x;
      a1 += v.y;
    }
  }
  red[ty][tx * 2] = a0;
  red[ty][tx * 2 + 1] = a1;
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { a0 += red[k][tx * 2]; a1 += red[k][tx * 2 + 1]; }
    const float sc = scale_ptr ? __ldg(scale_ptr) : 1.f;
    if (col < N) atomicAdd(out + col, a0 * sc);
    if (col + 1 < N) atomicAdd(out + col + 1, a1 * sc);
  }
}

// 16-byte variant: each thread owns 8 adjacent columns (one uint4 per row), rows unrolled x4 so that four independent loads
// are in flight per thread.  Needs a 16-byte aligned base and ld % 8 == 0; columns in [N, ld) of the last vector are read
// but never written back.
// TY = row lanes per CTA: 8 (256 threads) in the foreground, 4 (128 threads, fits next to a resident persistent GEMM CTA) in
// background mode (runtime.cu).
template <int TY>
__global__ void __launch_bounds__(32 * TY) colsum8_kernel(const bf16* __restrict__ x, long long ld, float* __restrict__ out, int M, int N,
                                                          int rows_per_block, const float* __restrict__ scale_ptr) {
  __shared__ float red[TY][256 + 8];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * 256 + tx * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (col < N) {
    const bf16* base = x + col;
    int r = r0 + ty;
    for (; r + 3 * TY < r1; r += 4 * TY) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(r + TY * k) * ld));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 p0 = unpack_bf16x2(u[k].x), p1 = unpack_bf16x2(u[k].y), p2 = unpack_bf16x2(u[k].z), p3 = unpack_bf16x2(u[k].w);
        a[0] += p0.x; a[1] += p0.y; a[2] += p1.x; a[3] += p1.y; a[4] += p2.x; a[5] += p2.y; a[6] += p3.x; a[7] += p3.y;
      }
    }
    for (; r < r1; r += TY) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (size_t)r * ld));
      const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
      a[0] += p0.x; a[1] += p0.y; a[2] += p1.x; a[3] += p1.y; a[4] += p2.x; a[5] += p2.y; a[6] += p3.x; a[7] += p3.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = a[j];
  __syncthreads();
  const float sc = scale_ptr ? __ldg(scale_ptr) : 1.f;
  for (int c = ty * 32 + tx; c < 256; c += 32 * TY) {     // threads <-> the 256 columns of the block
    const int gc = blockIdx.x * 256 + c;
    if (gc < N) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < TY; ++k) t += red[k][c];
      atomicAdd(out + gc, t * sc);
    }
  }
}

// ---------------------------------------------------------------- input pipeline: crop + flip + ToTensor + Normalize
// One thread per output pixel (all three channels): the three input bytes of a pixel are adjacent (HWC), the three output
// planes are written with unit stride along x.  Arithmetic order = torchvision: x.float().div(255) -> sub(mean) -> div(std),
// with IEEE round-to-nearest division and no FMA contraction, so the result is bit-identical to the CPU transform
// (vilmedic/datasets/base/ImageDataset.py:97-104).
__global__ void image_crop_flip_normalize_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, const int* __restrict__ top,
                                                 const int* __restrict__ left, const uint8_t* __restrict__ flip, int B, int Hin, int Win,
                                                 int crop, float m0, float m1, float m2, float s0, float s1, float s2) {
  const long long n = (long long)B * crop * crop;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % crop);
    const int y = (int)((i / crop) % crop);
    const int b = (int)(i / ((long long)crop * crop));
    const int sx = left[b] + (flip[b] ? crop - 1 - x : x);
    const int sy = top[b] + y;
    const uint8_t* px = in + (((long long)b * Hin + sy) * Win + sx) * 3;
    const float r = __fdiv_rn((float)px[0], 255.0f), g = __fdiv_rn((float)px[1], 255.0f), bl = __fdiv_rn((float)px[2], 255.0f);
    float* o = out + (long long)b * 3 * crop * crop + (long long)y * crop + x;
    o[0] = __fdiv_rn(__fsub_rn(r, m0), s0);
    o[(long long)crop * crop] = __fdiv_rn(__fsub_rn(g, m1), s1);
    o[2LL * crop * crop] = __fdiv_rn(__fsub_rn(bl, m2), s2);
  }
}

// ---------------------------------------------------------------- Resize of the reference's transform on the device
// torchvision Resize on a PIL image (vilmedic/datasets/base/ImageDataset.py:97-104 `transforms.Resize(resize)`) is Pillow's two-pass
// separable convolution resampling (ImagingResample: horizontal pass into an 8-bit intermediate, then the vertical pass), bilinear
// = triangle filter whose support grows with the down-scaling factor (antialias), coefficients normalised in double and applied in
// fixed point: acc = 2^21 + sum pixel * round(k * 2^22); out = clip8(acc >> 22).  The coefficient tables (bounds = first source index
// and tap count per output coordinate, ksize taps each) are computed on the host exactly as Pillow computes them
// (blocks/vision/preprocess.py: pil_resample_tables) and passed in; this kernel is one pass along one axis of uint8 HWC images.
// axis 0: out[b, y, xx, c] = sum_x in[b, y, xmin + x, c] * k[xx][x]  (Wout columns);  axis 1: the same along rows.
__global__ void resample_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int* __restrict__ bounds,
                                   const int* __restrict__ coefs, int ksize, int B, int Hin, int Win, int Hout, int Wout, int axis) {
  const long long n = (long long)B * Hout * Wout * 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    const int xx = (int)((i / 3) % Wout);
    const int yy = (int)((i / (3LL * Wout)) % Hout);
    const int b = (int)(i / (3LL * Wout * Hout));
    const int o = axis == 0 ? xx : yy;
    const int lo = bounds[2 * o], cnt = bounds[2 * o + 1];
    const int* k = coefs + (long long)o * ksize;
    int acc = 1 << 21;
    if (axis == 0) {
      const uint8_t* src = in + (((long long)b * Hin + yy) * Win + lo) * 3 + c;
      for (int x = 0; x < cnt; ++x) acc += (int)src[3 * x] * k[x];
    } else {
      const uint8_t* src = in + (((long long)b * Hin + lo) * Win + xx) * 3 + c;
      for (int y = 0; y < cnt; ++y) acc += (int)src[(long long)y * Win * 3] * k[y];
    }
    acc >>= 22;
    out[i] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
  }
}

// ---------------------------------------------------------------- VisualEncoder.encode feature mask
// mask[r] = (sum_d |f[r,d]| != 0)      (vilmedic/blocks/vision/visual_encoder.py:138)
__global__ void features_mask_kernel(const bf16* __restrict__ f, uint8_t* __restrict__ mask, int R, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const bf16* fr = f + (size_t)warp * D;
  float s = 0.f;
  for (int v = lane; v < D / 8; v += 32) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(fr) + v);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    s += fabsf(a.x) + fabsf(a.y) + fabsf(b.x) + fabsf(b.y) + fabsf(c.x) + fabsf(c.y) + fabsf(d.x) + fabsf(d.y);
  }
  s = warp_sum(s);
  if (lane == 0) mask[warp] = (s != 0.f) ? 1 : 0;
}

// ---------------------------------------------------------------- token + position embeddings
// z[r,:] = word[ids[r],:] + pos[pos_offset + r % T,:]   (HF modeling_bert_generation.py:410-429, before LayerNorm)
// pos_ids (optional int32 [R]): explicit position index per token (RoBERTa: padding_idx + running count of non-pad tokens,
// HF:roberta/modeling_roberta.py create_position_ids_from_input_ids), else pos_offset + (r % T).  tt_row (optional fp32 [D]): row 0 of
// the token-type table, added to every token (BERT / RoBERTa embeddings with token_type_ids = 0, which is all the reference passes).
__global__ void embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ pos,
                                 bf16* __restrict__ z, int R, int T, int D, int V, int pos_offset, const int* __restrict__ pos_ids,
                                 const float* __restrict__ tt_row) {
  const int nvec = D / 8;
  const long long total = (long long)R * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const int r = (int)(i / nvec);
    long long id = ids[r];
    if (id < 0 || id >= V) id = 0;  // defensive: the reference would raise an index error
    const float* w = word + (size_t)id * D + v * 8;
    const int pi = pos_ids ? pos_ids[r] : pos_offset + r % T;
    const float* p = pos + (size_t)pi * D + v * 8;
    float4 w0 = __ldg(reinterpret_cast<const float4*>(w)), w1 = __ldg(reinterpret_cast<const float4*>(w) + 1);
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(p)), p1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
    if (tt_row) {       // HF sums word + token_type first, then + position (modeling_bert.py BertEmbeddings.forward)
      const float4 t0 = __ldg(reinterpret_cast<const float4*>(tt_row + v * 8)), t1 = __ldg(reinterpret_cast<const float4*>(tt_row + v * 8) + 1);
      w0.x += t0.x; w0.y += t0.y; w0.z += t0.z; w0.w += t0.w;
      w1.x += t1.x; w1.y += t1.y; w1.z += t1.z; w1.w += t1.w;
    }
    uint4 u;
    u.x = pack_bf16x2(w0.x + p0.x, w0.y + p0.y); u.y = pack_bf16x2(w0.z + p0.z, w0.w + p0.w);
    u.z = pack_bf16x2(w1.x + p1.x, w1.y + p1.y); u.w = pack_bf16x2(w1.z + p1.z, w1.w + p1.w);
    reinterpret_cast<uint4*>(z)[i] = u;
  }
}

// dword[ids[r],:] += dz[r,:] ; dpos[position of r,:] += dz[r,:]
__global__ void embed_bwd_kernel(const long long* __restrict__ ids, const bf16* __restrict__ dz, float* __restrict__ dword,
                                 float* __restrict__ dpos, int R, int T, int D, int V, int pos_offset, int padding_idx,
                                 const int* __restrict__ pos_ids) {
  const int nvec = D / 2;
  const long long total = (long long)R * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const int r = (int)(i / nvec);
    long long id = ids[r];
    if (id < 0 || id >= V) id = 0;
    const float2 g = unpack_bf16x2(reinterpret_cast<const uint32_t*>(dz)[i]);
    if (dword && id != padding_idx) {  // nn.Embedding(padding_idx=...) never accumulates into the padding row
      atomicAdd(dword + (size_t)id * D + v * 2, g.x);
      atomicAdd(dword + (size_t)id * D + v * 2 + 1, g.y);
    }
    if (dpos) {
      const int pi = pos_ids ? pos_ids[r] : pos_offset + r % T;
      atomicAdd(dpos + (size_t)pi * D + v * 2, g.x);
      atomicAdd(dpos + (size_t)pi * D + v * 2 + 1, g.y);
    }
  }
}

// ---------------------------------------------------------------- dropout (Philox4x32-10, 8 elements / thread)
// y = x * keep / (1-p); the same (seed, offset) regenerates the mask for the backward pass.
__global__ void dropout_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n8, float p, float scale,
                               unsigned long long seed, unsigned long long offset, const unsigned long long* offset_ptr) {
  const Philox rng(seed);
  if (offset_ptr) offset += __ldg(offset_ptr);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = reinterpret_cast<const uint4*>(x)[i];
    const uint4 r0 = rng(2 * i, offset), r1 = rng(2 * i + 1, offset);
    const uint32_t thr = (uint32_t)(p * 4294967296.0f);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    a.x = r0.x >= thr ? a.x * scale : 0.f; a.y = r0.y >= thr ? a.y * scale : 0.f;
    b.x = r0.z >= thr ? b.x * scale : 0.f; b.y = r0.w >= thr ? b.y * scale : 0.f;
    c.x = r1.x >= thr ? c.x * scale : 0.f; c.y = r1.y >= thr ? c.y * scale : 0.f;
    d.x = r1.z >= thr ? d.x * scale : 0.f; d.y = r1.w >= thr ? d.y * scale : 0.f;
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(b.x, b.y); o.z = pack_bf16x2(c.x, c.y); o.w = pack_bf16x2(d.x, d.y);
    reinterpret_cast<uint4*>(y)[i] = o;
  }
}

// ---------------------------------------------------------------- zero the rows of masked-out images
// y[r,:] = mask[r / rows_per_mask] ? x[r,:] : 0      (vilmedic/blocks/vision/visual_encoder.py:170-171)
__global__ void mask_rows_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const uint8_t* __restrict__ mask, long long n8,
                                 int vec_per_row, int rows_per_mask) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const uint4 u = reinterpret_cast<const uint4*>(x)[i];
    reinterpret_cast<uint4*>(y)[i] = mask[r / rows_per_mask] ? u : make_uint4(0, 0, 0, 0);
  }
}

// ---------------------------------------------------------------- small fp32 activations (BertPooler tanh, ConVIRT projection ReLU)
// kind 0 = tanh, 1 = relu.  Backward takes the forward OUTPUT y.
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int kind) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = kind == 0 ? tanhf(x[i]) : fmaxf(x[i], 0.f);
}
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n, int kind) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = kind == 0 ? dy[i] * (1.f - y[i] * y[i]) : (y[i] > 0.f ? dy[i] : 0.f);
}

// ---------------------------------------------------------------- deterministic sum of a small fp32 vector
__global__ void sum_scale_kernel(const float* __restrict__ x, int n, float scale, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) *out = s * scale;
  }
}

static inline int grid_for(long long work, int block, int per_sm = 8) {
  long long g = (work + block - 1) / block;
  const long long cap = (long long)num_sms() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace vlm

using namespace vlm;

extern "C" int vlm_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  VLM_REQUIRE(src && dst && n >= 0, "vlm_cast_f32_to_bf16: bad args");
  if (n == 0) return 0;
  const long long n8 = n / 8;
  cast_f32_bf16_kernel<<<grid_for(n8 + 8, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n8, n);
  return check_launch("cast_f32_bf16");
}

extern "C" int vlm_patchify(const float* images, void* patches, int B, int C, int H, int W, int P, int n_prefix, void* stream) {
  VLM_REQUIRE(images && patches && B > 0 && C > 0 && n_prefix >= 0 && n_prefix <= 2, "vlm_patchify: bad args");
  VLM_REQUIRE(P % 8 == 0 && H % P == 0 && W % P == 0 && W % 4 == 0, "vlm_patchify: need P%%8==0, H,W multiples of P (H=%d W=%d P=%d)", H, W, P);
  const long long total = (long long)B * (n_prefix + (H / P) * (W / P)) * (C * P * P / 8);
  patchify_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(images, (bf16*)patches, B, C, H, W, P, n_prefix);
  return check_launch("patchify");
}

extern "C" int vlm_vit_cls_pos(void* x, int x_is_fp32, const float* cls, const float* pos, int B, int S, int D, int row, void* stream) {
  VLM_REQUIRE(x && cls && pos && B > 0 && S > 0 && D > 0 && row >= 0 && row < S, "vlm_vit_cls_pos: bad args");
  const int n = B * D;
  if (x_is_fp32) vit_cls_kernel<float><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((float*)x, cls, pos, B, S, D, row);
  else vit_cls_kernel<bf16><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((bf16*)x, cls, pos, B, S, D, row);
  return check_launch("vit_cls_pos");
}

extern "C" int vlm_vit_embed_bwd(const void* dx, int dx_is_fp32, float* dpos, float* dcls, float* ddist, float* dbias, int B, int S, int D,
                                 int n_prefix, void* stream) {
  VLM_REQUIRE(dx && B > 0 && S > 0 && D > 0 && n_prefix >= 1 && n_prefix <= 2, "vlm_vit_embed_bwd: bad args");
  dim3 grid((D + 127) / 128, S);
  if (dx_is_fp32) vit_embed_bwd_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>((const float*)dx, dpos, dcls, ddist, dbias, B, S, D, n_prefix);
  else vit_embed_bwd_kernel<bf16><<<grid, 128, 0, (cudaStream_t)stream>>>((const bf16*)dx, dpos, dcls, ddist, dbias, B, S, D, n_prefix);
  return check_launch("vit_embed_bwd");
}

extern "C" int vlm_colsum_bf16(const void* x, long long ld, float* out, int M, int N, const float* scale_ptr, void* stream) {
  VLM_REQUIRE(x && out && M > 0 && N > 0 && ld % 2 == 0 && ld >= N + (N & 1), "vlm_colsum_bf16: bad args (M=%d N=%d ld=%lld)", M, N, ld);
  if (ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (long long)((N + 7) / 8 * 8) <= ld) {
    const int col_blocks = (N + 255) / 256;
    const bool bg = background_mode();
    int row_blocks = ((bg ? num_sms_all() * 8 : num_sms() * 4) + col_blocks - 1) / col_blocks;
    if (row_blocks > (M + 63) / 64) row_blocks = (M + 63) / 64;
    if (row_blocks < 1) row_blocks = 1;
    const int rows_per_block = (M + row_blocks - 1) / row_blocks;
    if (bg) colsum8_kernel<4><<<dim3(col_blocks, row_blocks), dim3(32, 4), 0, (cudaStream_t)stream>>>((const bf16*)x, ld, out, M, N, rows_per_block, scale_ptr);
    else colsum8_kernel<8><<<dim3(col_blocks, row_blocks), dim3(32, 8), 0, (cudaStream_t)stream>>>((const bf16*)x, ld, out, M, N, rows_per_block, scale_ptr);
    return check_launch("colsum8");
  }
  const int col_blocks = (N + 63) / 64;
  int row_blocks = (num_sms() * 4 + col_blocks - 1) / col_blocks;
  if (row_blocks > (M + 63) / 64) row_blocks = (M + 63) / 64;
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_block = (M + row_blocks - 1) / row_blocks;
  colsum_kernel<<<dim3(col_blocks, row_blocks), dim3(32, 8), 0, (cudaStream_t)stream>>>((const bf16*)x, ld, out, M, N, rows_per_block, scale_ptr);
  return check_launch("colsum");
}

extern "C" int vlm_image_crop_flip_normalize(const uint8_t* in, float* out, const int* top, const int* left, const uint8_t* flip, int B,
                                             int Hin, int Win, int crop, const float* mean3, const float* std3, void* stream) {
  VLM_REQUIRE(in && out && top && left && flip && mean3 && std3, "vlm_image_crop_flip_normalize: null pointer");
  VLM_REQUIRE(B > 0 && crop > 0 && Hin >= crop && Win >= crop, "vlm_image_crop_flip_normalize: crop %d larger than image %dx%d (B=%d)",
              crop, Hin, Win, B);
  VLM_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "vlm_image_crop_flip_normalize: std must be non-zero");
  const long long n = (long long)B * crop * crop;
  image_crop_flip_normalize_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, top, left, flip, B, Hin, Win, crop, mean3[0],
                                                                                     mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  return check_launch("image_crop_flip_normalize");
}

extern "C" int vlm_image_resample_u8(const uint8_t* in, uint8_t* out, const int* bounds, const int* coefs, int ksize, int B, int Hin,
                                     int Win, int Hout, int Wout, int axis, void* stream) {
  VLM_REQUIRE(in && out && bounds && coefs && ksize > 0 && B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "vlm_image_resample_u8: bad args");
  VLM_REQUIRE(axis == 0 ? Hout == Hin : (axis == 1 && Wout == Win), "vlm_image_resample_u8: one axis per pass (axis 0: columns, axis 1: rows)");
  const long long n = (long long)B * Hout * Wout * 3;
  resample_u8_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, bounds, coefs, ksize, B, Hin, Win, Hout, Wout, axis);
  return check_launch("image_resample_u8");
}

extern "C" int vlm_features_mask(const void* feats, uint8_t* mask, int R, int D, void* stream) {
  VLM_REQUIRE(feats && mask && R > 0 && D % 8 == 0, "vlm_features_mask: bad args");
  features_mask_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>((const bf16*)feats, mask, R, D);
  return check_launch("features_mask");
}

extern "C" int vlm_embed_fwd(const long long* ids, const float* word, const float* pos, void* z, int R, int T, int D, int V,
                             int pos_offset, const int* pos_ids, const float* tt_row, void* stream) {
  VLM_REQUIRE(ids && word && pos && z && R > 0 && T > 0 && D % 8 == 0 && V > 0, "vlm_embed_fwd: bad args");
  embed_fwd_kernel<<<grid_for((long long)R * D / 8, 256), 256, 0, (cudaStream_t)stream>>>(ids, word, pos, (bf16*)z, R, T, D, V, pos_offset, pos_ids, tt_row);
  return check_launch("embed_fwd");
}

extern "C" int vlm_embed_bwd(const long long* ids, const void* dz, float* dword, float* dpos, int R, int T, int D, int V,
                             int pos_offset, int padding_idx, const int* pos_ids, void* stream) {
  VLM_REQUIRE(ids && dz && R > 0 && T > 0 && D % 2 == 0 && V > 0, "vlm_embed_bwd: bad args");
  embed_bwd_kernel<<<grid_for((long long)R * D / 2, 256), 256, 0, (cudaStream_t)stream>>>(ids, (const bf16*)dz, dword, dpos, R, T, D, V, pos_offset, padding_idx, pos_ids);
  return check_launch("embed_bwd");
}

extern "C" int vlm_dropout_bf16(const void* x, void* y, long long n, float p, unsigned long long seed,
                                unsigned long long offset, const unsigned long long* rng_offset_ptr, void* stream) {
  VLM_REQUIRE(x && y && n >= 0 && n % 8 == 0 && p >= 0.f && p < 1.f, "vlm_dropout_bf16: need n%%8==0 and 0<=p<1");
  if (n == 0) return 0;
  dropout_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, n / 8, p, 1.f / (1.f - p), seed, offset, rng_offset_ptr);
  return check_launch("dropout");
}

extern "C" int vlm_mask_rows_bf16(const void* x, void* y, const uint8_t* mask, int R, int D, int rows_per_mask, void* stream) {
  VLM_REQUIRE(x && y && mask && R > 0 && D % 8 == 0 && rows_per_mask > 0, "vlm_mask_rows_bf16: bad args");
  const long long n8 = (long long)R * (D / 8);
  mask_rows_kernel<<<grid_for(n8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, mask, n8, D / 8, rows_per_mask);
  return check_launch("mask_rows");
}

extern "C" int vlm_act_fwd_f32(const float* x, float* y, long long n, int kind, void* stream) {
  VLM_REQUIRE(x && y && n > 0 && (kind == 0 || kind == 1), "vlm_act_fwd_f32: bad args");
  act_fwd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, kind);
  return check_launch("act_fwd");
}

extern "C" int vlm_act_bwd_f32(const float* dy, const float* y, float* dx, long long n, int kind, void* stream) {
  VLM_REQUIRE(dy && y && dx && n > 0 && (kind == 0 || kind == 1), "vlm_act_bwd_f32: bad args");
  act_bwd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n, kind);
  return check_launch("act_bwd");
}

__global__ void u64_add_kernel(unsigned long long* p, unsigned long long v) { *p += v; }

extern "C" int vlm_rng_advance(unsigned long long* counter, unsigned long long delta, void* stream) {
  VLM_REQUIRE(counter, "vlm_rng_advance: null counter");
  u64_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta);
  return check_launch("rng_advance");
}

extern "C" int vlm_sum_scale_f32(const float* x, int n, float scale, float* out, void* stream) {
  VLM_REQUIRE(x && out && n > 0, "vlm_sum_scale_f32: bad args");
  sum_scale_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, scale, out);
  return check_launch("sum_scale");
}
