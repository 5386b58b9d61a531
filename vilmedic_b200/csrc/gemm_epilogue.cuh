// Fused GEMM epilogue shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels: one thread owns one accumulator row and
// processes it in chunks of 16 or 32 fp32 columns read from TMEM.  The epilogue is instruction-issue / latency bound, not
// bandwidth bound (ncu, profiles/ncu_gemm_r1_epilogue.txt): 2 epilogue warps per SM sub-partition reached only 38 % issue
// utilisation on the bias+GELU epilogue, so the 1-CTA kernel runs 16 epilogue warps (4 per sub-partition, 16-column chunks,
// <= 112 registers) and the operand loads of a chunk are issued under its TMEM load.
#pragma once
#include <cuda.h>
#include "common.cuh"

// Packed fp32x2 (FFMA2 / FMUL2 / FADD2) evaluation of GELU + GELU' in the staged epilogue: 13 packed + 10 scalar instructions per
// pair instead of 2 x 19 scalar ones; measured -5 % on the FFN-up forward GEMM (gpurun_out r2a: 87 vs 92 us), same results.
// Build with -DVLM_GELU_F32X2=0 for the scalar form.
#ifndef VLM_GELU_F32X2
#define VLM_GELU_F32X2 1
#endif

namespace vlm {

struct GemmEpilogue {
  void* c;
  long long ldc;
  int c_fp32;
  const float* bias;          // [N] or null
  const void* residual;       // same dtype as c, ld = ldr
  long long ldr;
  int act;                    // 0 none, 1 GELU (optionally stash GELU'(pre-activation) to aux_out), 2 multiply by aux_in
  const bf16* aux_in;         // [M, ld_aux]
  bf16* aux_out;              // [M, ld_aux]
  long long ld_aux;
  float alpha;
  const float* alpha_ptr;     // optional device scalar multiplied into alpha (upstream loss gradient)
  int accumulate;             // c += result
  int atomic;                 // split-K: fp32 C is updated with atomic adds (implies accumulate)
  float p_drop;               // dropout on the activation (after bias/act, before the residual add)
  unsigned long long seed, offset;
  const unsigned long long* offset_ptr;  // optional device-side addend to `offset` (CUDA-graph replayable RNG stream)
  int drop_ld;                // logical row width used for the dropout element index (row * drop_ld + col)
};

// One chunk = W (16 or 32) consecutive fp32 accumulator columns of this thread's row, read from TMEM at `taddr`.
// Order of operations: issue the TMEM load, issue the global loads the epilogue needs (GELU' input, residual, old C) while
// it is in flight, wait, do the math, store.  Must be called by all 32 lanes (tcgen05.ld / wait::ld are warp-collective).
template <int W>
__device__ __forceinline__ void epilogue_chunk(uint32_t taddr, int row, int col0, int M, int N, const GemmEpilogue& e) {
  static_assert(W == 16 || W == 32, "chunk width");
  uint32_t acc[W];
  if constexpr (W == 32) tmem_ld32(taddr, acc);
  else tmem_ld16(taddr, acc);
  const bool live = row < M && col0 < N;
  const bool full = live && (col0 + W <= N);
  // ---- prefetch (bf16 operands: W/8 x 16 B per row; fp32: W/4 x 16 B)
  uint4 aux[W / 8], res16[W / 8];
  float4 res32[W / 4];
  const bool pre_aux = full && e.act == 2;
  const bool pre_res16 = full && !e.c_fp32 && e.residual != nullptr;
  const bool pre_res32 = full && e.c_fp32 && e.residual != nullptr;
  if (pre_aux) {
    const uint4* src = reinterpret_cast<const uint4*>(e.aux_in + (long long)row * e.ld_aux + col0);
#pragma unroll
    for (int j = 0; j < W / 8; ++j) aux[j] = __ldg(src + j);
  }
  if (pre_res16) {
    const uint4* r4 = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(e.residual) + (long long)row * e.ldr + col0);
#pragma unroll
    for (int j = 0; j < W / 8; ++j) res16[j] = __ldg(r4 + j);
  }
  if (pre_res32) {
    const float4* r4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.residual) + (long long)row * e.ldr + col0);
#pragma unroll
    for (int j = 0; j < W / 4; ++j) res32[j] = __ldg(r4 + j);
  }
  tmem_ld_wait();
  if (!live) return;
  float v[W];
  const float alpha = e.alpha_ptr ? e.alpha * __ldg(e.alpha_ptr) : e.alpha;
  if (alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] = __uint_as_float(acc[j]) * alpha;
  } else {
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] = __uint_as_float(acc[j]);
  }

  if (full) {
    if (e.bias) {
      const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
      for (int j = 0; j < W / 4; ++j) {
        const float4 b = __ldg(b4 + j);
        v[4 * j + 0] += b.x;
        v[4 * j + 1] += b.y;
        v[4 * j + 2] += b.z;
        v[4 * j + 3] += b.w;
      }
    }
    if (e.act == 1) {
      if (e.aux_out) {
        float dg[W];
#pragma unroll
        for (int j = 0; j < W; ++j) v[j] = gelu_erf_both(v[j], dg[j]);
        uint4* dst = reinterpret_cast<uint4*>(e.aux_out + (long long)row * e.ld_aux + col0);
#pragma unroll
        for (int j = 0; j < W / 8; ++j) {
          uint4 u;
          u.x = pack_bf16x2(dg[8 * j + 0], dg[8 * j + 1]);
          u.y = pack_bf16x2(dg[8 * j + 2], dg[8 * j + 3]);
          u.z = pack_bf16x2(dg[8 * j + 4], dg[8 * j + 5]);
          u.w = pack_bf16x2(dg[8 * j + 6], dg[8 * j + 7]);
          dst[j] = u;
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) v[j] = gelu_erf(v[j]);
      }
    } else if (e.act == 2) {
#pragma unroll
      for (int j = 0; j < W / 8; ++j) {
        const uint4 u = aux[j];
        const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
        v[8 * j + 0] *= p0.x;
        v[8 * j + 1] *= p0.y;
        v[8 * j + 2] *= p1.x;
        v[8 * j + 3] *= p1.y;
        v[8 * j + 4] *= p2.x;
        v[8 * j + 5] *= p2.y;
        v[8 * j + 6] *= p3.x;
        v[8 * j + 7] *= p3.y;
      }
    }
    if (e.p_drop > 0.f) {
      const Philox rng(e.seed);
      const uint32_t thr = (uint32_t)(e.p_drop * 4294967296.0f);
      const float inv_keep = 1.f / (1.f - e.p_drop);
      const unsigned long long base = ((unsigned long long)row * (unsigned long long)e.drop_ld + (unsigned long long)col0) >> 2;
#pragma unroll
      for (int j = 0; j < W / 4; ++j) {
        const uint4 r = rng(base + j, e.offset);
        v[4 * j + 0] = r.x >= thr ? v[4 * j + 0] * inv_keep : 0.f;
        v[4 * j + 1] = r.y >= thr ? v[4 * j + 1] * inv_keep : 0.f;
        v[4 * j + 2] = r.z >= thr ? v[4 * j + 2] * inv_keep : 0.f;
        v[4 * j + 3] = r.w >= thr ? v[4 * j + 3] * inv_keep : 0.f;
      }
    }
    if (e.c_fp32) {
      float* crow = reinterpret_cast<float*>(e.c) + (long long)row * e.ldc + col0;
      if (e.residual) {
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
          const float4 r = res32[j];
          v[4 * j + 0] += r.x;
          v[4 * j + 1] += r.y;
          v[4 * j + 2] += r.z;
          v[4 * j + 3] += r.w;
        }
      }
      float4* c4 = reinterpret_cast<float4*>(crow);
      if (e.atomic) {
#pragma unroll
        for (int j = 0; j < W / 4; ++j) atomicAdd(c4 + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        return;
      }
      if (e.accumulate) {
        float4 old[W / 4];
#pragma unroll
        for (int j = 0; j < W / 4; ++j) old[j] = c4[j];
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
          v[4 * j + 0] += old[j].x;
          v[4 * j + 1] += old[j].y;
          v[4 * j + 2] += old[j].z;
          v[4 * j + 3] += old[j].w;
        }
      }
#pragma unroll
      for (int j = 0; j < W / 4; ++j) c4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      bf16* crow = reinterpret_cast<bf16*>(e.c) + (long long)row * e.ldc + col0;
      if (e.residual) {
#pragma unroll
        for (int j = 0; j < W / 8; ++j) {
          const uint4 u = res16[j];
          const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z),
                       p3 = unpack_bf16x2(u.w);
          v[8 * j + 0] += p0.x;
          v[8 * j + 1] += p0.y;
          v[8 * j + 2] += p1.x;
          v[8 * j + 3] += p1.y;
          v[8 * j + 4] += p2.x;
          v[8 * j + 5] += p2.y;
          v[8 * j + 6] += p3.x;
          v[8 * j + 7] += p3.y;
        }
      }
      uint4* c4 = reinterpret_cast<uint4*>(crow);
#pragma unroll
      for (int j = 0; j < W / 8; ++j) {
        if (e.accumulate) {
          const uint4 u = c4[j];
          const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z),
                       p3 = unpack_bf16x2(u.w);
          v[8 * j + 0] += p0.x;
          v[8 * j + 1] += p0.y;
          v[8 * j + 2] += p1.x;
          v[8 * j + 3] += p1.y;
          v[8 * j + 4] += p2.x;
          v[8 * j + 5] += p2.y;
          v[8 * j + 6] += p3.x;
          v[8 * j + 7] += p3.y;
        }
        uint4 u;
        u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
        u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
        u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        c4[j] = u;
      }
    }
  } else {
    // ragged last column tile: scalar, fully predicated (unrolled so that v[] stays in registers)
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int col = col0 + j;
      if (col >= N) continue;
      float x = v[j];
      if (e.bias) x += e.bias[col];
      if (e.act == 1) {
        float dg;
        x = gelu_erf_both(x, dg);
        if (e.aux_out) e.aux_out[(long long)row * e.ld_aux + col] = __float2bfloat16(dg);
      } else if (e.act == 2) {
        x *= __bfloat162float(e.aux_in[(long long)row * e.ld_aux + col]);
      }
      if (e.p_drop > 0.f) {
        const Philox rng(e.seed);
        const unsigned long long el = (unsigned long long)row * (unsigned long long)e.drop_ld + (unsigned long long)col;
        const uint4 r = rng(el >> 2, e.offset);
        const uint32_t w = (el & 3) == 0 ? r.x : ((el & 3) == 1 ? r.y : ((el & 3) == 2 ? r.z : r.w));
        x = w >= (uint32_t)(e.p_drop * 4294967296.0f) ? x / (1.f - e.p_drop) : 0.f;
      }
      if (e.c_fp32) {
        float* c = reinterpret_cast<float*>(e.c) + (long long)row * e.ldc + col;
        if (e.residual) x += reinterpret_cast<const float*>(e.residual)[(long long)row * e.ldr + col];
        if (e.atomic) {
          atomicAdd(c, x);
          continue;
        }
        if (e.accumulate) x += *c;
        *c = x;
      } else {
        bf16* c = reinterpret_cast<bf16*>(e.c) + (long long)row * e.ldc + col;
        if (e.residual) x += __bfloat162float(reinterpret_cast<const bf16*>(e.residual)[(long long)row * e.ldr + col]);
        if (e.accumulate) x += __bfloat162float(*c);
        *c = __float2bfloat16(x);
      }
    }
  }
}


// ---- staged fast path -------------------------------------------------------------------------------------------------
// One 32-column span of one accumulator row per lane, bf16 C without accumulation, interior of the N range (col0 + 32 <= N).
// `pre` holds the row's 32 bf16 inputs of this span (the stashed GELU' if act == 2, else the residual) loaded one tile ahead, so
// no global-load latency is exposed here; the bias of the tile sits in shared memory (staged by the epilogue warps while the
// mainloop of the tile runs).  The result (and, for act == 1 with aux_out, GELU' of the pre-activation) leaves through one of the
// warp's BUFS 32x32 bf16 staging tiles (dense 64-byte rows, SWIZZLE_64B: 16-byte unit ^= (row >> 1) & 3 — conflict-free for
// row-per-lane 16-byte accesses) and a TMA tensor store issued by lane 0; rows >= M are clipped by the tensor map.  With
// BUFS == 2 the tiles alternate, so a tile is only rewritten after the store issued two stores ago has drained it (ncu showed
// 18 % of the epilogue time waiting for the drain of a single tile, profiles/ncu_r1b_attention_gemm.txt).
// Must be called by all 32 lanes.
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_3d_s(const void* desc, uint32_t smem_addr, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_addr), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// Epilogue modes the staged fast path is specialised for at compile time (one kernel instantiation each: the generic
// runtime-flag epilogue needs ~150 live registers, the specialised ones fit the 128-register budget of 14 warps / CTA).
enum { EPI_GENERIC = 0, EPI_BIAS = 1, EPI_GELU = 2, EPI_GELUGRAD = 3, EPI_RESID = 4 };

// claims the warp's next staging tile: waits until the TMA store that last read it has drained, returns its smem address
template <int BUFS>
__device__ __forceinline__ uint32_t stg_acquire(uint32_t stg_base, uint32_t& stg_cnt, int lane) {
  if (lane == 0) bulk_wait_read<BUFS - 1>();
  __syncwarp();
  const uint32_t a = stg_base + (stg_cnt & (uint32_t)(BUFS - 1)) * 2048u;
  ++stg_cnt;
  return a;
}

// Row inputs of a span (the stashed GELU' if MODE == EPI_GELUGRAD, else the residual): 32 rows x 64 bytes.  They are requested ONE
// SPAN AHEAD with the coalesced mapping "lane l holds the 16-byte unit (l & 3) of rows 8 j + (l >> 2), j = 0..3" (one request touches
// 8 whole 64-byte row segments instead of 32 quarter-used sectors), kept in 16 registers while the current span is processed, and
// scattered into the warp's staging tile at the start of their own span, where every lane then reads back its own row.
// (Round 1 kept a whole tile of row-per-lane inputs in an array filled through a lambda: ptxas placed it in local memory, so every
// "prefetch" stalled on its own STL — profiles/ncu_r2a_gemm_out_resid.txt: the residual epilogue cost 14 of 33 us on 12608x768x768.)
struct PreReq {
  const bf16* base;     // element (row0, col0) of the span, nullptr = no request
  long long ld;
  int rows;             // valid rows from row0 on (M - row0), may exceed 32
  int cols;             // valid columns from col0 on (N - col0), may exceed 32; 16-byte units starting at or beyond it are not read
};
__device__ __forceinline__ void pre_issue(uint4 (&pre)[4], const PreReq& rq, int lane) {
  if (rq.base == nullptr) return;
  const int r = lane >> 2, c = lane & 3;
  if (8 * c >= rq.cols) return;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (8 * j + r < rq.rows) pre[j] = __ldg(reinterpret_cast<const uint4*>(rq.base + (long long)(8 * j + r) * rq.ld) + c);
}

// `next`: request for the span this warp processes after this one (issued right after this span's own inputs left the registers).
// `e` is the kernel parameter itself (constant bank): nothing of it is copied into registers ahead of use.
template <int MODE, int BUFS>
__device__ __forceinline__ void epilogue_span_fast(uint32_t taddr, int row0, int col0, int batch_idx, int lane, const GemmEpilogue& e,
                                                   float alpha, unsigned long long drop_offset, bool use_pre, uint4 (&pre)[4],
                                                   const PreReq& next, uint32_t stg_base, uint32_t& stg_cnt,
                                                   uint32_t sbias /* smem address of the span's 32 bias floats, 0 = none */,
                                                   const CUtensorMap* tmap_c, const CUtensorMap* tmap_aux) {
  constexpr bool HAS_PRE = MODE == EPI_GELUGRAD || MODE == EPI_RESID;
  const int row = row0 + lane;
  const bool stash = MODE == EPI_GELU && e.aux_out != nullptr;
  uint32_t stashed[MODE == EPI_GELU ? 16 : 1];
  const uint32_t stg = stg_acquire<BUFS>(stg_base, stg_cnt, lane);
  const uint32_t srow = stg + (uint32_t)lane * 64u;
  const int sw = (lane >> 1) & 3;
  if constexpr (HAS_PRE) {
    if (use_pre) {
      const int r = lane >> 2, c = lane & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = 8 * j + r;
        sts128(stg + (uint32_t)(rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4)), pre[j].x, pre[j].y, pre[j].z, pre[j].w);
      }
      __syncwarp();
      pre_issue(pre, next, lane);
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {                 // two 16-column halves
    uint32_t acc[16];
    tmem_ld16(taddr + 16 * h, acc);
    uint4 pa = make_uint4(0, 0, 0, 0), pb = make_uint4(0, 0, 0, 0);
    if constexpr (HAS_PRE) {
      if (use_pre) {
        pa = lds128u(srow + (uint32_t)(((2 * h) ^ sw) << 4));
        pb = lds128u(srow + (uint32_t)(((2 * h + 1) ^ sw) << 4));
      }
    }
    tmem_ld_wait();
    float v[16];
    if (sbias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b = lds128f(sbias + (uint32_t)(64 * h + 16 * j));
        v[4 * j + 0] = fmaf(__uint_as_float(acc[4 * j + 0]), alpha, b.x);
        v[4 * j + 1] = fmaf(__uint_as_float(acc[4 * j + 1]), alpha, b.y);
        v[4 * j + 2] = fmaf(__uint_as_float(acc[4 * j + 2]), alpha, b.z);
        v[4 * j + 3] = fmaf(__uint_as_float(acc[4 * j + 3]), alpha, b.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]) * alpha;
    }
    if constexpr (MODE == EPI_GELU) {
      if (stash) {                              // GELU and GELU' share every transcendental; the derivative is what is stashed
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float d0, d1;
#if defined(VLM_GELU_F32X2) && VLM_GELU_F32X2
          gelu_erf_both_x2(v[2 * j], v[2 * j + 1], v[2 * j], v[2 * j + 1], d0, d1);
#else
          v[2 * j] = gelu_erf_both(v[2 * j], d0);
          v[2 * j + 1] = gelu_erf_both(v[2 * j + 1], d1);
#endif
          stashed[8 * h + j] = pack_bf16x2(d0, d1);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
      }
    }
    if constexpr (HAS_PRE) {
      const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
      if constexpr (MODE == EPI_GELUGRAD) {   // aux_in holds GELU'(pre-activation), stashed by the forward epilogue
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 x = unpack_bf16x2(pw[j]);
          v[2 * j] *= x.x;
          v[2 * j + 1] *= x.y;
        }
      } else {
        if (e.p_drop > 0.f) {
          const Philox rng(e.seed);
          const uint32_t thr = (uint32_t)(e.p_drop * 4294967296.0f);
          const float inv_keep = 1.f / (1.f - e.p_drop);
          const unsigned long long base =
              ((unsigned long long)row * (unsigned long long)e.drop_ld + (unsigned long long)(col0 + 16 * h)) >> 2;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 r = rng(base + j, drop_offset);
            v[4 * j + 0] = r.x >= thr ? v[4 * j + 0] * inv_keep : 0.f;
            v[4 * j + 1] = r.y >= thr ? v[4 * j + 1] * inv_keep : 0.f;
            v[4 * j + 2] = r.z >= thr ? v[4 * j + 2] * inv_keep : 0.f;
            v[4 * j + 3] = r.w >= thr ? v[4 * j + 3] * inv_keep : 0.f;
          }
        }
        if (use_pre) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 x = unpack_bf16x2(pw[j]);
            v[2 * j] += x.x;
            v[2 * j + 1] += x.y;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      sts128(srow + (uint32_t)(((2 * h + u) ^ sw) << 4), pack_bf16x2(v[8 * u + 0], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
             pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_3d_s(tmap_c, stg, col0, row0, batch_idx);
    bulk_commit_group();
  }
  if constexpr (MODE == EPI_GELU) {
    if (stash) {                                // the GELU pre-activation leaves through the warp's next staging tile
      const uint32_t stg2 = stg_acquire<BUFS>(stg_base, stg_cnt, lane);
      const uint32_t srow2 = stg2 + (uint32_t)lane * 64u;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        sts128(srow2 + (uint32_t)((u ^ sw) << 4), stashed[4 * u], stashed[4 * u + 1], stashed[4 * u + 2], stashed[4 * u + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d_s(tmap_aux, stg2, col0, row0, batch_idx);
        bulk_commit_group();
      }
    }
  }
}

// host-side pieces shared by the GEMM translation units
int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t rows, uint64_t batch,
                   uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_rows);
int make_tmap_bf16_nd(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                      const uint32_t* box, int swizzle_bytes = 128);
void bind_context_for_driver_calls();

}  // namespace vlm
