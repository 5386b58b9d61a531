// tcgen05 attention BACKWARD for head dim 64 and Tq <= 256 (ViT S=197, decoder T=128, cross 128 x 197[*N]).
//
// One persistent CTA per SM walks over (batch, head) items; for every 128-key tile j of the item it runs the five
// products of the attention backward on the 5th-gen tensor cores with all accumulators in TMEM, in the TRANSPOSED
// formulation (rows = keys), so that P^T / dS^T are directly the A operands of the dV / dK products and dS (the A
// operand of dQ) is the same shared-memory tile read through an MN-major descriptor:
//     S^T  = K_j Q^T                 (M=128 keys, N=Nq, K=64)    TMEM cols [0, Nq)
//     dP^T = V_j dO^T                (same shape, same columns, after P^T has been extracted)
//     dV_j = P^T dO                  (M=128, N=64, K=Nq)         TMEM cols [256, 320)
//     dK_j = dS^T Q                  (M=128, N=64, K=Nq)         TMEM cols [320, 384)
//     dQ  += dS K_j                  (M=128 per q block, N=64, K=128 keys)  TMEM cols [384, 512), accumulated over j
// Warp roles: warp0 = TMA producer (4-D tensor maps straight over the packed QKV / dO layouts), warp1 = MMA issuer,
// warps 2..9 = 256 "softmax" threads (thread <-> key row, two warps per TMEM lane quadrant split the query columns):
//   pass A:  P^T = exp2(S^T*scale*log2e - lse2[q]) with key-padding / causal masks (+ Philox dropout) -> bf16, written
//            to shared memory in the canonical K-major SWIZZLE_128B layout the UMMA descriptors expect;
//   pass B:  dS^T = P^T * (dP^T - delta[q]) * scale, in place;
//   epilogue: dV_j, dK_j (and dQ after the last key tile) TMEM -> bf16 -> global.
// delta = rowsum(dO * O) (computed in the kernel from the O rows in global memory and the dO tile TMA has just loaded) and lse2 are
// staged per item by the same threads.  Same math / same dropout stream as the
// mma.sync kernels in attention.cu (those remain the path for other head dims, longer queries and the forward).
#include <cuda.h>
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "vlm_b200.h"

namespace vlm {

struct AttnTcParams {
  const bf16* o; const bf16* d_o;
  long long o_bs, o_rs, do_bs, do_rs;
  bf16* dq; bf16* dk; bf16* dv;
  long long dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs;
  const float* lse;        // [B,H,Tq] natural log
  const float* delta;      // unused since round 2 (delta = rowsum(dO * O) is computed per item inside the kernel)
  const uint8_t* kmask;    // [B,Sk] or null
  int B, H, Tq, Sk, Nq;    // Nq = Tq rounded up to 32
  int causal;
  float scale, p_drop;
  unsigned long long seed, offset;
  const unsigned long long* offset_ptr;
};

__device__ __forceinline__ void tma_load_4d(const void* desc, uint64_t* bar, void* smem_dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// calls f(std::integral_constant<int, n>) for the runtime n in [0, MAXN] — a compare chain, used once per pass and tile
template <int MAXN, typename F>
__device__ __forceinline__ void dispatch_count(int n, F&& f) {
  if constexpr (MAXN == 0) {
    f(std::integral_constant<int, 0>{});
  } else {
    if (n >= MAXN) f(std::integral_constant<int, MAXN>{});
    else dispatch_count<MAXN - 1>(n, f);
  }
}

static constexpr int ATC_SPLIT = 4;                              // compute warps per TMEM lane quadrant (each owns Nq / 4 query columns)
static constexpr int ATC_THREADS = 64 + 128 * ATC_SPLIT;         // 2 control warps + 16 compute warps
static constexpr int ATC_S_COL = 0, ATC_DV_COL = 256, ATC_DK_COL = 320, ATC_DQ_COL = 384;

template <int ATOMS, bool DROPOUT>
struct AtcSmem {
  static constexpr int NQ_MAX = ATOMS * 64;
  static constexpr int Q_BYTES = NQ_MAX * 128;
  static constexpr int KV_BYTES = 128 * 128;
  static constexpr int PT_BYTES = ATOMS * 16384;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_DO = OFF_Q + Q_BYTES;
  static constexpr int OFF_K = OFF_DO + Q_BYTES;           // 2 buffers
  static constexpr int OFF_V = OFF_K + 2 * KV_BYTES;       // 2 buffers
  static constexpr int OFF_PT = OFF_V + 2 * KV_BYTES;
  static constexpr int OFF_P2 = OFF_PT + PT_BYTES;
  static constexpr int OFF_LSE = OFF_P2 + (DROPOUT ? PT_BYTES : 0);
  static constexpr int OFF_DELTA = OFF_LSE + NQ_MAX * 4;
  static constexpr int OFF_BAR = OFF_DELTA + NQ_MAX * 4;
  static constexpr int NUM_BARS = 13;
  static constexpr int TOTAL = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
};

// byte offset of element (row r, column q) inside a K-major SWIZZLE_128B tile made of 64-column atoms of 128 rows
__device__ __forceinline__ uint32_t pt_offset16(int r, int q0) {  // q0 multiple of 8: start of a 16-byte unit
  return (uint32_t)((q0 >> 6) * 16384 + r * 128 + ((((q0 & 63) >> 3) ^ (r & 7)) << 4));
}

template <int ATOMS, bool DROPOUT>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do, AttnTcParams p) {
  using S = AtcSmem<ATOMS, DROPOUT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + S::OFF_Q;
  uint8_t* sDO = smem + S::OFF_DO;
  uint8_t* sK = smem + S::OFF_K;
  uint8_t* sV = smem + S::OFF_V;
  uint8_t* sPT = smem + S::OFF_PT;
  uint8_t* sP2 = smem + S::OFF_P2;
  float* sLse = reinterpret_cast<float*>(smem + S::OFF_LSE);
  float* sDelta = reinterpret_cast<float*>(smem + S::OFF_DELTA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* qdo_full = bars + 0;
  uint64_t* qdo_empty = bars + 1;
  uint64_t* kv_full = bars + 2;    // [2]
  uint64_t* kv_empty = bars + 4;   // [2]
  uint64_t* s_full = bars + 6;
  uint64_t* p_ready = bars + 7;
  uint64_t* dp_full = bars + 8;
  uint64_t* dv_done = bars + 9;
  uint64_t* ds_ready = bars + 10;
  uint64_t* out_full = bars + 11;
  uint64_t* out_free = bars + 12;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + S::NUM_BARS);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = p.B * p.H;
  const int ntiles = (p.Sk + 127) / 128;
  const int Nq = p.Nq;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    mbar_init(qdo_full, 1);
    mbar_init(qdo_empty, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1);
    mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4 * ATC_SPLIT);
    mbar_init(dp_full, 1);
    mbar_init(dv_done, 1);
    mbar_init(ds_ready, 4 * ATC_SPLIT);
    mbar_init(out_full, 1);
    mbar_init(out_free, 4 * ATC_SPLIT);
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc<512>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t item_cnt = 0, tile_cnt = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++item_cnt) {
        const int b = item / p.H, h = item % p.H;
        mbar_wait(qdo_empty, (item_cnt & 1u) ^ 1u);
        mbar_expect_tx(qdo_full, 2u * (uint32_t)Nq * 128u);
        tma_load_4d(&tm_q, qdo_full, sQ, 0, h, 0, b);
        tma_load_4d(&tm_do, qdo_full, sDO, 0, h, 0, b);
        for (int j = 0; j < ntiles; ++j, ++tile_cnt) {
          const int buf = tile_cnt & 1u;
          const uint32_t ph = (tile_cnt >> 1) & 1u;
          mbar_wait(&kv_empty[buf], ph ^ 1u);
          mbar_expect_tx(&kv_full[buf], 2u * S::KV_BYTES);
          tma_load_4d(&tm_k, &kv_full[buf], sK + buf * S::KV_BYTES, 0, h, j * 128, b);
          tma_load_4d(&tm_v, &kv_full[buf], sV + buf * S::KV_BYTES, 0, h, j * 128, b);
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================================== MMA issuer
    // whole warp runs the loop (uniform datapath for descriptors / barriers), one elected lane issues the tcgen05 ops
    {
      const bool leader = elect_one();
      const uint32_t id_s = make_idesc_bf16(128, Nq, false, false);     // S^T, dP^T : K-major x K-major
      const uint32_t id_dv = make_idesc_bf16(128, 64, false, true);      // dV, dK   : K-major A, MN-major B
      const uint32_t id_dq = make_idesc_bf16(128, 64, true, true);       // dQ       : MN-major A and B
      const uint32_t aQ = smem_u32(sQ), aDO = smem_u32(sDO), aPT = smem_u32(sPT);
      const int nq16 = Nq / 16;
      const int q_blocks = (Nq + 127) / 128;
      uint32_t item_cnt = 0, tile_cnt = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++item_cnt) {
        mbar_wait(qdo_full, item_cnt & 1u);
        tc_fence_after();
        for (int j = 0; j < ntiles; ++j, ++tile_cnt) {
          const int buf = tile_cnt & 1u;
          const uint32_t kvph = (tile_cnt >> 1) & 1u, tph = tile_cnt & 1u;
          const uint32_t aK = smem_u32(sK + buf * S::KV_BYTES), aV = smem_u32(sV + buf * S::KV_BYTES);
          mbar_wait(&kv_full[buf], kvph);
          tc_fence_after();
          // (1) S^T = K_j Q^T
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + ATC_S_COL, make_smem_desc(aK + k * 32, 16, 1024), make_smem_desc(aQ + k * 32, 16, 1024), id_s, k > 0);
            umma_commit(s_full);
          }
          __syncwarp();
          mbar_wait(p_ready, tph);
          tc_fence_after();
          // (2) dP^T = V_j dO^T  (same TMEM columns; P^T has been extracted)
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + ATC_S_COL, make_smem_desc(aV + k * 32, 16, 1024), make_smem_desc(aDO + k * 32, 16, 1024), id_s, k > 0);
            umma_commit(dp_full);
          }
          __syncwarp();
          mbar_wait(out_free, tph ^ 1u);   // dV / dK / dQ accumulators of the previous tile have been drained
          tc_fence_after();
          // (3) dV_j = P^T dO
          if (leader) {
            for (int k = 0; k < nq16; ++k)
              umma_bf16(tmem_base + ATC_DV_COL, make_smem_desc(aPT + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                        make_smem_desc(aDO + k * 2048, 16384, 1024), id_dv, k > 0);
            umma_commit(dv_done);
          }
          __syncwarp();
          mbar_wait(ds_ready, tph);
          tc_fence_after();
          if (leader) {
            // (4) dK_j = dS^T Q
            for (int k = 0; k < nq16; ++k)
              umma_bf16(tmem_base + ATC_DK_COL, make_smem_desc(aPT + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                        make_smem_desc(aQ + k * 2048, 16384, 1024), id_dv, k > 0);
            // (5) dQ[mb] += dS K_j   (dS = the dS^T tile through an MN-major descriptor: 64-query blocks 16 KB apart)
            for (int mb = 0; mb < q_blocks; ++mb)
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_bf16(tmem_base + ATC_DQ_COL + mb * 64, make_smem_desc(aPT + mb * 32768 + k * 2048, 16384, 1024),
                          make_smem_desc(aK + k * 2048, 16384, 1024), id_dq, (j > 0 || k > 0) ? 1u : 0u);
            umma_commit(out_full);
            umma_commit(&kv_empty[buf]);
            if (j == ntiles - 1) umma_commit(qdo_empty);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================================================== compute warps (128 * ATC_SPLIT threads)
    // Each softmax-side thread runs one long dependent chain per chunk (TMEM load -> ex2 -> pack -> st.shared); with two warps per
    // scheduler the issue slots sat idle ~90 % of the time (ncu, round 2), so every lane quadrant is shared by ATC_SPLIT warps.
    const int quad = warp_idx & 3;             // TMEM lane quadrant
    const int part = (warp_idx - 2) >> 2;      // which slice of the query columns / output columns
    const int r = quad * 32 + lane;            // key row inside the tile
    const int ct = threadIdx.x - 64;           // 0 .. 128 * ATC_SPLIT - 1
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    const unsigned long long off_eff = p.offset + ((DROPOUT && p.offset_ptr) ? __ldg(p.offset_ptr) : 0ull);
    const uint32_t thr = (uint32_t)(p.p_drop * 4294967296.0f);
    const float inv_keep = DROPOUT ? 1.f / (1.f - p.p_drop) : 1.f;
    const uint32_t pt_row = (uint32_t)(r * 128);   // row offset inside a 64-column atom of the P^T / dS^T tile
    const uint32_t aPTs = smem_u32(sPT), aP2s = smem_u32(sP2), aLse = smem_u32(sLse), aDelta = smem_u32(sDelta);   // shared-space addresses
    const int r7 = r & 7;
    uint32_t tile_cnt = 0, item_cnt = 0;
    const uint32_t aDOs = smem_u32(sDO);
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++item_cnt) {
      const int b = item / p.H, h = item % p.H;
      const uint32_t dkey = attn_drop_key(p.seed, off_eff, item);
      // ---- per-item staging: lse (log2 units, +inf beyond Tq) and delta = rowsum(dO * O).
      // delta is computed HERE (round 2; it used to be a separate pre-pass kernel per launch: 36 launches, 0.32 ms / step and a second
      // read of dO): two threads per query row, each multiplies four 16-byte units of the O row (global) with the same units of
      // the dO row that TMA has just put into shared memory (swizzled K-major rows of 128 bytes).  The O rows of the CTA's NEXT
      // item are requested into L2 now, so that this read is an L2 hit next time.
      named_bar_sync(1, 128 * ATC_SPLIT);
      {
        const int qrow = ct >> 1, half = ct & 1;
        uint4 ov[4];
        const bool live = qrow < p.Tq;
        if (live) {
          const uint4* op = reinterpret_cast<const uint4*>(p.o + (long long)b * p.o_bs + (long long)qrow * p.o_rs + h * 64) + half * 4;
#pragma unroll
          for (int u = 0; u < 4; ++u) ov[u] = __ldg(op + u);
          const int nitem = item + gridDim.x;
          if (nitem < nitems) {
            const bf16* np = p.o + (long long)(nitem / p.H) * p.o_bs + (long long)qrow * p.o_rs + (nitem % p.H) * 64 + half * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(np));
          }
        }
        if (ct < Nq) sLse[ct] = ct < p.Tq ? p.lse[(long long)item * p.Tq + ct] * 1.4426950408889634f : INFINITY;
        mbar_wait(qdo_full, item_cnt & 1u);        // dO (and Q) of this item are in shared memory
        float acc = 0.f;
        if (live) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint4 dv = lds128u(aDOs + (uint32_t)(qrow * 128) + (uint32_t)((((half << 2) + u) ^ (qrow & 7)) << 4));
            const uint32_t ow[4] = {ov[u].x, ov[u].y, ov[u].z, ov[u].w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 a = unpack_bf16x2(ow[e]), g = unpack_bf16x2(dw[e]);
              acc = fmaf(a.x, g.x, acc);
              acc = fmaf(a.y, g.y, acc);
            }
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (half == 0 && qrow < Nq) sDelta[qrow] = acc;      // rows in [Tq, Nq): 0
      }
      named_bar_sync(1, 128 * ATC_SPLIT);
      for (int j = 0; j < ntiles; ++j, ++tile_cnt) {
        const uint32_t tph = tile_cnt & 1u;
        const int kk = j * 128 + r;
        const bool kvalid = (kk < p.Sk) && (!p.kmask || p.kmask[(long long)b * p.Sk + kk]);
        // ---------------- pass A: P^T
        mbar_wait(s_full, tph);
        tc_fence_after();
        // Query columns are dealt to the ATC_SPLIT warps of a quadrant in 8-query chunks, round robin: chunk c of this thread covers
        // queries [32 c + 8 part, +8) = one 16-byte unit of the P^T row.  causal: query q sees key kk iff kk <= q, so for this warp's
        // key rows [k_lo, k_lo + 31] chunk c is fully masked iff q0 + 7 < k_lo (c < c_any), fully visible iff q0 >= k_lo + 31
        // (c >= c_full) and tests per element only on the diagonal in between.  The fully visible chunks run straight-line code
        // selected by their COUNT (compile-time offsets relative to the first of them).  Key rows that are padding / masked
        // (kvalid == false) produce zeros; warps whose 32 key rows are all padding only write zeros.
        const int k_lo = j * 128 + quad * 32;
        const int nch = Nq >> 5;                 // chunks per thread
        int c_any = 0, c_full = 0;
        if (p.causal) {
          c_any = (k_lo - 7 - part * 8 + 31) >> 5;             // first c with 32 c + 8 part + 7 >= k_lo
          c_full = (k_lo + 31 - part * 8 + 31) >> 5;           // first c with 32 c + 8 part >= k_lo + 31
          c_any = c_any < 0 ? 0 : (c_any > nch ? nch : c_any);
          c_full = c_full < 0 ? 0 : (c_full > nch ? nch : c_full);
        }
        const bool any_valid = __any_sync(0xffffffffu, kvalid);
        // offsets of chunk c inside the P^T / dS^T tile: 64-query atom c >> 1, 16-byte unit (c & 1) * 4 + part, XOR-swizzled
        auto pt_off = [&](int c) -> uint32_t {
          return (uint32_t)((c >> 1) * 16384) + pt_row + (uint32_t)(((((c & 1) << 2) + part) ^ r7) << 4);
        };
        const uint32_t s_col = lane_taddr + ATC_S_COL + (uint32_t)(part * 8);
        uint32_t va[8], vb[8];
        auto store_p = [&](float (&pv)[8], int q0, uint32_t off) {
          if (DROPOUT) {
            float pd[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pd[i] = attn_drop_rand(dkey, q0 + i, kk, p.Sk) >= thr ? pv[i] * inv_keep : 0.f;
            sts128(aPTs + off, pack_bf16x2(pd[0], pd[1]), pack_bf16x2(pd[2], pd[3]), pack_bf16x2(pd[4], pd[5]), pack_bf16x2(pd[6], pd[7]));
          }
          sts128((DROPOUT ? aP2s : aPTs) + off, pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]),
                 pack_bf16x2(pv[6], pv[7]));
        };
        if (any_valid) {
          // ---- masked and diagonal chunks (causal only)
#pragma unroll 1
          for (int c = 0; c < c_full; ++c) {
            const int q0 = c * 32 + part * 8;
            const uint32_t off = pt_off(c);
            float pv[8];
            if (c < c_any) {
#pragma unroll
              for (int i = 0; i < 8; ++i) pv[i] = 0.f;
              sts128(aPTs + off, 0u, 0u, 0u, 0u);
              if (DROPOUT) sts128(aP2s + off, 0u, 0u, 0u, 0u);
              continue;
            }
            tmem_ld8(s_col + (uint32_t)(c * 32), va);
            const float4 l0 = lds128f(aLse + (uint32_t)(q0 * 4)), l1 = lds128f(aLse + (uint32_t)(q0 * 4 + 16));
            const float ls[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i)
              pv[i] = (kvalid && kk <= q0 + i) ? ex2_approx(fmaf(__uint_as_float(va[i]), sl2, -ls[i])) : 0.f;
            store_p(pv, q0, off);
          }
          // ---- fully visible chunks [c_full, nch): straight-line, TMEM loads two deep
          const uint32_t colb = s_col + (uint32_t)(c_full * 32);
          const uint32_t lseb = aLse + (uint32_t)((c_full * 32 + part * 8) * 4);
          const int q0b = c_full * 32 + part * 8;
          const uint32_t off0 = pt_off(c_full), off1 = pt_off(c_full + 1);
          dispatch_count<ATOMS * 2>(nch - c_full, [&](auto nconst) {
            constexpr int N = decltype(nconst)::value;
            auto body = [&](const uint32_t* v, int i) {
              const float4 l0 = lds128f(lseb + (uint32_t)(i * 128)), l1 = lds128f(lseb + (uint32_t)(i * 128 + 16));
              const float ls[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
              float pv[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) pv[e] = kvalid ? ex2_approx(fmaf(__uint_as_float(v[e]), sl2, -ls[e])) : 0.f;
              store_p(pv, q0b + i * 32, ((i & 1) ? off1 : off0) + (uint32_t)((i >> 1) * 16384));
            };
            if (N > 0) tmem_ld8(colb, va);
#pragma unroll
            for (int i = 0; i < N; i += 2) {
              tmem_ld_wait();
              if (i + 1 < N) tmem_ld8(colb + (uint32_t)((i + 1) * 32), vb);
              body(va, i);
              if (i + 1 < N) {
                tmem_ld_wait();
                if (i + 2 < N) tmem_ld8(colb + (uint32_t)((i + 2) * 32), va);
                body(vb, i + 1);
              }
            }
          });
        } else {
          for (int c = 0; c < nch; ++c) {
            const uint32_t off = pt_off(c);
            sts128(aPTs + off, 0u, 0u, 0u, 0u);          // P^T = 0  =>  dS^T = 0 as well: pass B leaves these rows alone
            if (DROPOUT) sts128(aP2s + off, 0u, 0u, 0u, 0u);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        // ---------------- pass B: dS^T (in place over P^T)
        mbar_wait(dp_full, tph);
        mbar_wait(dv_done, tph);
        tc_fence_after();
        // dS^T = P^T * (dP^T - delta) — the softmax scale is applied once per dK / dQ output element in the epilogue instead
        // of once per score element here.
        // Chunks [0, c_any) hold P^T = 0 (fully masked): dS^T = 0 is already there.  The rest is one straight-line sequence
        // selected by its length.
        if (any_valid) {
          const uint32_t colb = s_col + (uint32_t)(c_any * 32);
          const uint32_t delb = aDelta + (uint32_t)((c_any * 32 + part * 8) * 4);
          const uint32_t off0 = pt_off(c_any), off1 = pt_off(c_any + 1);
          dispatch_count<ATOMS * 2>(nch - c_any, [&](auto nconst) {
            constexpr int N = decltype(nconst)::value;
            auto body = [&](const uint32_t* v, int i) {
              const float4 d0 = lds128f(delb + (uint32_t)(i * 128)), d1 = lds128f(delb + (uint32_t)(i * 128 + 16));
              const float dl[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
              const uint32_t off = ((i & 1) ? off1 : off0) + (uint32_t)((i >> 1) * 16384);
              const uint4 pw = lds128u((DROPOUT ? aP2s : aPTs) + off);
              uint4 kw = pw;
              if (DROPOUT) kw = lds128u(aPTs + off);   // dropped P: zero <=> dropped (or P == 0)
              const uint32_t pw4[4] = {pw.x, pw.y, pw.z, pw.w}, kw4[4] = {kw.x, kw.y, kw.z, kw.w};
              uint32_t ow[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 pp = unpack_bf16x2(pw4[e]), kp = unpack_bf16x2(kw4[e]);
                float x0 = __uint_as_float(v[2 * e]), x1 = __uint_as_float(v[2 * e + 1]);
                if (DROPOUT) {
                  x0 = kp.x != 0.f ? x0 * inv_keep : 0.f;
                  x1 = kp.y != 0.f ? x1 * inv_keep : 0.f;
                }
                ow[e] = pack_bf16x2(pp.x * (x0 - dl[2 * e]), pp.y * (x1 - dl[2 * e + 1]));
              }
              sts128(aPTs + off, ow[0], ow[1], ow[2], ow[3]);
            };
            if (N > 0) tmem_ld8(colb, va);
#pragma unroll
            for (int i = 0; i < N; i += 2) {
              tmem_ld_wait();
              if (i + 1 < N) tmem_ld8(colb + (uint32_t)((i + 1) * 32), vb);
              body(va, i);
              if (i + 1 < N) {
                tmem_ld_wait();
                if (i + 2 < N) tmem_ld8(colb + (uint32_t)((i + 2) * 32), va);
                body(vb, i + 1);
              }
            }
          });
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_ready);
        // ---------------- epilogue: dV_j, dK_j (+ dQ after the last key tile)
        mbar_wait(out_full, tph);
        tc_fence_after();
        {
          constexpr int OC = 64 / ATC_SPLIT;     // output columns per part (16)
          uint32_t acc[OC];
          auto store16 = [&](bf16* base, float sc) {
            uint4* dst = reinterpret_cast<uint4*>(base);
#pragma unroll
            for (int i = 0; i < OC / 8; ++i)
              dst[i] = make_uint4(pack_bf16x2(__uint_as_float(acc[8 * i]) * sc, __uint_as_float(acc[8 * i + 1]) * sc),
                                  pack_bf16x2(__uint_as_float(acc[8 * i + 2]) * sc, __uint_as_float(acc[8 * i + 3]) * sc),
                                  pack_bf16x2(__uint_as_float(acc[8 * i + 4]) * sc, __uint_as_float(acc[8 * i + 5]) * sc),
                                  pack_bf16x2(__uint_as_float(acc[8 * i + 6]) * sc, __uint_as_float(acc[8 * i + 7]) * sc));
          };
          tmem_ld16(lane_taddr + ATC_DV_COL + part * OC, acc);
          tmem_ld_wait();
          if (kk < p.Sk) store16(p.dv + (long long)b * p.dv_bs + (long long)kk * p.dv_rs + h * 64 + part * OC, 1.f);
          const float sc = p.scale;          // dS was left unscaled (pass B)
          tmem_ld16(lane_taddr + ATC_DK_COL + part * OC, acc);
          tmem_ld_wait();
          if (kk < p.Sk) store16(p.dk + (long long)b * p.dk_bs + (long long)kk * p.dk_rs + h * 64 + part * OC, sc);
          if (j == ntiles - 1) {
            const int q_blocks = (Nq + 127) / 128;
            for (int mb = 0; mb < q_blocks; ++mb) {
              tmem_ld16(lane_taddr + ATC_DQ_COL + mb * 64 + part * OC, acc);
              tmem_ld_wait();
              const int q = mb * 128 + r;
              if (q < p.Tq) store16(p.dq + (long long)b * p.dq_bs + (long long)q * p.dq_rs + h * 64 + part * OC, sc);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(out_free);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


template <int ATOMS, bool DROPOUT>
static int launch_attn_bwd_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                              const AttnTcParams& p, cudaStream_t stream) {
  using S = AtcSmem<ATOMS, DROPOUT>;
  auto kern = attn_bwd_tc_kernel<ATOMS, DROPOUT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (err != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_bwd_tc smem=%d): %s", S::TOTAL, cudaGetErrorString(err));
      return -1;
    }
    attr_set = true;
  }
  const int items = p.B * p.H;
  const int grid = items < num_sms() ? items : num_sms();
  kern<<<grid, ATC_THREADS, S::TOTAL, stream>>>(tq, tk, tv, tdo, p);
  return check_launch("attn_bwd_tc");
}



// returns 1 if the shape is handled by the tcgen05 kernel (and it was launched), 0 if not supported, <0 on error
int attention_bwd_tc_dispatch(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                              const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                              const void* d_o, long long do_bs, long long do_rs, const float* lse, float* delta, void* dq,
                              long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs, void* dv, long long dv_bs,
                              long long dv_rs, const uint8_t* kmask, int B, int H, int Tq, int Sk, int DH, int causal,
                              float scale, float p_drop, unsigned long long seed, unsigned long long offset,
                              const unsigned long long* rng_offset_ptr, cudaStream_t stream) {
  if (DH != 64 || Tq > 256 || (p_drop > 0.f && Tq > 128)) return 0;
  if ((o_rs % 8) || (do_rs % 8) || (dq_rs % 8) || (dk_rs % 8) || (dv_rs % 8) || (o_bs % 8) || (do_bs % 8) || (dq_bs % 8) ||
      (dk_bs % 8) || (dv_bs % 8))
    return 0;  // 16-byte vector access on the row pointers
  bind_context_for_driver_calls();
  const int Nq = (Tq + 31) / 32 * 32;
  CUtensorMap tq, tk, tv, tdo;
  auto mk = [&](CUtensorMap* tm, const void* ptr, long long bs, long long rs, int rows, int box_rows) {
    const uint64_t dims[4] = {64, (uint64_t)H, (uint64_t)rows, (uint64_t)B};
    const uint64_t strides[3] = {64, (uint64_t)rs, (uint64_t)bs};
    const uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
    return make_tmap_bf16_nd(tm, ptr, 4, dims, strides, box);
  };
  if (mk(&tq, q, q_bs, q_rs, Tq, Nq) || mk(&tdo, d_o, do_bs, do_rs, Tq, Nq) || mk(&tk, k, k_bs, k_rs, Sk, 128) ||
      mk(&tv, v, v_bs, v_rs, Sk, 128))
    return -1;
  // (delta = rowsum(dO * O) is computed inside the kernel since round 2; `delta` stays in the signature for the mma.sync path)
  (void)delta;
  AttnTcParams p;
  p.delta = nullptr;
  p.o = (const bf16*)o; p.d_o = (const bf16*)d_o;
  p.o_bs = o_bs; p.o_rs = o_rs; p.do_bs = do_bs; p.do_rs = do_rs;
  p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv;
  p.dq_bs = dq_bs; p.dq_rs = dq_rs; p.dk_bs = dk_bs; p.dk_rs = dk_rs; p.dv_bs = dv_bs; p.dv_rs = dv_rs;
  p.lse = lse; p.kmask = kmask; p.B = B; p.H = H; p.Tq = Tq; p.Sk = Sk; p.Nq = Nq; p.causal = causal;
  p.scale = scale; p.p_drop = p_drop; p.seed = seed; p.offset = offset; p.offset_ptr = rng_offset_ptr;
  int rc;
  if (p_drop > 0.f) rc = launch_attn_bwd_tc<2, true>(tq, tk, tv, tdo, p, stream);
  else if (Nq <= 128) rc = launch_attn_bwd_tc<2, false>(tq, tk, tv, tdo, p, stream);
  else rc = launch_attn_bwd_tc<4, false>(tq, tk, tv, tdo, p, stream);
  return rc ? rc : 1;
}

// =====================================================================================================================
// tcgen05 attention FORWARD (head dim 64, Sk <= 224 — single key block, so no online-softmax rescaling is needed).
// Item = (batch, head): K / V are staged once, then every 128-query tile runs  S = Q K^T  (N = Sk padded to 16) ->
// row softmax by the 256 compute threads (thread <-> query row, two warps per TMEM lane quadrant split the key
// columns; row max / row sum are combined through shared memory) -> P (bf16, dropout applied) written in the K-major
// SWIZZLE_128B layout -> O = P V (V consumed MN-major straight from its [keys][64] tile) -> O / l, LSE.
// Same dropout stream and LSE convention (natural log) as attention.cu, so either backward can follow.
struct AttnTcFwdParams {
  bf16* o;
  long long o_bs, o_rs;
  float* lse;
  const uint8_t* kmask;
  int B, H, Tq, Sk, Nk;    // Nk = Sk rounded up to 32
  int causal;
  float scale, p_drop;
  unsigned long long seed, offset;
  const unsigned long long* offset_ptr;
};

// Softmax warps: ATF_SPLIT per TMEM lane quadrant (thread <-> query row, the warps of a quadrant split the key columns).  Round 1 ran 2 per
// quadrant: one warp then owned up to 112 score columns of its rows and its dependent instruction stream (TMEM load -> FFMA -> EX2 ->
// dropout hash -> pack -> STS, ~10 cycles per issued instruction with 2 warps per scheduler: profiles/ncu_r1c_attention_fwd.txt) WAS the
// tile time.  4 per quadrant halve the stream of every warp and double the warps each scheduler can interleave.
static constexpr int ATF_SPLIT = 4;
static constexpr int ATF_THREADS = 64 + 128 * ATF_SPLIT;     // 2 control warps + 16 softmax warps

// Shared memory (dynamic, sized by Nk = keys rounded up to 32): two Q tiles, two K and two V buffers (item parity), one P tile.
struct AtcFwdSmem {
  static constexpr int Q_BYTES = 128 * 128;      // one 128-query tile
  static constexpr int NUM_BARS = 13;
  __host__ __device__ static int kv_bytes(int Nk) { return Nk * 128; }
  __host__ __device__ static int p_bytes(int Nk) { return ((Nk + 63) / 64) * 16384; }
  __host__ __device__ static int off_k(int) { return 2 * Q_BYTES; }
  __host__ __device__ static int off_v(int Nk) { return off_k(Nk) + 2 * kv_bytes(Nk); }
  __host__ __device__ static int off_p(int Nk) { return off_v(Nk) + 2 * kv_bytes(Nk); }
  __host__ __device__ static int off_red(int Nk) { return off_p(Nk) + p_bytes(Nk); }   // float [2][ATF_SPLIT][128]
  __host__ __device__ static int off_kok(int Nk) { return off_red(Nk) + 2 * ATF_SPLIT * 128 * 4; }   // uint8 [256]
  __host__ __device__ static int off_bal(int Nk) { return off_kok(Nk) + 256; }             // uint32 [4 * ATF_SPLIT warps][8]
  __host__ __device__ static int off_bar(int Nk) { return off_bal(Nk) + 4 * ATF_SPLIT * 32; }
  __host__ __device__ static int total(int Nk) { return off_bar(Nk) + NUM_BARS * 8 + 16 + 1024; }
};

// Pipeline (per CTA, tiles t = 0, 1, ... over its (item, 128-query tile) sequence):
//   TMA warp   : K/V of item n -> buffer n&1 (freed by the last P V of item n-2), Q of tile t -> buffer t&1 (freed by S(t-2)).
//   MMA warp   : S(0), S(1); then for every t: wait P(t) -> O = P V (TMEM cols [2 Nk, 2 Nk + 64)) -> S(t+2) into the S buffer
//                t&1 that the softmax of tile t has just drained.  So S(t+1) is always complete when the softmax warps get to it.
//   softmax    : pass 1 (row max) of tile t, then the EPILOGUE OF TILE t-1 (its P V ran under pass 1), then pass 2 (P -> smem).
template <int NC, bool DROPOUT>
__global__ void __launch_bounds__(ATF_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, AttnTcFwdParams p) {
  using S = AtcFwdSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int Nk = p.Nk;
  const int KVB = S::kv_bytes(Nk);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + S::off_k(Nk);
  uint8_t* sV = smem + S::off_v(Nk);
  uint8_t* sP = smem + S::off_p(Nk);
  float* sRed = reinterpret_cast<float*>(smem + S::off_red(Nk));
  uint32_t* sBal = reinterpret_cast<uint32_t*>(smem + S::off_bal(Nk));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::off_bar(Nk));
  uint64_t* kv_full = bars + 0;    // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* q_full = bars + 4;     // [2]
  uint64_t* q_empty = bars + 6;    // [2]
  uint64_t* s_full = bars + 8;     // [2]
  uint64_t* p_ready = bars + 10;
  uint64_t* o_full = bars + 11;
  uint64_t* o_free = bars + 12;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + S::NUM_BARS);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = p.B * p.H;
  const int qtiles = (p.Tq + 127) / 128;
  const int my_items = (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int ntot = my_items * qtiles;             // tiles this CTA processes
  const int SB = Nk;                              // TMEM column stride between the two S buffers (Nk <= 224)
  const int O_COL = 2 * Nk;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(p_ready, 4 * ATF_SPLIT);
    mbar_init(o_full, 1);
    mbar_init(o_free, 4 * ATF_SPLIT);
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc<512>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    if (lane == 0) {
      int t = 0, it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        const int b = item / p.H, h = item % p.H;
        const int kvb = it & 1;
        mbar_wait(&kv_empty[kvb], ((it >> 1) & 1u) ^ 1u);
        mbar_expect_tx(&kv_full[kvb], 2u * (uint32_t)KVB);
        tma_load_4d(&tm_k, &kv_full[kvb], sK + kvb * KVB, 0, h, 0, b);
        tma_load_4d(&tm_v, &kv_full[kvb], sV + kvb * KVB, 0, h, 0, b);
        for (int i = 0; i < qtiles; ++i, ++t) {
          const int qb = t & 1;
          mbar_wait(&q_empty[qb], ((t >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&q_full[qb], S::Q_BYTES);
          tma_load_4d(&tm_q, &q_full[qb], sQ + qb * S::Q_BYTES, 0, h, i * 128, b);
        }
      }
    }
  } else if (warp_idx == 1) {
    // whole warp runs the loop (uniform datapath for descriptors / barriers), one elected lane issues the tcgen05 ops
    const bool leader = elect_one();
    const uint32_t id_s = make_idesc_bf16(128, Nk, false, false);
    const uint32_t id_o = make_idesc_bf16(128, 64, false, true);
    const uint32_t aP = smem_u32(sP);
    const int nk16 = Nk / 16;
    auto issue_s = [&](int t) {
      const int it = t / qtiles;
      const int kvb = it & 1, qb = t & 1;
      mbar_wait(&kv_full[kvb], (it >> 1) & 1u);
      mbar_wait(&q_full[qb], (t >> 1) & 1u);
      tc_fence_after();
      if (leader) {
        const uint32_t aQ = smem_u32(sQ + qb * S::Q_BYTES), aK = smem_u32(sK + kvb * KVB);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + (uint32_t)(qb * SB), make_smem_desc(aQ + k * 32, 16, 1024), make_smem_desc(aK + k * 32, 16, 1024),
                    id_s, k > 0);
        umma_commit(&s_full[qb]);
        umma_commit(&q_empty[qb]);
      }
      __syncwarp();
    };
    if (ntot > 0) issue_s(0);
    if (ntot > 1) issue_s(1);
    for (int t = 0; t < ntot; ++t) {
      const int it = t / qtiles, i = t - it * qtiles;
      const int kvb = it & 1;
      mbar_wait(p_ready, t & 1u);
      mbar_wait(o_free, (t & 1u) ^ 1u);          // epilogue of tile t-1 has drained the O accumulator
      tc_fence_after();
      if (leader) {
        const uint32_t aV = smem_u32(sV + kvb * KVB);
        for (int k = 0; k < nk16; ++k)
          umma_bf16(tmem_base + (uint32_t)O_COL, make_smem_desc(aP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    make_smem_desc(aV + k * 2048, 16384, 1024), id_o, k > 0);
        umma_commit(o_full);
        if (i == qtiles - 1) umma_commit(&kv_empty[kvb]);
      }
      __syncwarp();
      if (t + 2 < ntot) issue_s(t + 2);
    }
  } else {
    const int quad = warp_idx & 3;
    const int part = (warp_idx - 2) >> 2;        // which slice of the key columns this warp handles
    const int r = quad * 32 + lane;              // query row inside the tile
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;   // > 0 (host checks): max and scaling commute
    const unsigned long long off_eff = p.offset + ((DROPOUT && p.offset_ptr) ? __ldg(p.offset_ptr) : 0ull);
    const uint32_t thr = (uint32_t)(p.p_drop * 4294967296.0f);
    const float inv_keep = DROPOUT ? 1.f / (1.f - p.p_drop) : 1.f;
    // Key columns are dealt to the ATF_SPLIT warps of a quadrant in 8-key chunks, round robin: chunk c of this thread covers keys
    // [32 c + 8 part, +8).  (A contiguous slice per warp left the warps of the high slices idle under a causal mask: 30 % of all
    // stall samples were barrier waits, profiles/ncu_r2_attention.txt.)  NC = max chunks per thread (Nk <= 32 NC): the chunk loops
    // are fully unrolled and the raw scores stay in registers between the max pass and the exp pass — S is read from TMEM once.
    const int nchunks = Nk >> 5;
    const uint32_t p_row = smem_u32(sP) + (uint32_t)(r * 128);  // shared-space address of this row inside atom 0 of the P tile
    const int r7 = r & 7;
    // 16-byte unit of chunk c inside its 64-key atom: (c & 1) * 4 + part, XOR-swizzled with the row
    const uint32_t p_even = p_row + (uint32_t)(((part) ^ r7) << 4), p_odd = p_row + (uint32_t)(((4 + part) ^ r7) << 4);
    float* redmax = sRed;                        // [ATF_SPLIT][128]
    float* redsum = sRed + ATF_SPLIT * 128;      // [ATF_SPLIT][128]
    uint32_t* sBalW = sBal + (warp_idx - 2) * 8; // this warp's copy of the key-validity words (generic masks only)
    // deferred epilogue state (tile t-1)
    float l_prev = 0.f, mref_prev = 0.f;
    long long o_off_prev = 0, lse_off_prev = 0;
    bool row_prev = false;
    auto epilogue_prev = [&](int tprev) {
      mbar_wait(o_full, tprev & 1u);
      tc_fence_after();
      uint32_t acc[16];
      tmem_ld16(lane_taddr + (uint32_t)O_COL + part * 16, acc);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);        // accumulator is in registers: the next P V may start
      if (row_prev) {
        const float l = l_prev;
        const float inv = l > 0.f ? 1.f / l : 0.f;
        uint4* dst = reinterpret_cast<uint4*>(p.o + o_off_prev + part * 16);
#pragma unroll
        for (int e = 0; e < 2; ++e)
          dst[e] = make_uint4(pack_bf16x2(__uint_as_float(acc[8 * e]) * inv, __uint_as_float(acc[8 * e + 1]) * inv),
                              pack_bf16x2(__uint_as_float(acc[8 * e + 2]) * inv, __uint_as_float(acc[8 * e + 3]) * inv),
                              pack_bf16x2(__uint_as_float(acc[8 * e + 4]) * inv, __uint_as_float(acc[8 * e + 5]) * inv),
                              pack_bf16x2(__uint_as_float(acc[8 * e + 6]) * inv, __uint_as_float(acc[8 * e + 7]) * inv));
        if (part == 0 && p.lse) p.lse[lse_off_prev] = l > 0.f ? (mref_prev + log2f(l)) * 0.6931471805599453f : -INFINITY;
      }
    };
    int t = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int b = item / p.H, h = item % p.H;
      const uint32_t dkey = attn_drop_key(p.seed, off_eff, item);
      // ---- per-item key validity, computed by every warp for itself (no CTA-wide barrier in the tile loop).  The usual masks
      // (no mask, right padding) make the valid keys a PREFIX [0, nvalid): then an 8-key chunk is either fully valid for the whole
      // warp (no per-element predicate at all), fully masked (skipped) or one of the few mixed ones (causal diagonal / last
      // partial chunk).  Arbitrary masks keep a per-element test against the ballot words.
      int nvalid = 0;
      bool prefix = true;
      {
        uint32_t bal[7];
#pragma unroll
        for (int w = 0; w < 7; ++w) {
          const int key = w * 32 + lane;
          const bool ok = (key < p.Sk) && (!p.kmask || p.kmask[(long long)b * p.Sk + key]);
          bal[w] = __ballot_sync(0xffffffffu, ok);
          nvalid += __popc(bal[w]);
        }
#pragma unroll
        for (int w = 0; w < 7; ++w) {
          const int lo = nvalid - 32 * w;
          const uint32_t expect = lo >= 32 ? 0xffffffffu : (lo <= 0 ? 0u : ((1u << lo) - 1u));
          prefix = prefix && (bal[w] == expect);
        }
        if (!prefix) {
          __syncwarp();
          if (lane < 7) {
            uint32_t mine = 0;
#pragma unroll
            for (int w = 0; w < 7; ++w) mine = (lane == w) ? bal[w] : mine;
            sBalW[lane] = mine;
          }
          __syncwarp();
        }
      }
      for (int i = 0; i < qtiles; ++i, ++t) {
        const int sb = t & 1;
        const uint32_t s_taddr = lane_taddr + (uint32_t)(sb * SB + part * 8);
        const int qq = i * 128 + r;
        // key k is visible to this row iff k < kmax_row (prefix masks); every lane of the warp sees all keys < kfull and no
        // key >= kany.  Generic masks: kfull = 0, kany = Nk, visibility from the ballot words.
        int kmax_row = 0, kfull = 0, kany = Nk;
        if (prefix) {
          const int q_lo = i * 128 + quad * 32;
          kmax_row = p.causal ? min(nvalid, qq + 1) : nvalid;
          kfull = p.causal ? min(nvalid, q_lo + 1) : nvalid;
          kany = p.causal ? min(nvalid, q_lo + 32) : nvalid;
        }
        auto visible = [&](int kk) -> bool {     // generic masks only
          return ((sBalW[kk >> 5] >> (kk & 31)) & 1u) && (!p.causal || kk <= qq);
        };
        int nact = (kany - part * 8 + 31) >> 5;  // chunks of this thread that hold at least one visible key (warp-uniform)
        nact = nact < 0 ? 0 : (nact > nchunks ? nchunks : nact);
        mbar_wait(&s_full[sb], (t >> 1) & 1u);
        tc_fence_after();
        // ---- pass 1: row max.  Chunks [0, nfull) are fully visible to every lane of the warp: straight-line code selected by
        // nfull (compile-time chunk count: no per-chunk compare / branch, constant TMEM columns), TMEM loads two deep.  The few
        // chunks [nfull, nact) on the causal diagonal / at the end of the valid keys take the predicated loop.
        // (S stays in TMEM between the passes: keeping it in registers was measured SLOWER at 96 registers per thread —
        //  62 vs 52 us on the ViT shape, tools/jobs/r2p.sh.)
        int nfull = kfull - part * 8 + 24;       // = 32 * (#chunks with k0 + 8 <= kfull)  rounded down below
        nfull = nfull < 0 ? 0 : (nfull >> 5);
        nfull = nfull > nact ? nact : nfull;
        uint32_t ta[8], tb[8];
        float mx = -INFINITY;
        dispatch_count<NC>(nfull, [&](auto nconst) {
          constexpr int N = decltype(nconst)::value;
          if (N > 0) tmem_ld8(s_taddr, ta);
#pragma unroll
          for (int c = 0; c < N; c += 2) {
            tmem_ld_wait();
            if (c + 1 < N) tmem_ld8(s_taddr + (uint32_t)((c + 1) * 32), tb);
#pragma unroll
            for (int e = 0; e < 8; ++e) mx = fmaxf(mx, __uint_as_float(ta[e]));
            if (c + 1 < N) {
              tmem_ld_wait();
              if (c + 2 < N) tmem_ld8(s_taddr + (uint32_t)((c + 2) * 32), ta);
#pragma unroll
              for (int e = 0; e < 8; ++e) mx = fmaxf(mx, __uint_as_float(tb[e]));
            }
          }
        });
#pragma unroll 1
        for (int c = nfull; c < nact; ++c) {
          tmem_ld8(s_taddr + (uint32_t)(c * 32), ta);
          tmem_ld_wait();
          const int k0 = c * 32 + part * 8;
          if (prefix) {
            const int n = kmax_row - k0;         // visible keys of this row inside the chunk
#pragma unroll
            for (int e = 0; e < 8; ++e) mx = (e < n) ? fmaxf(mx, __uint_as_float(ta[e])) : mx;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (visible(k0 + e)) mx = fmaxf(mx, __uint_as_float(ta[e]));
          }
        }
        redmax[part * 128 + r] = mx;
        named_bar_sync(2 + quad, 32 * ATF_SPLIT);          // only the warps that share these 32 rows
#pragma unroll
        for (int s2 = 0; s2 < ATF_SPLIT; ++s2) mx = fmaxf(mx, redmax[s2 * 128 + r]);
        const float mref = (mx == -INFINITY) ? 0.f : mx * sl2;
        // ---- deferred epilogue of the previous tile: its P V product ran while pass 1 was executing.  Waiting for it also
        // guarantees that the tensor core has finished reading the (single) P tile before pass 2 overwrites it.
        if (t > 0) epilogue_prev(t - 1);
        // ---- pass 2: P = exp2(s * scale * log2e - m), row sum, dropout, bf16 -> swizzled smem (same fast / predicated split)
        float sum = 0.f;
        auto finish_chunk = [&](float (&pe)[8], int k0, uint32_t dst) {
#pragma unroll
          for (int e = 0; e < 8; ++e) sum += pe[e];
          if (DROPOUT) {
#pragma unroll
            for (int e = 0; e < 8; ++e) pe[e] = attn_drop_rand(dkey, qq, k0 + e, p.Sk) >= thr ? pe[e] * inv_keep : 0.f;
          }
          sts128(dst, pack_bf16x2(pe[0], pe[1]), pack_bf16x2(pe[2], pe[3]), pack_bf16x2(pe[4], pe[5]), pack_bf16x2(pe[6], pe[7]));
        };
        dispatch_count<NC>(nfull, [&](auto nconst) {
          constexpr int N = decltype(nconst)::value;
          if (N > 0) tmem_ld8(s_taddr, ta);
#pragma unroll
          for (int c = 0; c < N; c += 2) {
            tmem_ld_wait();
            if (c + 1 < N) tmem_ld8(s_taddr + (uint32_t)((c + 1) * 32), tb);
            {
              float pe[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) pe[e] = ex2_approx(fmaf(__uint_as_float(ta[e]), sl2, -mref));
              finish_chunk(pe, c * 32 + part * 8, p_even + (uint32_t)((c >> 1) * 16384));          // c is even here
            }
            if (c + 1 < N) {
              tmem_ld_wait();
              if (c + 2 < N) tmem_ld8(s_taddr + (uint32_t)((c + 2) * 32), ta);
              float pe[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) pe[e] = ex2_approx(fmaf(__uint_as_float(tb[e]), sl2, -mref));
              finish_chunk(pe, (c + 1) * 32 + part * 8, p_odd + (uint32_t)((c >> 1) * 16384));
            }
          }
        });
#pragma unroll 1
        for (int c = nfull; c < nchunks; ++c) {
          const uint32_t dst = ((c & 1) ? p_odd : p_even) + (uint32_t)((c >> 1) * 16384);
          if (c >= nact) {                       // no visible key: P = 0 (the P V product runs over all Nk columns)
            sts128(dst, 0u, 0u, 0u, 0u);
            continue;
          }
          tmem_ld8(s_taddr + (uint32_t)(c * 32), ta);
          tmem_ld_wait();
          const int k0 = c * 32 + part * 8;
          float pe[8];
          if (prefix) {
            const int n = kmax_row - k0;
#pragma unroll
            for (int e = 0; e < 8; ++e) pe[e] = (e < n) ? ex2_approx(fmaf(__uint_as_float(ta[e]), sl2, -mref)) : 0.f;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) pe[e] = visible(k0 + e) ? ex2_approx(fmaf(__uint_as_float(ta[e]), sl2, -mref)) : 0.f;
          }
          finish_chunk(pe, k0, dst);
        }
        redsum[part * 128 + r] = sum;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        named_bar_sync(2 + quad, 32 * ATF_SPLIT);
        float l = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < ATF_SPLIT; ++s2) l += redsum[s2 * 128 + r];
        l_prev = l;
        mref_prev = mref;
        row_prev = qq < p.Tq;
        o_off_prev = (long long)b * p.o_bs + (long long)qq * p.o_rs + h * 64;
        lse_off_prev = (long long)item * p.Tq + qq;
      }
    }
    if (t > 0) epilogue_prev(t - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// returns 1 if launched, 0 if the shape is outside the envelope, <0 on error
int attention_fwd_tc_dispatch(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                              const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse,
                              const uint8_t* kmask, int B, int H, int Tq, int Sk, int DH, int causal, float scale, float p_drop,
                              unsigned long long seed, unsigned long long offset, const unsigned long long* rng_offset_ptr,
                              cudaStream_t stream) {
  if (DH != 64 || Sk > 224 || (o_rs % 8) || (o_bs % 8) || !(scale > 0.f)) return 0;   // two S buffers + O in 512 TMEM columns
  bind_context_for_driver_calls();
  const int Nk = (Sk + 31) / 32 * 32;
  CUtensorMap tq, tk, tv;
  auto mk = [&](CUtensorMap* tm, const void* ptr, long long bs, long long rs, int rows, int box_rows) {
    const uint64_t dims[4] = {64, (uint64_t)H, (uint64_t)rows, (uint64_t)B};
    const uint64_t strides[3] = {64, (uint64_t)rs, (uint64_t)bs};
    const uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
    return make_tmap_bf16_nd(tm, ptr, 4, dims, strides, box);
  };
  if (mk(&tq, q, q_bs, q_rs, Tq, 128) || mk(&tk, k, k_bs, k_rs, Sk, Nk) || mk(&tv, v, v_bs, v_rs, Sk, Nk)) return -1;
  AttnTcFwdParams p;
  p.o = (bf16*)o; p.o_bs = o_bs; p.o_rs = o_rs; p.lse = lse; p.kmask = kmask;
  p.B = B; p.H = H; p.Tq = Tq; p.Sk = Sk; p.Nk = Nk; p.causal = causal; p.scale = scale; p.p_drop = p_drop;
  p.seed = seed; p.offset = offset; p.offset_ptr = rng_offset_ptr;
  static bool attr_set = false;
  if (!attr_set) {
    const int big = AtcFwdSmem::total(224), small = AtcFwdSmem::total(128);
    cudaError_t err = cudaFuncSetAttribute(attn_fwd_tc_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(attn_fwd_tc_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(attn_fwd_tc_kernel<7, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(attn_fwd_tc_kernel<7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (err != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_fwd_tc smem=%d): %s", big, cudaGetErrorString(err));
      return -1;
    }
    attr_set = true;
  }
  const int items = B * H;
  const int grid = items < num_sms() ? items : num_sms();
  const int smem_bytes = AtcFwdSmem::total(Nk);
  const bool drop = p_drop > 0.f;
  if (Nk <= 128) {
    if (drop) attn_fwd_tc_kernel<4, true><<<grid, ATF_THREADS, smem_bytes, stream>>>(tq, tk, tv, p);
    else attn_fwd_tc_kernel<4, false><<<grid, ATF_THREADS, smem_bytes, stream>>>(tq, tk, tv, p);
  } else {
    if (drop) attn_fwd_tc_kernel<7, true><<<grid, ATF_THREADS, smem_bytes, stream>>>(tq, tk, tv, p);
    else attn_fwd_tc_kernel<7, false><<<grid, ATF_THREADS, smem_bytes, stream>>>(tq, tk, tv, p);
  }
  const int rc = check_launch("attn_fwd_tc");
  return rc ? rc : 1;
}

}  // namespace vlm
