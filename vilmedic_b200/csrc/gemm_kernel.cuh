// tcgen05 GEMM kernel + launchers, shared by the per-tile-width translation units (gemm_tcgen05_bn*.cu: one explicit
// instantiation of gemm_launch_bn<BN> each, so that the ~100 kernel variants compile in parallel) and the host entry
// point in gemm_tcgen05.cu.  See gemm_tcgen05.cu for the description of the kernel.
#pragma once
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "vlm_b200.h"

namespace vlm {


static constexpr int GEMM_BM = 128;
static constexpr int GEMM_BK = 64;  // 64 bf16 = 128 B = one swizzle row
// Epilogue warps: PARTS per TMEM lane quadrant (warp w drains quadrant w%4, 16-column chunks c = part, part+PARTS, ...).
// Register budget is per SM sub-partition (16384 registers, warps are dealt round-robin): 2 + 4*4 = 18 warps -> 5 on one
// sub-partition -> <= 96 registers; 2 + 4*3 = 14 warps -> 4 per sub-partition -> <= 128 registers.
__host__ __device__ constexpr int gemm_threads(int parts) { return 64 + 128 * parts; }
static constexpr int GEMM_MAX_EPI_WARPS = 12;

// Shared-memory plan.  Specialised (staged-epilogue) kernels carry, per epilogue warp, STG_BUFS 32x32 bf16 staging tiles for
// the TMA-store epilogue plus two BN-float bias buffers; they pay for it with one pipeline stage (K = 768..3072 mainloops are
// consumer-bound: the ring was full at 5-6 stages, profiles/ncu_gemm_r1_epilogue.txt).
// SBUFS = 2 (double-buffered staging, one stage less) is used for short mainloops (K <= 1536: epilogue-bound), SBUFS = 1 for
// long ones (K = 3072: mainloop-bound, the extra stage matters more) — measured with tools/gemm_bench.py.
template <int BN, int MODE, int SBUFS>
struct GemmSmem {
  static constexpr bool FAST = MODE != EPI_GENERIC;
  static constexpr int STG_BUFS = FAST ? SBUFS : 0;
  static constexpr int STAGES = (FAST && SBUFS == 2) ? (BN >= 192 ? 4 : (BN >= 128 ? 5 : 6)) : ((BN >= 256) ? 4 : (BN >= 192 ? 5 : 6));
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;            // epilogue staging tiles
  static constexpr int STG_BYTES = 32 * 32 * 2;
  static constexpr int BIAS_OFFSET = STG_OFFSET + GEMM_MAX_EPI_WARPS * STG_BUFS * STG_BYTES;   // float [2][BN]
  static constexpr int BAR_OFFSET = BIAS_OFFSET + (FAST ? 2 * BN * 4 : 0);
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;  // + alignment slack
  static_assert(TOTAL <= 232448, "shared memory plan exceeds 227 KB");
};

template <int BN, bool A_MN, bool B_MN, int PARTS, int GEMM_EPI_W, int MODE, int SBUFS>
__global__ void __launch_bounds__(gemm_threads(PARTS), 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_aux, int tma_store,
                         int M, int N, int K, int batch, int a_bmul, int b_bmul, int split_k,
                         long long c_batch_stride, long long aux_batch_stride, long long res_batch_stride,
                         GemmEpilogue epi) {
  using S = GemmSmem<BN, MODE, SBUFS>;
  constexpr int STAGES = S::STAGES;
  constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;  // two accumulator buffers of BN fp32 columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
  const int tiles_per_batch = m_tiles * n_tiles;
  const int total_tiles = tiles_per_batch * batch;
  // split-K: work item = (tile, split); split s covers k-blocks [s*kb_per, min(k_blocks, (s+1)*kb_per))
  const int kb_per = (k_blocks + split_k - 1) / split_k;
  const int total_work = total_tiles * split_k;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (tma_store) {
      tma_prefetch_desc(&tmap_c);
      tma_prefetch_desc(&tmap_aux);
    }
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4 * PARTS);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int tile = work % total_tiles, split = work / total_tiles;
        const int kb0 = split * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
        if (kb0 >= kb1) continue;
        const int b = tile / tiles_per_batch;
        const int t = tile - b * tiles_per_batch;
        const int m0 = (t / n_tiles) * GEMM_BM;
        const int n0 = (t % n_tiles) * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          const int k0 = kb * GEMM_BK;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j)
              tma_load_3d(&tmap_a, &full_bar[stage], sa + j * (GEMM_BK * 128), m0 + j * 64, k0, b * a_bmul);
          } else {
            tma_load_3d(&tmap_a, &full_bar[stage], sa, k0, m0, b * a_bmul);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_3d(&tmap_b, &full_bar[stage], sb + j * (GEMM_BK * 128), n0 + j * 64, k0, b * b_bmul);
          } else {
            tma_load_3d(&tmap_b, &full_bar[stage], sb, k0, n0, b * b_bmul);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop (so tile / stage / descriptor arithmetic stays on the uniform datapath) and one elected
    // lane issues.  ncu showed the single-lane version spending ~100 dependent instructions (ELECT + R2UR per operand) per
    // k-block: 930 cycles of issue latency for 349 cycles of tensor work (profiles/ncu_gemm_r1_epilogue.txt).
    constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, A_MN, B_MN);
    // K-major: 8-row groups are 1024 B apart; advancing 16 k = +32 B inside the swizzled row.
    // MN-major: 64-wide MN blocks are BK*128 B apart (LBO), 8-k groups 1024 B apart (SBO); 16 k = +2048 B.
    constexpr uint32_t A_LBO = A_MN ? GEMM_BK * 128 : 0, B_LBO = B_MN ? GEMM_BK * 128 : 0;
    constexpr uint32_t A_KSTEP = A_MN ? 2048 : 32, B_KSTEP = B_MN ? 2048 : 32;
    // descriptor = hi word (SBO 1024 B, version 1, SWIZZLE_128B) : lo word (start address >> 4 | LBO >> 4 << 16)
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | ((A_LBO >> 4) << 16);
    const uint32_t b_lo0 = (((smem_u32(smem) + S::A_BYTES) >> 4) & 0x3FFFu) | ((B_LBO >> 4) << 16);
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      const int split = work / total_tiles;
      const int kb0 = split * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      if (kb0 >= kb1) continue;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)stage * (S::STAGE_BYTES >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)stage * (S::STAGE_BYTES >> 4);
        if (leader) {
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t da = ((uint64_t)DESC_HI << 32) | (uint64_t)(a_lo + k * (A_KSTEP >> 4));
            const uint64_t db = ((uint64_t)DESC_HI << 32) | (uint64_t)(b_lo + k * (B_KSTEP >> 4));
            umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs have read it
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (leader) umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps (2 .. 2+4*PARTS) =====================
    // Warp w drains TMEM lane quadrant w%4 (rows quad*32 .. +32 of the tile) and the 32-column spans s = part, part+PARTS, ...
    // Each span is two 16-column chunks.  bf16 C without accumulation leaves through shared memory: every lane writes its
    // row into a 64B-swizzled 32x32 staging tile and one lane issues a TMA tensor store (full-line writes, clipped at the
    // M / N edges by the tensor map), so the epilogue warps never wait on global stores.
    static_assert(GEMM_EPI_W == 16, "a span is two 16-column chunks");
    const int quad = warp_idx & 3;                 // TMEM lane quadrant this warp may access
    const int part = (warp_idx - 2) >> 2;          // which spans of the tile's columns this warp drains
    constexpr int SPANS = BN / 32;
    const unsigned long long rng_add = (epi.p_drop > 0.f && epi.offset_ptr) ? __ldg(epi.offset_ptr) : 0ull;
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (MODE != EPI_GENERIC) {
      // ---- staged path only (host guarantees: bf16 C by TMA store, no accumulation, split_k == 1).  Ragged M / N edges are clipped
      // by the C / aux tensor maps; bias and row-input reads are guarded.  The kernel carries no direct-store fallback, which keeps
      // the epilogue warps inside their 128-register budget without spilling the in-flight row inputs.
      constexpr int STG_BUFS = S::STG_BUFS;
      constexpr int NSP = (SPANS + PARTS - 1) / PARTS;
      constexpr bool HAS_PRE = MODE == EPI_GELUGRAD || MODE == EPI_RESID;
      const uint32_t stg_base = smem_u32(smem + S::STG_OFFSET) + (uint32_t)((warp_idx - 2) * STG_BUFS * S::STG_BYTES);
      uint32_t stg_cnt = 0;
      const uint32_t sbias_base = smem_u32(smem + S::BIAS_OFFSET);
      const unsigned long long drop_off = epi.offset + rng_add;
      // Row inputs (GELU' argument / residual) are requested one span ahead into 16 registers (gemm_epilogue.cuh: PreReq / pre_issue);
      // the request for the first span of the NEXT tile is issued under the last span of this one.
      uint4 pre[4];
      bool have_pre = false;
      const bf16* const pre_base = MODE == EPI_GELUGRAD ? epi.aux_in : reinterpret_cast<const bf16*>(epi.residual);
      const long long pre_ld = MODE == EPI_GELUGRAD ? epi.ld_aux : epi.ldr;
      const long long pre_bs = MODE == EPI_GELUGRAD ? aux_batch_stride : res_batch_stride;
      const bool use_pre = HAS_PRE && pre_base != nullptr;
      const int nv = max(0, min(NSP, SPANS - part * NSP));      // spans of a tile this warp drains
      auto pre_req = [&](int i, int tb, int tm0, int tn0) {
        PreReq rq;
        const int r0 = tm0 + quad * 32, c0 = tn0 + (part * NSP + i) * 32;
        rq.base = (use_pre && i < nv && c0 < N) ? pre_base + (size_t)tb * pre_bs + (long long)r0 * pre_ld + c0 : nullptr;
        rq.ld = pre_ld;
        rq.rows = M - r0;
        rq.cols = N - c0;
        return rq;
      };
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int b = work / tiles_per_batch;
        const int t = work - b * tiles_per_batch;
        const int m0 = (t / n_tiles) * GEMM_BM;
        const int n0 = (t % n_tiles) * BN;
        // the tile this CTA processes next (its first span's inputs are requested under this tile's last span)
        const int nwork = work + gridDim.x;
        int nb = 0, nm0 = 0, nn0 = 0;
        const bool has_next = use_pre && nwork < total_work;
        if (has_next) {
          nb = nwork / tiles_per_batch;
          const int nt = nwork - nb * tiles_per_batch;
          nm0 = (nt / n_tiles) * GEMM_BM;
          nn0 = (nt % n_tiles) * BN;
        }
        if constexpr (HAS_PRE) {
          if (use_pre && !have_pre) pre_issue(pre, pre_req(0, b, m0, n0), lane);   // first tile of this CTA
        }
        // bias of this tile -> shared memory (buffer by accumulator parity), while the mainloop of the tile is still running
        if (epi.bias) {
          for (int i = threadIdx.x - 64; i < BN; i += 128 * PARTS)
            reinterpret_cast<float*>(smem + S::BIAS_OFFSET)[acc * BN + i] = (n0 + i < N) ? __ldg(epi.bias + n0 + i) : 0.f;
          named_bar_sync(1, 128 * PARTS);
        }
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        const float alpha = epi.alpha_ptr ? epi.alpha * __ldg(epi.alpha_ptr) : epi.alpha;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
        for (int i = 0; i < NSP; ++i) {
          const int sp = part * NSP + i;       // adjacent spans: a warp's row inputs form whole 128-byte lines
          if (i < nv && n0 + sp * 32 < N) {    // warp-uniform
            PreReq nx;
            nx.base = nullptr;
            if constexpr (HAS_PRE) {
              if (i + 1 < nv && n0 + (sp + 1) * 32 < N) nx = pre_req(i + 1, b, m0, n0);
              else if (has_next) nx = pre_req(0, nb, nm0, nn0);
            }
            epilogue_span_fast<MODE, STG_BUFS>(taddr + sp * 32, m0 + quad * 32, n0 + sp * 32, b, lane, epi, alpha, drop_off, use_pre, pre, nx,
                                               stg_base, stg_cnt, epi.bias ? sbias_base + (uint32_t)((acc * BN + sp * 32) * 4) : 0u, &tmap_c,
                                               &tmap_aux);
          }
        }
        have_pre = has_next;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    } else {
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int tile = work % total_tiles, split = work / total_tiles;
        const int kb0 = split * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
        if (kb0 >= kb1) continue;
        const int b = tile / tiles_per_batch;
        const int t = tile - b * tiles_per_batch;
        const int m0 = (t / n_tiles) * GEMM_BM;
        const int n0 = (t % n_tiles) * BN;
        GemmEpilogue e = epi;
        if (split > 0) {  // bias / residual are added once, by the first K split
          e.bias = nullptr;
          e.residual = nullptr;
        }
        e.offset += rng_add;
        if (b > 0) {
          const size_t esz = e.c_fp32 ? 4 : 2;
          e.c = reinterpret_cast<uint8_t*>(e.c) + (size_t)b * c_batch_stride * esz;
          if (e.residual) e.residual = reinterpret_cast<const uint8_t*>(e.residual) + (size_t)b * res_batch_stride * esz;
          if (e.aux_in) e.aux_in += (size_t)b * aux_batch_stride;
          if (e.aux_out) e.aux_out += (size_t)b * aux_batch_stride;
        }
        const int row = m0 + quad * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int sp = part; sp < SPANS; sp += PARTS) {
          const int col_s = n0 + sp * 32;
          if (col_s >= N) break;
          epilogue_chunk<16>(taddr + sp * 32, row, col_s, M, N, e);
          epilogue_chunk<16>(taddr + sp * 32 + 16, row, col_s + 16, M, N, e);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
    if (tma_store && lane == 0) bulk_wait_all();                     // global writes complete before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


static constexpr int GEMM_PARTS = 3;   // epilogue warps per TMEM lane quadrant (14 warps per CTA, <= 128 registers each)

template <int BN, bool A_MN, bool B_MN, int MODE, int SBUFS>
static int launch_gemm_m(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tx, int tma_store,
                       int M, int N, int K,
                       int batch, int a_bmul, int b_bmul, int split_k, long long c_bs, long long aux_bs, long long res_bs,
                       const GemmEpilogue& epi, int max_ctas, cudaStream_t stream) {
  using S = GemmSmem<BN, MODE, SBUFS>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, A_MN, B_MN, GEMM_PARTS, 16, MODE, SBUFS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (err != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm smem=%d): %s", S::TOTAL, cudaGetErrorString(err));
      return -1;
    }
    attr_set = true;
  }
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM, n_tiles = (N + BN - 1) / BN;
  const long long tiles = (long long)m_tiles * n_tiles * batch * split_k;
  int grid = (int)(tiles < (long long)num_sms() ? tiles : (long long)num_sms());
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  kern<<<grid, gemm_threads(GEMM_PARTS), S::TOTAL, stream>>>(ta, tb, tc, tx, tma_store, M, N, K, batch, a_bmul, b_bmul, split_k,
                                                             c_bs, aux_bs, res_bs, epi);
  return check_launch("gemm_bf16_tcgen05");
}

// `tma_store` carries the epilogue mode (EPI_GENERIC = direct stores).  Specialised kernels exist for the operand layouts
// the modes occur with: forward (both operands K-major): bias / GELU+stash / dropout+residual; dgrad (B MN-major): plain / GELU' /
// residual.
template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tx, int mode,
                       int M, int N, int K, int batch, int a_bmul, int b_bmul, int split_k, long long c_bs, long long aux_bs,
                       long long res_bs, const GemmEpilogue& epi, int max_ctas, cudaStream_t stream) {
  // double-buffered staging only where it pays (short K) and where the smem plan has room for it (BN 128 / 192)
  const bool dbl = (BN == 128 || BN == 192) && K <= 1536;
#define VLM_MODE_CASE(M_)                                                                                                            \
  do {                                                                                                                               \
    if constexpr ((BN == 128 || BN == 192) && M_ != EPI_GENERIC) {                                                                   \
      if (dbl)                                                                                                                       \
        return launch_gemm_m<BN, A_MN, B_MN, M_, 2>(ta, tb, tc, tx, 1, M, N, K, batch, a_bmul, b_bmul, split_k, c_bs, aux_bs, res_bs, \
                                                    epi, max_ctas, stream);                                                          \
    }                                                                                                                                \
    return launch_gemm_m<BN, A_MN, B_MN, M_, 1>(ta, tb, tc, tx, M_ != EPI_GENERIC, M, N, K, batch, a_bmul, b_bmul, split_k, c_bs,      \
                                                aux_bs, res_bs, epi, max_ctas, stream);                                              \
  } while (0)
  if constexpr (!A_MN && !B_MN) {
    if (mode == EPI_BIAS) VLM_MODE_CASE(EPI_BIAS);
    if (mode == EPI_GELU) VLM_MODE_CASE(EPI_GELU);
    if (mode == EPI_RESID) VLM_MODE_CASE(EPI_RESID);
  }
  if constexpr (!A_MN && B_MN) {
    if (mode == EPI_BIAS) VLM_MODE_CASE(EPI_BIAS);
    if (mode == EPI_GELUGRAD) VLM_MODE_CASE(EPI_GELUGRAD);
    if (mode == EPI_RESID) VLM_MODE_CASE(EPI_RESID);     // dgrad + residual-gradient add (post-LN blocks)
  }
  VLM_MODE_CASE(EPI_GENERIC);
#undef VLM_MODE_CASE
}

// Runtime operand-layout dispatch for one tile width; explicitly instantiated in gemm_tcgen05_bn<BN>.cu.
template <int BN>
int gemm_launch_bn(int a_mn_major, int b_mn_major, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                   const CUtensorMap& tx, int mode, int M, int N, int K, int batch, int a_bmul, int b_bmul, int split_k, long long c_bs,
                   long long aux_bs, long long res_bs, const GemmEpilogue& epi, int max_ctas, cudaStream_t stream) {
  if (a_mn_major) {
    if (b_mn_major)
      return launch_gemm<BN, true, true>(ta, tb, tc, tx, mode, M, N, K, batch, a_bmul, b_bmul, split_k, c_bs, aux_bs, res_bs, epi, max_ctas, stream);
    return launch_gemm<BN, true, false>(ta, tb, tc, tx, mode, M, N, K, batch, a_bmul, b_bmul, split_k, c_bs, aux_bs, res_bs, epi, max_ctas, stream);
  }
  if (b_mn_major)
    return launch_gemm<BN, false, true>(ta, tb, tc, tx, mode, M, N, K, batch, a_bmul, b_bmul, split_k, c_bs, aux_bs, res_bs, epi, max_ctas, stream);
  return launch_gemm<BN, false, false>(ta, tb, tc, tx, mode, M, N, K, batch, a_bmul, b_bmul, split_k, c_bs, aux_bs, res_bs, epi, max_ctas, stream);
}

}  // namespace vlm
