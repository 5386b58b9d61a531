// 2-CTA (cta_group::2) variant of the bf16 GEMM: a CTA pair on one TPC computes a 256 x BN tile with ONE
// tcgen05.mma.cta_group::2 instruction stream issued by the leader CTA.  Each CTA stages its own 128 A rows and HALF of
// the B tile, so the L2->smem traffic per FLOP drops 1.5x versus the 1-CTA kernel (which is L2-bandwidth bound at
// ~87 FLOP/B) and the smem ring gets 6 stages of 32 KB.  Same operand conventions / epilogue as gemm_tcgen05.cu.
//
// Protocol (per pair; CTA rank r in {0 = leader, 1}):
//   full[s]   lives in the leader: 1 arrival (leader producer, expect_tx = bytes of BOTH CTAs); both producers' TMA
//             (cp.async.bulk.tensor ... .cta_group::2) complete_tx on the leader's barrier (peer bit cleared).
//   empty[s]  one per CTA: the leader's tcgen05.commit.cta_group::2 multicasts the arrival to both.
//   tmem_full[a]  one per CTA (multicast commit);  tmem_empty[a] in the leader: 2 x EPI_WARPS arrivals (remote arrive
//             from the peer's epilogue warps).
#include <cuda.h>
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "vlm_b200.h"

namespace vlm {

static constexpr int G2_BM = 128;   // rows per CTA (pair: 256)
static constexpr int G2_BK = 64;
static constexpr int G2_EPI_WARPS = 8;
static constexpr int G2_THREADS = 64 + 32 * G2_EPI_WARPS;
static constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address

// MODE != EPI_GENERIC: the staged TMA-store epilogue of the 1-CTA kernel (gemm_epilogue.cuh: epilogue_span_fast) — two 32x32 bf16
// staging tiles per epilogue warp and two BN-float bias buffers live behind the operand ring.
template <int BN, int MODE = EPI_GENERIC>
struct G2Smem {
  static constexpr bool FAST = MODE != EPI_GENERIC;
  static constexpr int A_BYTES = G2_BM * G2_BK * 2;            // 16 KB
  static constexpr int B_BYTES = (BN / 2) * G2_BK * 2;         // this CTA's half of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EXTRA = FAST ? (G2_EPI_WARPS * 2 * 2048 + 2 * BN * 4) : 0;
  static constexpr int STAGES = (200 * 1024 - EXTRA) / STAGE_BYTES > 8 ? 8 : (200 * 1024 - EXTRA) / STAGE_BYTES;
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFFSET = STG_OFFSET + (FAST ? G2_EPI_WARPS * 2 * 2048 : 0);
  static constexpr int BAR_OFFSET = BIAS_OFFSET + (FAST ? 2 * BN * 4 : 0);
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"((uint32_t)NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)NCOLS) : "memory");
}
// TMA load whose completion bytes are signalled on the LEADER CTA's mbarrier.
__device__ __forceinline__ void tma_load_3d_2cta(const void* desc, uint64_t* bar_local_addr, void* smem_dst, int c0, int c1, int c2) {
  const uint32_t bar = smem_u32(bar_local_addr) & PEER_BIT_MASK;
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit -> arrive on the barrier at the same smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  const uint16_t mask = 0x3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
// arrive on the leader CTA's copy of a barrier (works from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar_local_addr) {
  const uint32_t bar = smem_u32(bar_local_addr) & PEER_BIT_MASK;
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int BN, bool A_MN, bool B_MN, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                          const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_aux, int M, int N,
                          int K, GemmEpilogue epi) {
  using S = G2Smem<BN, MODE>;
  constexpr int STAGES = S::STAGES;
  constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  const int m_tiles = (M + 2 * G2_BM - 1) / (2 * G2_BM);
  const int n_tiles = (N + BN - 1) / BN;
  const int k_blocks = (K + G2_BK - 1) / G2_BK;
  const int total_tiles = m_tiles * n_tiles;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 2 * G2_EPI_WARPS);
    }
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc_2cta<TMEM_COLS>(tmem_ptr_smem);
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation visible pair-wide before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        const int m0 = (tile / n_tiles) * (2 * G2_BM) + (int)rank * G2_BM;
        const int n0 = (tile % n_tiles) * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);
          const int k0 = kb * G2_BK;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < G2_BM / 64; ++j)
              tma_load_3d_2cta(&tmap_a, &full_bar[stage], sa + j * (G2_BK * 128), m0 + j * 64, k0, 0);
          } else {
            tma_load_3d_2cta(&tmap_a, &full_bar[stage], sa, k0, m0, 0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j)
              tma_load_3d_2cta(&tmap_b, &full_bar[stage], sb + j * (G2_BK * 128), n0 + j * 64, k0, 0);
          } else {
            tma_load_3d_2cta(&tmap_b, &full_bar[stage], sb, k0, n0, 0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp runs the loop so that stage / descriptor arithmetic stays on the uniform datapath and one elected lane issues
    // (the single-lane loop of round 1 cost ~900 cycles of dependent issue latency per k-block, profiles/ncu_gemm_r1_epilogue.txt).
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * G2_BM, BN, A_MN, B_MN);
      constexpr uint32_t A_LBO = A_MN ? G2_BK * 128 : 16, B_LBO = B_MN ? G2_BK * 128 : 16;
      constexpr uint32_t A_KSTEP = A_MN ? 2048 : 32, B_KSTEP = B_MN ? 2048 : 32;
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      const uint32_t a_lo0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | ((A_LBO >> 4) << 16);
      const uint32_t b_lo0 = (((smem_u32(smem) + S::A_BYTES) >> 4) & 0x3FFFu) | ((B_LBO >> 4) << 16);
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (S::STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)stage * (S::STAGE_BYTES >> 4);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k) {
              const uint64_t da = ((uint64_t)DESC_HI << 32) | (uint64_t)(a_lo + k * (A_KSTEP >> 4));
              const uint64_t db = ((uint64_t)DESC_HI << 32) | (uint64_t)(b_lo + k * (B_KSTEP >> 4));
              umma_bf16_2cta(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2cta(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (issuer) umma_commit_2cta(&tmem_full_bar[acc]);
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..9), both CTAs: own 128 rows =====================
    const int quad = warp_idx & 3;
    const int half = (warp_idx - 2) >> 2;
    constexpr int CHUNKS = BN / 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (MODE != EPI_GENERIC) {
      // ---- staged path only: see the 1-CTA kernel (gemm_kernel.cuh) for the scheme; ragged edges are clipped by the tensor maps
      constexpr int PARTS = G2_EPI_WARPS / 4;                 // epilogue warps per TMEM lane quadrant
      constexpr int NSP = (CHUNKS + PARTS - 1) / PARTS;       // 32-column spans per warp
      constexpr bool HAS_PRE = MODE == EPI_GELUGRAD || MODE == EPI_RESID;
      const int part = half;
      const uint32_t stg_base = smem_u32(smem + S::STG_OFFSET) + (uint32_t)((warp_idx - 2) * 2 * 2048);
      uint32_t stg_cnt = 0;
      const uint32_t sbias_base = smem_u32(smem + S::BIAS_OFFSET);
      const unsigned long long drop_off = epi.offset + ((epi.p_drop > 0.f && epi.offset_ptr) ? __ldg(epi.offset_ptr) : 0ull);
      uint4 pre[4];                                           // row inputs of the NEXT span (gemm_epilogue.cuh: PreReq / pre_issue)
      bool have_pre = false;
      const bf16* const pre_base = MODE == EPI_GELUGRAD ? epi.aux_in : reinterpret_cast<const bf16*>(epi.residual);
      const long long pre_ld = MODE == EPI_GELUGRAD ? epi.ld_aux : epi.ldr;
      const bool use_pre = HAS_PRE && pre_base != nullptr;
      const int nv = max(0, min(NSP, CHUNKS - part * NSP));
      auto pre_req = [&](int i, int tm0, int tn0) {
        PreReq rq;
        const int r0 = tm0 + quad * 32, c0 = tn0 + (part * NSP + i) * 32;
        rq.base = (use_pre && i < nv && c0 < N) ? pre_base + (long long)r0 * pre_ld + c0 : nullptr;
        rq.ld = pre_ld;
        rq.rows = M - r0;
        rq.cols = N - c0;
        return rq;
      };
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        const int m0 = (tile / n_tiles) * (2 * G2_BM) + (int)rank * G2_BM;
        const int n0 = (tile % n_tiles) * BN;
        const int ntile = tile + num_pairs;                            // the tile this pair processes next
        const bool has_next = use_pre && ntile < total_tiles;
        int nm0 = 0, nn0 = 0;
        if (has_next) {
          nm0 = (ntile / n_tiles) * (2 * G2_BM) + (int)rank * G2_BM;
          nn0 = (ntile % n_tiles) * BN;
        }
        if constexpr (HAS_PRE) {
          if (use_pre && !have_pre) pre_issue(pre, pre_req(0, m0, n0), lane);
        }
        if (epi.bias) {                                                 // bias of this tile -> smem under the mainloop
          for (int i = threadIdx.x - 64; i < BN; i += 32 * G2_EPI_WARPS)
            reinterpret_cast<float*>(smem + S::BIAS_OFFSET)[acc * BN + i] = (n0 + i < N) ? __ldg(epi.bias + n0 + i) : 0.f;
          named_bar_sync(1, 32 * G2_EPI_WARPS);
        }
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        const float alpha = epi.alpha_ptr ? epi.alpha * __ldg(epi.alpha_ptr) : epi.alpha;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
        for (int i = 0; i < NSP; ++i) {
          const int sp = part * NSP + i;
          if (i < nv && n0 + sp * 32 < N) {
            PreReq nx;
            nx.base = nullptr;
            if constexpr (HAS_PRE) {
              if (i + 1 < nv && n0 + (sp + 1) * 32 < N) nx = pre_req(i + 1, m0, n0);
              else if (has_next) nx = pre_req(0, nm0, nn0);
            }
            epilogue_span_fast<MODE, 2>(taddr + sp * 32, m0 + quad * 32, n0 + sp * 32, 0, lane, epi, alpha, drop_off, use_pre, pre, nx,
                                        stg_base, stg_cnt, epi.bias ? sbias_base + (uint32_t)((acc * BN + sp * 32) * 4) : 0u, &tmap_c,
                                        &tmap_aux);
          }
        }
        have_pre = has_next;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    } else {
      const int c_begin = half * (CHUNKS / 2), c_end = (half + 1) * (CHUNKS / 2);
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        const int m0 = (tile / n_tiles) * (2 * G2_BM) + (int)rank * G2_BM;
        const int n0 = (tile % n_tiles) * BN;
        GemmEpilogue e = epi;
        if (e.p_drop > 0.f && e.offset_ptr) e.offset += __ldg(e.offset_ptr);
        const int row = m0 + quad * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) epilogue_chunk<32>(taddr + c * 32, row, n0 + c * 32, M, N, e);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
    if (MODE != EPI_GENERIC && lane == 0) bulk_wait_all();              // TMA stores complete before the CTA retires
  }

  tc_fence_before();
  cluster_sync_all();   // nobody exits (or frees TMEM) while the peer may still multicast / read
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
  }
}

template <int BN, bool A_MN, bool B_MN, int MODE>
static int launch_gemm2_m(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tx, int M, int N, int K,
                          const GemmEpilogue& epi, cudaStream_t stream) {
  using S = G2Smem<BN, MODE>;
  auto kern = gemm2_bf16_tcgen05_kernel<BN, A_MN, B_MN, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (err != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm2 smem=%d): %s", S::TOTAL, cudaGetErrorString(err));
      return -1;
    }
    attr_set = true;
  }
  const int m_tiles = (M + 2 * G2_BM - 1) / (2 * G2_BM), n_tiles = (N + BN - 1) / BN;
  const long long tiles = (long long)m_tiles * n_tiles;
  const int max_pairs = num_sms() / 2;
  const int pairs = (int)(tiles < max_pairs ? tiles : max_pairs);
  kern<<<2 * pairs, G2_THREADS, S::TOTAL, stream>>>(ta, tb, tc, tx, M, N, K, epi);
  return check_launch("gemm2_bf16_tcgen05");
}

// mode = EPI_* of the staged epilogue (for the layouts the 1-CTA kernel specialises: forward bias / GELU / residual, dgrad bias /
// GELU' / residual), else the generic direct-store epilogue.
template <int BN, bool A_MN, bool B_MN>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tx, int mode, int M, int N,
                        int K, const GemmEpilogue& epi, cudaStream_t stream) {
  if constexpr (!A_MN && !B_MN) {
    if (mode == EPI_BIAS) return launch_gemm2_m<BN, A_MN, B_MN, EPI_BIAS>(ta, tb, tc, tx, M, N, K, epi, stream);
    if (mode == EPI_GELU) return launch_gemm2_m<BN, A_MN, B_MN, EPI_GELU>(ta, tb, tc, tx, M, N, K, epi, stream);
    if (mode == EPI_RESID) return launch_gemm2_m<BN, A_MN, B_MN, EPI_RESID>(ta, tb, tc, tx, M, N, K, epi, stream);
  }
  if constexpr (!A_MN && B_MN) {
    if (mode == EPI_BIAS) return launch_gemm2_m<BN, A_MN, B_MN, EPI_BIAS>(ta, tb, tc, tx, M, N, K, epi, stream);
    if (mode == EPI_GELUGRAD) return launch_gemm2_m<BN, A_MN, B_MN, EPI_GELUGRAD>(ta, tb, tc, tx, M, N, K, epi, stream);
    if (mode == EPI_RESID) return launch_gemm2_m<BN, A_MN, B_MN, EPI_RESID>(ta, tb, tc, tx, M, N, K, epi, stream);
  }
  (void)mode;
  return launch_gemm2_m<BN, A_MN, B_MN, EPI_GENERIC>(ta, tb, tc, tx, M, N, K, epi, stream);
}

// Entry used by vlm_gemm_bf16 for large non-batched problems.  bn in {128, 256}.  tc / tx / mode: C and aux tensor maps + epilogue
// mode prepared by the caller for the staged epilogue (ignored by the generic path).
int gemm2_dispatch(const void* a, long long lda, int a_mn, const void* b, long long ldb, int b_mn, int M, int N, int K,
                   int bn, const GemmEpilogue& e, const CUtensorMap* tc_in, const CUtensorMap* tx_in, int mode, cudaStream_t s) {
  CUtensorMap ta, tb;
  if (a_mn) {
    if (make_tmap_bf16(&ta, a, (uint64_t)M, (uint64_t)K, 1, lda, 0, G2_BK)) return -1;
  } else {
    if (make_tmap_bf16(&ta, a, (uint64_t)K, (uint64_t)M, 1, lda, 0, G2_BM)) return -1;
  }
  if (b_mn) {
    if (make_tmap_bf16(&tb, b, (uint64_t)N, (uint64_t)K, 1, ldb, 0, G2_BK)) return -1;
  } else {
    if (make_tmap_bf16(&tb, b, (uint64_t)K, (uint64_t)N, 1, ldb, 0, bn / 2)) return -1;
  }
  const CUtensorMap& tc = tc_in ? *tc_in : ta;
  const CUtensorMap& tx = tx_in ? *tx_in : ta;
  if (!tc_in) mode = EPI_GENERIC;
#define G2_DISPATCH(BN_)                                                                           \
  if (a_mn) {                                                                                      \
    if (b_mn) return launch_gemm2<BN_, true, true>(ta, tb, tc, tx, mode, M, N, K, e, s);           \
    return launch_gemm2<BN_, true, false>(ta, tb, tc, tx, mode, M, N, K, e, s);                    \
  } else {                                                                                         \
    if (b_mn) return launch_gemm2<BN_, false, true>(ta, tb, tc, tx, mode, M, N, K, e, s);          \
    return launch_gemm2<BN_, false, false>(ta, tb, tc, tx, mode, M, N, K, e, s);                   \
  }
  if (bn == 128) { G2_DISPATCH(128) }
  G2_DISPATCH(256)
#undef G2_DISPATCH
}

}  // namespace vlm
