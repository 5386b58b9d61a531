// Fused scaled-dot-product attention, forward + backward, for the three attention flavours on the hot path:
//   * ViT bidirectional self-attention        (HF modeling_vit.py:232-247, S=197, no mask, no dropout)
//   * decoder causal self-attention + padding (HF modeling_bert_generation.py:114-153, create_causal_mask :568-574)
//   * decoder cross-attention over image feats (modeling_bert_generation.py:181-232, create_bidirectional_mask :582-588)
// Q/K/V are read in place from the packed projection outputs (row stride / head offset given by the caller), the
// [Tq,Sk] score matrix never touches HBM (online softmax over 64-key blocks), and the backward recomputes P from the
// saved log-sum-exp.  Dropout on the probabilities uses the counter-based Philox stream (seed, offset, element id),
// regenerated bit-identically in the backward.
//
// Round-1 implementation note: the inner products use the register-level mma.sync m16n8k16 bf16 path (HMMA).  The
// tiles here are 64x64x{48,64,96} per warp-group with Sk <= a few hundred, i.e. latency- not throughput-bound; the
// tcgen05 rewrite (S/P in TMEM) is the next step for this file.  GEMM-shaped projections around it are tcgen05.
#include <cstdlib>
#include "common.cuh"
#include "vlm_b200.h"

namespace vlm {

struct AttnParams {
  const bf16* q; const bf16* k; const bf16* v;
  bf16* o;
  float* lse;                 // [B,H,Tq]
  const uint8_t* kmask;       // [B,Sk] (1 = attend) or null
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;  // batch / row strides in elements; head h at +h*DH
  int B, H, Tq, Sk;
  int causal;
  float scale;
  float p_drop;
  unsigned long long seed, offset;
  const unsigned long long* offset_ptr;  // optional device-side addend to `offset`
  // backward only
  const bf16* d_o; long long do_bs, do_rs;
  bf16* dq; bf16* dk; bf16* dv;
  long long dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs;
  float* delta;               // [B,H,Tq]  rowsum(dO * O)
};

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm_x4(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const int sz = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// A [ROWS x DH] bf16 tile in shared memory; row pitch DH+8 elements (16 B pad -> conflict-free ldmatrix).
template <int ROWS, int DH>
struct Tile {
  static constexpr int PITCH = DH + 8;
  bf16 d[ROWS * PITCH];
  __device__ __forceinline__ bf16* at(int r, int c) { return d + r * PITCH + c; }
  // cooperative async load of rows [row0, row0+ROWS) (zero fill beyond nrows) by 128 threads
  __device__ __forceinline__ void load_async(const bf16* base, long long row_stride, int row0, int nrows) {
    constexpr int CH = DH / 8;
    for (int i = threadIdx.x; i < ROWS * CH; i += 128) {
      const int r = i / CH, c = i % CH;
      const bool ok = (row0 + r) < nrows;
      const bf16* src = base + (long long)(ok ? row0 + r : 0) * row_stride + c * 8;
      cp_async16(at(r, c * 8), src, ok);
    }
  }
};

// ============================================================================================ forward
template <int DH>
__global__ void __launch_bounds__(128) attn_fwd_kernel(AttnParams p) {
  constexpr int KS = DH / 16;   // k-steps over the head dim
  constexpr int NT = DH / 8;    // n-tiles of the output
  constexpr int NBUF = (DH <= 64) ? 2 : 1;   // K/V blocks are double-buffered (cp.async prefetch) when they fit 48 KB
  __shared__ __align__(16) Tile<64, DH> sQ;
  __shared__ __align__(16) Tile<64, DH> sKb[NBUF];
  __shared__ __align__(16) Tile<64, DH> sVb[NBUF];
  __shared__ uint8_t sMaskb[NBUF][64];
  __shared__ int sFull[NBUF][2];

  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int q0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  const bf16* qb = p.q + (long long)b * p.q_bs + h * DH;
  const bf16* kb = p.k + (long long)b * p.k_bs + h * DH;
  const bf16* vb = p.v + (long long)b * p.v_bs + h * DH;

  sQ.load_async(qb, p.q_rs, q0, p.Tq);
  cp_async_wait_all();
  __syncthreads();
  uint32_t qf[KS][4];
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) ldsm_x4(qf[kk], sQ.at(warp * 16 + (lane & 15), kk * 16 + (lane >> 4) * 8));

  float o_acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sl2 = p.scale * 1.4426950408889634f;
  const int qrow[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};
  const unsigned long long off_eff = p.offset + ((p.p_drop > 0.f && p.offset_ptr) ? __ldg(p.offset_ptr) : 0ull);
  const uint32_t dkey = attn_drop_key(p.seed, off_eff, bh);
  const uint32_t thr = (uint32_t)(p.p_drop * 4294967296.0f);
  const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;

  int k_end = p.Sk;
  if (p.causal) k_end = min(p.Sk, q0 + 64);
  const int nblk = (k_end + 63) / 64;
  auto issue = [&](int blk, int buf) {
    const int kk0 = blk * 64;
    sKb[buf].load_async(kb, p.k_rs, kk0, p.Sk);
    sVb[buf].load_async(vb, p.v_rs, kk0, p.Sk);
    if (threadIdx.x < 64) {
      const int kk = kk0 + threadIdx.x;
      const bool ok = (kk < p.Sk) && (!p.kmask || p.kmask[(long long)b * p.Sk + kk]);
      sMaskb[buf][threadIdx.x] = ok;
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) sFull[buf][warp] = (bal == 0xffffffffu);
    }
    cp_async_commit();
  };
  if (NBUF == 2 && nblk > 0) issue(0, 0);
  for (int it = 0; it < nblk; ++it) {
    const int k0 = it * 64;
    const int cur = (NBUF == 2) ? (it & 1) : 0;
    if (NBUF == 2) {
      if (it + 1 < nblk) { issue(it + 1, cur ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    } else {
      issue(it, 0);
      cp_async_wait<0>();
    }
    __syncthreads();
    Tile<64, DH>& sK = sKb[cur];
    Tile<64, DH>& sV = sVb[cur];
    const uint8_t* sMask = sMaskb[cur];
    // whole block attendable (no padding, not on / above the causal diagonal) -> predicate-free fast path
    const bool nomask = sFull[cur][0] && sFull[cur][1] && (!p.causal || (k0 + 63 <= q0));

    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key tiles
        uint32_t kf[4];
        // matrices: (keys np*16+0..7, dh kk*16+0..7), (same keys, dh +8), (keys +8, dh 0..7), (keys +8, dh +8)
        ldsm_x4(kf, sK.at(np * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 16 + ((lane >> 3) & 1) * 8));
        mma_bf16_16816(s[2 * np], qf[kk], kf);
        mma_bf16_16816(s[2 * np + 1], qf[kk], kf + 2);
      }
    }
    // mask + running max
    float m_new[2] = {m_run[0], m_run[1]};
    if (nomask) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s[nt][e] *= sl2;
          m_new[e >> 1] = fmaxf(m_new[e >> 1], s[nt][e]);
        }
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kc = nt * 8 + 2 * t + (e & 1);
          const int r = e >> 1;
          bool ok = sMask[kc];
          if (p.causal) ok = ok && (k0 + kc <= qrow[r]);
          s[nt][e] = ok ? s[nt][e] * sl2 : -INFINITY;
          m_new[r] = fmaxf(m_new[r], s[nt][e]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 1));
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float mref = (m_new[r] == -INFINITY) ? 0.f : m_new[r];
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f(m_run[r] - mref);
      m_run[r] = m_new[r];
      m_new[r] = mref;
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float pe[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        pe[e] = exp2f(s[nt][e] - m_new[e >> 1]);
        rs[e >> 1] += pe[e];
      }
      if (p.p_drop > 0.f) {
        // element (q, k): k = k0 + nt*8 + 2t + {0,1}; Philox word = k & 3 of call (k >> 2)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int kcol = k0 + nt * 8 + 2 * t;
          const uint32_t r0 = attn_drop_rand(dkey, qrow[r], kcol, p.Sk), r1 = attn_drop_rand(dkey, qrow[r], kcol + 1, p.Sk);
          pe[2 * r] = r0 >= thr ? pe[2 * r] * inv_keep : 0.f;
          pe[2 * r + 1] = r1 >= thr ? pe[2 * r + 1] * inv_keep : 0.f;
        }
      }
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(pe[0], pe[1]);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(pe[2], pe[3]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      o_acc[i][0] *= corr[0]; o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1]; o_acc[i][3] *= corr[1];
    }
    // O += P V   (A = P [16 x 64 keys], B = V [keys x dh] via ldmatrix.trans)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t vf[4];
        // matrices: (keys j*16+0..7, dh np*16+0..7), (keys +8, same dh), (keys 0..7, dh +8), (keys +8, dh +8)
        ldsm_x4_t(vf, sV.at(j * 16 + (lane & 15), np * 16 + (lane >> 4) * 8));
        mma_bf16_16816(o_acc[2 * np], pf[j], vf);
        mma_bf16_16816(o_acc[2 * np + 1], pf[j], vf + 2);
      }
    }
    __syncthreads();  // everyone is done with buffer `cur` before the next prefetch overwrites it
  }
  // finalize: l over the 4 lanes of a row quad
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  bf16* ob = p.o + (long long)b * p.o_bs + h * DH;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (qrow[r] < p.Tq) {
      const float inv = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        *reinterpret_cast<uint32_t*>(ob + (long long)qrow[r] * p.o_rs + i * 8 + 2 * t) =
            pack_bf16x2(o_acc[i][2 * r] * inv, o_acc[i][2 * r + 1] * inv);
      }
      if (t == 0 && p.lse)
        p.lse[(long long)bh * p.Tq + qrow[r]] =
            l_run[r] > 0.f ? (m_run[r] + log2f(l_run[r])) * 0.6931471805599453f : -INFINITY;  // natural log units
    }
  }
}

// ============================================================================================ backward, pass 1: dQ (+ delta)
// grid (q blocks, B*H).  Recomputes S/P for each key block, dP = dO V^T, dS = P*(dP - delta), dQ += dS K.
template <int DH>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(AttnParams p) {
  constexpr int KS = DH / 16;
  constexpr int NT = DH / 8;
  constexpr int NBUF = (DH <= 64) ? 2 : 1;
  __shared__ __align__(16) Tile<64, DH> sDO;  // holds Q first (fragments go to registers), then dO
  __shared__ __align__(16) Tile<64, DH> sKb[NBUF];
  __shared__ __align__(16) Tile<64, DH> sVb[NBUF];
  __shared__ uint8_t sMaskb[NBUF][64];
  __shared__ int sFull[NBUF][2];

  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int q0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const bf16* qb = p.q + (long long)b * p.q_bs + h * DH;
  const bf16* kb = p.k + (long long)b * p.k_bs + h * DH;
  const bf16* vb = p.v + (long long)b * p.v_bs + h * DH;
  const bf16* dob = p.d_o + (long long)b * p.do_bs + h * DH;
  const bf16* ob = p.o + (long long)b * p.o_bs + h * DH;

  uint32_t qf[KS][4], dof[KS][4];
  sDO.load_async(qb, p.q_rs, q0, p.Tq);
  cp_async_wait_all();
  __syncthreads();
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) ldsm_x4(qf[kk], sDO.at(warp * 16 + (lane & 15), kk * 16 + (lane >> 4) * 8));
  __syncthreads();
  sDO.load_async(dob, p.do_rs, q0, p.Tq);
  cp_async_wait_all();
  __syncthreads();
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) ldsm_x4(dof[kk], sDO.at(warp * 16 + (lane & 15), kk * 16 + (lane >> 4) * 8));
  const int qrow[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};
  // delta = rowsum(dO * O) for this warp's 16 rows (each quad of lanes owns a row pair)
  float delta[2] = {0.f, 0.f}, lse[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (qrow[r] < p.Tq) {
      for (int c = t * 2; c < DH; c += 8) {
        const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(ob + (long long)qrow[r] * p.o_rs + c));
        const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sDO.at(qrow[r] - q0, c)));
        delta[r] += a.x * d.x + a.y * d.y;
      }
    }
    delta[r] += __shfl_xor_sync(0xffffffffu, delta[r], 1);
    delta[r] += __shfl_xor_sync(0xffffffffu, delta[r], 2);
    lse[r] = qrow[r] < p.Tq ? p.lse[(long long)bh * p.Tq + qrow[r]] : 0.f;
    if (t == 0 && qrow[r] < p.Tq) p.delta[(long long)bh * p.Tq + qrow[r]] = delta[r];
    lse[r] *= 1.4426950408889634f;  // to log2 units
  }
  float dq_acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) dq_acc[i][0] = dq_acc[i][1] = dq_acc[i][2] = dq_acc[i][3] = 0.f;
  const float sl2 = p.scale * 1.4426950408889634f;
  const unsigned long long off_eff = p.offset + ((p.p_drop > 0.f && p.offset_ptr) ? __ldg(p.offset_ptr) : 0ull);
  const uint32_t dkey = attn_drop_key(p.seed, off_eff, bh);
  const uint32_t thr = (uint32_t)(p.p_drop * 4294967296.0f);
  const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;

  int k_end = p.Sk;
  if (p.causal) k_end = min(p.Sk, q0 + 64);
  const int nblk = (k_end + 63) / 64;
  auto issue = [&](int blk, int buf) {
    const int kk0 = blk * 64;
    sKb[buf].load_async(kb, p.k_rs, kk0, p.Sk);
    sVb[buf].load_async(vb, p.v_rs, kk0, p.Sk);
    if (threadIdx.x < 64) {
      const int kk = kk0 + threadIdx.x;
      const bool ok = (kk < p.Sk) && (!p.kmask || p.kmask[(long long)b * p.Sk + kk]);
      sMaskb[buf][threadIdx.x] = ok;
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) sFull[buf][warp] = (bal == 0xffffffffu);
    }
    cp_async_commit();
  };
  __syncthreads();  // the delta pass above read sDO; K/V buffers are free
  if (NBUF == 2 && nblk > 0) issue(0, 0);
  for (int it = 0; it < nblk; ++it) {
    const int k0 = it * 64;
    const int cur = (NBUF == 2) ? (it & 1) : 0;
    if (NBUF == 2) {
      if (it + 1 < nblk) { issue(it + 1, cur ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    } else {
      issue(it, 0);
      cp_async_wait<0>();
    }
    __syncthreads();
    Tile<64, DH>& sK = sKb[cur];
    Tile<64, DH>& sV = sVb[cur];
    const uint8_t* sMask = sMaskb[cur];
    const bool nomask = sFull[cur][0] && sFull[cur][1] && (!p.causal || (k0 + 63 <= q0)) && (q0 + 64 <= p.Tq);
    float s[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kf[4], vf[4];
        ldsm_x4(kf, sK.at(np * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 16 + ((lane >> 3) & 1) * 8));
        ldsm_x4(vf, sV.at(np * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 16 + ((lane >> 3) & 1) * 8));
        mma_bf16_16816(s[2 * np], qf[kk], kf);
        mma_bf16_16816(s[2 * np + 1], qf[kk], kf + 2);
        mma_bf16_16816(dp[2 * np], dof[kk], vf);       // dP = dO V^T  (V rows are the "n" index)
        mma_bf16_16816(dp[2 * np + 1], dof[kk], vf + 2);
      }
    }
    uint32_t dsf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kc = nt * 8 + 2 * t + (e & 1);
        const int r = e >> 1;
        bool ok = true;
        if (!nomask) {
          ok = sMask[kc] && (qrow[r] < p.Tq);
          if (p.causal) ok = ok && (k0 + kc <= qrow[r]);
        }
        const float pe = ok ? exp2f(s[nt][e] * sl2 - lse[r]) : 0.f;
        float dpe = dp[nt][e];
        if (p.p_drop > 0.f) {
          const uint32_t rv = attn_drop_rand(dkey, qrow[r], k0 + kc, p.Sk);
          dpe = rv >= thr ? dpe * inv_keep : 0.f;
        }
        ds[e] = pe * (dpe - delta[r]) * p.scale;
      }
      dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(ds[0], ds[1]);
      dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
    // dQ += dS K   (B = K [keys x dh] via ldmatrix.trans)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t kf[4];
        ldsm_x4_t(kf, sK.at(j * 16 + (lane & 15), np * 16 + (lane >> 4) * 8));
        mma_bf16_16816(dq_acc[2 * np], dsf[j], kf);
        mma_bf16_16816(dq_acc[2 * np + 1], dsf[j], kf + 2);
      }
    }
    __syncthreads();
  }
  bf16* dqb = p.dq + (long long)b * p.dq_bs + h * DH;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (qrow[r] < p.Tq) {
#pragma unroll
      for (int i = 0; i < NT; ++i)
        *reinterpret_cast<uint32_t*>(dqb + (long long)qrow[r] * p.dq_rs + i * 8 + 2 * t) =
            pack_bf16x2(dq_acc[i][2 * r], dq_acc[i][2 * r + 1]);
    }
  }
}

// ============================================================================================ backward, pass 2: dK, dV
// grid (key blocks, B*H).  Each warp owns 16 keys; computes S^T = K Q^T so that P^T / dS^T are directly A operands.
//   dV += P^T dO,  dK += dS^T Q  accumulated over query blocks.
template <int DH>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(AttnParams p) {
  constexpr int KS = DH / 16;
  constexpr int NT = DH / 8;
  constexpr int NBUF = (DH <= 64) ? 2 : 1;
  __shared__ __align__(16) Tile<64, DH> sQb[NBUF];   // buffer 0 holds K first (fragments go to registers), then Q blocks
  __shared__ __align__(16) Tile<64, DH> sDOb[NBUF];  // buffer 0 holds V first, then dO blocks
  __shared__ float sLseb[NBUF][64], sDeltab[NBUF][64];

  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int k0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const bf16* qb = p.q + (long long)b * p.q_bs + h * DH;
  const bf16* kb = p.k + (long long)b * p.k_bs + h * DH;
  const bf16* vb = p.v + (long long)b * p.v_bs + h * DH;
  const bf16* dob = p.d_o + (long long)b * p.do_bs + h * DH;

  sQb[0].load_async(kb, p.k_rs, k0, p.Sk);
  sDOb[0].load_async(vb, p.v_rs, k0, p.Sk);
  cp_async_wait_all();
  __syncthreads();
  uint32_t kf[KS][4], vf[KS][4];
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    ldsm_x4(kf[kk], sQb[0].at(warp * 16 + (lane & 15), kk * 16 + (lane >> 4) * 8));
    ldsm_x4(vf[kk], sDOb[0].at(warp * 16 + (lane & 15), kk * 16 + (lane >> 4) * 8));
  }
  __syncthreads();  // K/V fragments are in registers: buffer 0 may be refilled with Q / dO blocks
  const int krow[2] = {k0 + warp * 16 + g, k0 + warp * 16 + g + 8};
  bool kok[2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
    kok[r] = (krow[r] < p.Sk) && (!p.kmask || p.kmask[(long long)b * p.Sk + krow[r]]);

  float dk_acc[NT][4], dv_acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    dk_acc[i][0] = dk_acc[i][1] = dk_acc[i][2] = dk_acc[i][3] = 0.f;
    dv_acc[i][0] = dv_acc[i][1] = dv_acc[i][2] = dv_acc[i][3] = 0.f;
  }
  const float sl2 = p.scale * 1.4426950408889634f;
  const unsigned long long off_eff = p.offset + ((p.p_drop > 0.f && p.offset_ptr) ? __ldg(p.offset_ptr) : 0ull);
  const uint32_t dkey = attn_drop_key(p.seed, off_eff, bh);
  const uint32_t thr = (uint32_t)(p.p_drop * 4294967296.0f);
  const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;

  int q_begin = 0;
  if (p.causal) q_begin = (k0 / 64) * 64;  // queries before this key block never attend to it
  const int nblk = (p.Tq - q_begin + 63) / 64;
  auto issue = [&](int blk, int buf) {
    const int qq0 = q_begin + blk * 64;
    sQb[buf].load_async(qb, p.q_rs, qq0, p.Tq);
    sDOb[buf].load_async(dob, p.do_rs, qq0, p.Tq);
    if (threadIdx.x < 64) {
      const int qq = qq0 + threadIdx.x;
      sLseb[buf][threadIdx.x] = qq < p.Tq ? p.lse[(long long)bh * p.Tq + qq] * 1.4426950408889634f : INFINITY;
      sDeltab[buf][threadIdx.x] = qq < p.Tq ? p.delta[(long long)bh * p.Tq + qq] : 0.f;
    }
    cp_async_commit();
  };
  const bool keys_ok = __all_sync(0xffffffffu, kok[0] && kok[1]);
  if (NBUF == 2 && nblk > 0) issue(0, 0);
  for (int it = 0; it < nblk; ++it) {
    const int q0 = q_begin + it * 64;
    const int cur = (NBUF == 2) ? (it & 1) : 0;
    if (NBUF == 2) {
      if (it + 1 < nblk) { issue(it + 1, cur ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    } else {
      issue(it, 0);
      cp_async_wait<0>();
    }
    __syncthreads();
    Tile<64, DH>& sQ = sQb[cur];
    Tile<64, DH>& sDO = sDOb[cur];
    const float* sLse = sLseb[cur];
    const float* sDelta = sDeltab[cur];
    // rows beyond Tq carry lse = +inf (probability 0), so only key validity and the causal diagonal need predicates
    const bool nomask = keys_ok && (!p.causal || (q0 >= k0 + 63));
    // S^T[key, query] and dP^T[key, query]
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
      dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t qf[4], dof[4];
        ldsm_x4(qf, sQ.at(np * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 16 + ((lane >> 3) & 1) * 8));
        ldsm_x4(dof, sDO.at(np * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 16 + ((lane >> 3) & 1) * 8));
        mma_bf16_16816(st[2 * np], kf[kk], qf);
        mma_bf16_16816(st[2 * np + 1], kf[kk], qf + 2);
        mma_bf16_16816(dpt[2 * np], vf[kk], dof);
        mma_bf16_16816(dpt[2 * np + 1], vf[kk], dof + 2);
      }
    }
    uint32_t ptf[4][4], dstf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float pe[4], ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qc = nt * 8 + 2 * t + (e & 1);  // query column within the block
        const int r = e >> 1;                     // key row
        const int qq = q0 + qc;
        bool ok = true;
        if (!nomask) {
          ok = kok[r] && (qq < p.Tq);
          if (p.causal) ok = ok && (krow[r] <= qq);
        }
        float pv = ok ? exp2f(st[nt][e] * sl2 - sLse[qc]) : 0.f;
        float dpe = dpt[nt][e];
        float pdrop = pv;
        if (p.p_drop > 0.f) {
          const bool keep = attn_drop_rand(dkey, qq, krow[r], p.Sk) >= thr;
          dpe = keep ? dpe * inv_keep : 0.f;
          pdrop = keep ? pv * inv_keep : 0.f;
        }
        pe[e] = pdrop;
        ds[e] = pv * (dpe - sDelta[qc]) * p.scale;
      }
      ptf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(pe[0], pe[1]);
      ptf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(pe[2], pe[3]);
      dstf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(ds[0], ds[1]);
      dstf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
    // dV += P^T dO ; dK += dS^T Q   (B operands [queries x dh] via ldmatrix.trans)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t bq[4], bdo[4];
        ldsm_x4_t(bdo, sDO.at(j * 16 + (lane & 15), np * 16 + (lane >> 4) * 8));
        ldsm_x4_t(bq, sQ.at(j * 16 + (lane & 15), np * 16 + (lane >> 4) * 8));
        mma_bf16_16816(dv_acc[2 * np], ptf[j], bdo);
        mma_bf16_16816(dv_acc[2 * np + 1], ptf[j], bdo + 2);
        mma_bf16_16816(dk_acc[2 * np], dstf[j], bq);
        mma_bf16_16816(dk_acc[2 * np + 1], dstf[j], bq + 2);
      }
    }
    __syncthreads();
  }
  bf16* dkb = p.dk + (long long)b * p.dk_bs + h * DH;
  bf16* dvb = p.dv + (long long)b * p.dv_bs + h * DH;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (krow[r] < p.Sk) {
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        *reinterpret_cast<uint32_t*>(dkb + (long long)krow[r] * p.dk_rs + i * 8 + 2 * t) =
            pack_bf16x2(dk_acc[i][2 * r], dk_acc[i][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(dvb + (long long)krow[r] * p.dv_rs + i * 8 + 2 * t) =
            pack_bf16x2(dv_acc[i][2 * r], dv_acc[i][2 * r + 1]);
      }
    }
  }
}

int attention_bwd_tc_dispatch(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                              const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                              const void* d_o, long long do_bs, long long do_rs, const float* lse, float* delta, void* dq,
                              long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs, void* dv,
                              long long dv_bs, long long dv_rs, const uint8_t* kmask, int B, int H, int Tq, int Sk, int DH,
                              int causal, float scale, float p_drop, unsigned long long seed, unsigned long long offset,
                              const unsigned long long* rng_offset_ptr, cudaStream_t stream);

int attention_fwd_tc_dispatch(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                              const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse,
                              const uint8_t* kmask, int B, int H, int Tq, int Sk, int DH, int causal, float scale, float p_drop,
                              unsigned long long seed, unsigned long long offset, const unsigned long long* rng_offset_ptr,
                              cudaStream_t stream);

static int attn_tc_mode() {   // VLM_ATTN_TC: 1 (default) = tcgen05 kernels for supported shapes, 2 = tcgen05 backward
  static int mode = -1;       // only, 0 = mma.sync kernels everywhere
  if (mode < 0) {
    const char* env = getenv("VLM_ATTN_TC");
    mode = env ? atoi(env) : 1;
  }
  return mode;
}

static int check_attn_common(const char* who, int B, int H, int Tq, int Sk, int DH) {
  if (B <= 0 || H <= 0 || Tq <= 0 || Sk <= 0) { set_error("%s: bad shape B=%d H=%d Tq=%d Sk=%d", who, B, H, Tq, Sk); return -1; }
  if (DH != 48 && DH != 64 && DH != 96) { set_error("%s: head dim %d unsupported (48, 64, 96)", who, DH); return -1; }
  return 0;
}

}  // namespace vlm

using namespace vlm;

extern "C" int vlm_attention_fwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                 const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                                 float* lse, const uint8_t* kmask, int B, int H, int Tq, int Sk, int DH, int causal,
                                 float scale, float p_drop, unsigned long long seed, unsigned long long offset,
                                 const unsigned long long* rng_offset_ptr, void* stream) {
  if (check_attn_common("vlm_attention_fwd", B, H, Tq, Sk, DH)) return -1;
  VLM_REQUIRE(q && k && v && o, "vlm_attention_fwd: null pointer");
  VLM_REQUIRE(q_rs % 8 == 0 && k_rs % 8 == 0 && v_rs % 8 == 0 && o_rs % 2 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0,
              "vlm_attention_fwd: strides must keep 16B alignment");
  if (attn_tc_mode() == 1 && lse) {
    const int r = attention_fwd_tc_dispatch(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, lse, kmask, B, H, Tq, Sk, DH,
                                            causal, scale, p_drop, seed, offset, rng_offset_ptr, (cudaStream_t)stream);
    if (r < 0) return r;
    if (r == 1) return 0;
  }
  AttnParams p = {};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.o = (bf16*)o; p.lse = lse; p.kmask = kmask;
  p.q_bs = q_bs; p.q_rs = q_rs; p.k_bs = k_bs; p.k_rs = k_rs; p.v_bs = v_bs; p.v_rs = v_rs; p.o_bs = o_bs; p.o_rs = o_rs;
  p.B = B; p.H = H; p.Tq = Tq; p.Sk = Sk; p.causal = causal; p.scale = scale; p.p_drop = p_drop; p.seed = seed; p.offset = offset; p.offset_ptr = rng_offset_ptr;
  dim3 grid((Tq + 63) / 64, B * H);
  cudaStream_t s = (cudaStream_t)stream;
  if (DH == 48) attn_fwd_kernel<48><<<grid, 128, 0, s>>>(p);
  else if (DH == 64) attn_fwd_kernel<64><<<grid, 128, 0, s>>>(p);
  else attn_fwd_kernel<96><<<grid, 128, 0, s>>>(p);
  return check_launch("attention_fwd");
}

extern "C" int vlm_attention_bwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                 const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                                 const void* d_o, long long do_bs, long long do_rs, const float* lse, float* delta,
                                 void* dq, long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs,
                                 void* dv, long long dv_bs, long long dv_rs, const uint8_t* kmask, int B, int H, int Tq,
                                 int Sk, int DH, int causal, float scale, float p_drop, unsigned long long seed,
                                 unsigned long long offset, const unsigned long long* rng_offset_ptr, void* stream) {
  if (check_attn_common("vlm_attention_bwd", B, H, Tq, Sk, DH)) return -1;
  VLM_REQUIRE(q && k && v && o && d_o && lse && delta && dq && dk && dv, "vlm_attention_bwd: null pointer");
  {
    // tcgen05 path (head dim 64, Tq <= 256): VLM_ATTN_TC=1 routes supported shapes to attention_tc.cu
    if (attn_tc_mode() >= 1) {
      const int r = attention_bwd_tc_dispatch(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, d_o, do_bs, do_rs, lse,
                                              delta, dq, dq_bs, dq_rs, dk, dk_bs, dk_rs, dv, dv_bs, dv_rs, kmask, B, H, Tq, Sk, DH,
                                              causal, scale, p_drop, seed, offset, rng_offset_ptr, (cudaStream_t)stream);
      if (r < 0) return r;
      if (r == 1) return 0;
    }
  }
  VLM_REQUIRE(q_rs % 8 == 0 && k_rs % 8 == 0 && v_rs % 8 == 0 && do_rs % 8 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 &&
                  v_bs % 8 == 0 && do_bs % 8 == 0 && o_rs % 2 == 0 && dq_rs % 2 == 0 && dk_rs % 2 == 0 && dv_rs % 2 == 0,
              "vlm_attention_bwd: strides must keep 16B alignment");
  AttnParams p = {};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.o = (bf16*)const_cast<void*>(o);
  p.lse = const_cast<float*>(lse); p.kmask = kmask;
  p.q_bs = q_bs; p.q_rs = q_rs; p.k_bs = k_bs; p.k_rs = k_rs; p.v_bs = v_bs; p.v_rs = v_rs; p.o_bs = o_bs; p.o_rs = o_rs;
  p.B = B; p.H = H; p.Tq = Tq; p.Sk = Sk; p.causal = causal; p.scale = scale; p.p_drop = p_drop; p.seed = seed; p.offset = offset; p.offset_ptr = rng_offset_ptr;
  p.d_o = (const bf16*)d_o; p.do_bs = do_bs; p.do_rs = do_rs;
  p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv;
  p.dq_bs = dq_bs; p.dq_rs = dq_rs; p.dk_bs = dk_bs; p.dk_rs = dk_rs; p.dv_bs = dv_bs; p.dv_rs = dv_rs;
  p.delta = delta;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 gq((Tq + 63) / 64, B * H), gk((Sk + 63) / 64, B * H);
  if (DH == 48) { attn_bwd_dq_kernel<48><<<gq, 128, 0, s>>>(p); attn_bwd_dkv_kernel<48><<<gk, 128, 0, s>>>(p); }
  else if (DH == 64) { attn_bwd_dq_kernel<64><<<gq, 128, 0, s>>>(p); attn_bwd_dkv_kernel<64><<<gk, 128, 0, s>>>(p); }
  else { attn_bwd_dq_kernel<96><<<gq, 128, 0, s>>>(p); attn_bwd_dkv_kernel<96><<<gk, 128, 0, s>>>(p); }
  return check_launch("attention_bwd");
}

// Explicit entry to the tcgen05 backward (tests / benchmarks); fails if the shape is outside its envelope.
extern "C" int vlm_attention_bwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                    const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                                    const void* d_o, long long do_bs, long long do_rs, const float* lse, float* delta,
                                    void* dq, long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs,
                                    void* dv, long long dv_bs, long long dv_rs, const uint8_t* kmask, int B, int H, int Tq,
                                    int Sk, int DH, int causal, float scale, float p_drop, unsigned long long seed,
                                    unsigned long long offset, const unsigned long long* rng_offset_ptr, void* stream) {
  if (check_attn_common("vlm_attention_bwd_tc", B, H, Tq, Sk, DH)) return -1;
  VLM_REQUIRE(q && k && v && o && d_o && lse && delta && dq && dk && dv, "vlm_attention_bwd_tc: null pointer");
  VLM_REQUIRE(q_rs % 8 == 0 && k_rs % 8 == 0 && v_rs % 8 == 0 && do_rs % 8 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 &&
                  v_bs % 8 == 0 && do_bs % 8 == 0, "vlm_attention_bwd_tc: strides must keep 16B alignment");
  const int r = attention_bwd_tc_dispatch(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, d_o, do_bs, do_rs, lse, delta,
                                          dq, dq_bs, dq_rs, dk, dk_bs, dk_rs, dv, dv_bs, dv_rs, kmask, B, H, Tq, Sk, DH, causal,
                                          scale, p_drop, seed, offset, rng_offset_ptr, (cudaStream_t)stream);
  if (r == 0) {
    set_error("vlm_attention_bwd_tc: shape outside the tcgen05 envelope (DH=64, Tq<=256, Tq<=128 with dropout)");
    return -1;
  }
  return r < 0 ? r : 0;
}

extern "C" int vlm_attention_fwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                                    const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                                    float* lse, const uint8_t* kmask, int B, int H, int Tq, int Sk, int DH, int causal,
                                    float scale, float p_drop, unsigned long long seed, unsigned long long offset,
                                    const unsigned long long* rng_offset_ptr, void* stream) {
  if (check_attn_common("vlm_attention_fwd_tc", B, H, Tq, Sk, DH)) return -1;
  VLM_REQUIRE(q && k && v && o && lse, "vlm_attention_fwd_tc: null pointer");
  VLM_REQUIRE(q_rs % 8 == 0 && k_rs % 8 == 0 && v_rs % 8 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0,
              "vlm_attention_fwd_tc: strides must keep 16B alignment");
  const int r = attention_fwd_tc_dispatch(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, lse, kmask, B, H, Tq, Sk, DH,
                                          causal, scale, p_drop, seed, offset, rng_offset_ptr, (cudaStream_t)stream);
  if (r == 0) {
    set_error("vlm_attention_fwd_tc: shape outside the tcgen05 envelope (DH=64, Sk<=224)");
    return -1;
  }
  return r < 0 ? r : 0;
}
