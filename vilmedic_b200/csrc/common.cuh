// Shared device/host helpers for the vilmedic_b200 sm_100a kernels.
// PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), fences.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>

namespace vlm {

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define VLM_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::vlm::set_error(__VA_ARGS__);      \
      return -1;                          \
    }                                     \
  } while (0)

int num_sms();
int num_sms_all();      // physical SM count (ignores the margin)
int background_mode();  // runtime.cu: helper kernels launch as small co-resident CTAs while set

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Parity wait with a watchdog: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  long long t0 = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins == 1024u) t0 = clock64();
    if (spins > 1024u && (spins & 1023u) == 0u && clock64() - t0 > 4000000000LL) {
      printf("vlm: mbarrier watchdog block %d thread %d bar %u parity %u\n", (int)blockIdx.x, (int)threadIdx.x, addr,
             parity);
      __trap();
    }
  }
}
// TMA tensor store (shared::cta -> global) + bulk async-group bookkeeping.
__device__ __forceinline__ void tma_store_3d(const void* desc, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// named barrier over a subset of the CTA (all participating warps must use the same id / thread count)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* desc, uint64_t* bar, void* smem_dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const void* desc, uint64_t* bar, void* smem_dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- tcgen05 ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"((uint32_t)NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)NCOLS) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t gets lane (base_lane + t), columns [col, col+32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp of CUTLASS; restated, not copied) ----
// Shared-memory matrix descriptor for a 128B-swizzled tile whose rows are 128 bytes (64 bf16).
//   K-major operand : tile = [rows (M or N)][64 k] ; 8-row groups 1024 B apart  -> SBO = 1024, LBO unused.
//   MN-major operand: tile = [64-col MN block][k rows][64 mn]; k rows 128 B apart, 8-k groups 1024 B apart (SBO),
//                     MN blocks `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);          // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;    // [16,30) leading byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;    // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                               // [46,48) descriptor version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;                               // [61,64) layout type 2 = SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16, BF16 x BF16 -> F32.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                         // c_format = F32
         | (1u << 7)                       // a_format = BF16
         | (1u << 10)                      // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)  // a_major
         | ((b_mn_major ? 1u : 0u) << 16)  // b_major
         | ((uint32_t)(N >> 3) << 17)      // n_dim
         | ((uint32_t)(M >> 4) << 24);     // m_dim
}

// ---- misc math ----
// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the bf16 output resolution): one MUFU.RCP + one
// MUFU.EX2 + 7 FMA-class ops instead of the ~25-instruction branchy erff().  `e` returns exp(-x^2) for reuse.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float erf_as(float x, float& e) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));      // MUFU.RCP (2 ulp): abs error of erf stays < 3e-7
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  e = ex2_approx(ax * ax * -1.4426950408889634f);
  return copysignf(fmaf(-p, e, 1.0f), x);
}
// exact-erf GELU of HF (hidden_act="gelu"): 0.5 x (1 + erf(x / sqrt 2)), with erf by the same A&S 7.1.26 polynomial evaluated
// directly on |x| (constants folded):  w = 0.5 (1 - erf(|x| / sqrt 2)) = 0.5 P(t) exp(-x^2 / 2),  t = 1 / (1 + p |x| / sqrt 2).
// `e` returns exp(-x^2 / 2).  GELU(x) = relu(x) - |x| w  (x >= 0: x (1 - w); x < 0: x w — no cancellation in the negative tail).
__device__ __forceinline__ float gelu_half_erfc(float ax, float& e) {
  const float t = rcp_approx(fmaf(ax, 0.3275911f * 0.70710678118654752f, 1.0f));
  float q = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  q = fmaf(q, t, 0.5f * 1.421413741f);
  q = fmaf(q, t, 0.5f * -0.284496736f);
  q = fmaf(q, t, 0.5f * 0.254829592f);
  e = ex2_approx((ax * ax) * (-0.5f * 1.4426950408889634f));
  return (q * t) * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float ax = fabsf(x);
  const float w = gelu_half_erfc(ax, e);
  return fmaf(-ax, w, fmaxf(x, 0.f));
}
// d/dx GELU = Phi(x) + x phi(x);  Phi = 1 - w (x >= 0) | w (x < 0);  phi = exp(-x^2 / 2) / sqrt(2 pi)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;
  const float w = gelu_half_erfc(fabsf(x), e);
  const float cdf = 0.5f + copysignf(0.5f - w, x);
  return fmaf(x * 0.39894228040143268f, e, cdf);
}

// GELU and its derivative from one evaluation of the shared terms (the forward epilogue stashes the DERIVATIVE, so the
// backward epilogue is a plain multiply): 5 instructions on top of gelu_erf.
__device__ __forceinline__ float gelu_erf_both(float x, float& dgelu) {
  float e;
  const float ax = fabsf(x);
  const float w = gelu_half_erfc(ax, e);
  dgelu = fmaf(x * 0.39894228040143268f, e, 0.5f + copysignf(0.5f - w, x));
  return fmaf(-ax, w, fmaxf(x, 0.f));
}

// The same evaluation for two elements at
// a time on the packed fp32x2 FMA pipe of sm_100 (FFMA2 / FMUL2 / FADD2): 13 packed + 10 scalar instructions per pair instead of
// 2 x 19 scalar ones.  The polynomial carries the sign (wn = -w), so that GELU = fma(|x|, wn, relu(x)).
__device__ __forceinline__ void gelu_erf_both_x2(float x0, float x1, float& g0, float& g1, float& d0, float& d1) {
  const float2 x = make_float2(x0, x1);
  const float2 ax = make_float2(fabsf(x0), fabsf(x1));
  const float2 one = make_float2(1.0f, 1.0f), half = make_float2(0.5f, 0.5f);
  const float2 den = __ffma2_rn(ax, make_float2(0.3275911f * 0.70710678118654752f, 0.3275911f * 0.70710678118654752f), one);
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  float2 q = __ffma2_rn(make_float2(-0.5f * 1.061405429f, -0.5f * 1.061405429f), t, make_float2(0.5f * 1.453152027f, 0.5f * 1.453152027f));
  q = __ffma2_rn(q, t, make_float2(-0.5f * 1.421413741f, -0.5f * 1.421413741f));
  q = __ffma2_rn(q, t, make_float2(0.5f * 0.284496736f, 0.5f * 0.284496736f));
  q = __ffma2_rn(q, t, make_float2(-0.5f * 0.254829592f, -0.5f * 0.254829592f));
  const float2 arg = __fmul2_rn(__fmul2_rn(ax, ax), make_float2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f));
  const float2 e = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));
  const float2 wn = __fmul2_rn(__fmul2_rn(q, t), e);                            // -0.5 (1 - erf(|x| / sqrt 2))
  const float2 g = __ffma2_rn(ax, wn, make_float2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  const float2 hp = __fadd2_rn(half, wn);                                       // 0.5 - w
  const float2 cdf = __fadd2_rn(make_float2(copysignf(hp.x, x0), copysignf(hp.y, x1)), half);
  const float2 d = __ffma2_rn(__fmul2_rn(x, make_float2(0.39894228040143268f, 0.39894228040143268f)), e, cdf);
  g0 = g.x; g1 = g.y; d0 = d.x; d1 = d.y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// Attention-probability dropout uses a cheap per-element counter hash (one decision per (batch*head, query, key)): the
// tcgen05 backward walks the score matrix transposed, so a generator that amortises over 4 consecutive keys (Philox)
// would cost a full call per element there.  key = splitmix64(seed, offset, bh); word = lowbias32(key ^ elem * phi).
__device__ __forceinline__ uint32_t attn_drop_key(unsigned long long seed, unsigned long long offset, int bh) {
  unsigned long long z = seed ^ (offset * 0x9E3779B97F4A7C15ull) ^ ((unsigned long long)(bh + 1) * 0xD6E8FEB86659FD93ull);
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)z ^ (uint32_t)(z >> 32);
}
// Round 2: one multiply round instead of two (the softmax warps of the decoder attention kernels were integer-issue bound on this
// hash: LOP3 / ISETP / IMAD / SHF were 50 % of their instructions, profiles/ncu_r1c_attention_fwd.txt).  The word is only compared with
// a threshold, i.e. its high bits matter; Weyl-sequence input + xorshift-multiply keeps those equidistributed (tests/test_ops_gpu.py
// checks the drop rate, the independence across rows / keys / heads and the fwd = bwd mask identity).
__device__ __forceinline__ uint32_t attn_drop_rand(uint32_t key, int q, int k, int Sk) {
  uint32_t x = ((uint32_t)q * (uint32_t)Sk + (uint32_t)k) * 0x9E3779B1u ^ key;
  x ^= x >> 15; x *= 0x2C1B3C6Du;
  x ^= x >> 12;
  return x;
}

// Counter-based RNG for dropout: Philox4x32-10.  (seed, offset) fixed per launch; `idx` = element group index.
struct Philox {
  uint32_t key0, key1;
  __device__ __forceinline__ Philox(uint64_t seed) : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint64_t idx, uint64_t offset) const {
    uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ k0;
      c1 = lo1;
      c2 = hi0 ^ c3 ^ k1;
      c3 = lo0;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

}  // namespace vlm
