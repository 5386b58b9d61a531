// Data-parallel gradient exchange over NVLink / NVSwitch PEER MEMORY, fused into the optimizer kernel.
//
// The reference all-reduces fp32 gradients with NCCL (accelerate / DistributedDataParallel,
// vilmedic/executors/trainor_accelerate.py:122,132) and then runs torch.optim on every rank.  Here there is no collective
// kernel at all: every rank publishes the bf16 cast of a gradient bucket in a buffer that its peers have mapped through CUDA
// IPC, raises a flag in every peer's memory, and the fused optimizer kernel of each rank READS THE BUCKET OF ALL RANKS through
// peer loads (one 8-byte load per rank and 4 parameters), sums them in fp32 in rank order (so that all replicas stay bit-identical)
// and applies Adam / AdamW / RAdam in the same pass — the reduction never exists in HBM.  The kernel is launched bucket by
// bucket from the backward pass (vilmedic_b200/ddp.py) as small background CTAs that sit next to the persistent tensor kernels,
// so its NVLink reads run under the backward of the layers below.  NCCL's all-reduce kernels could not do that: their 32 CTAs do
// not fit on an SM next to a persistent GEMM CTA and the exchange added ~1 ms to a 22.6 ms step on 2 GPUs (ddp.py).
//
// Protocol (epoch e = step number, kept in device memory so that the whole step replays as a CUDA graph):
//   main stream :  epoch += 1 ; wait until every peer's DONE flag >= e - 1 (they have finished READING my buffers of step e-1)
//                  per bucket b: cast fp32 gradients -> my bf16 buffer ; signal READY[b][me] = e into every rank's flag block
//   side stream :  per bucket b: optimizer kernel: poll READY[b][r] >= e for all r (local memory), then peer-load + update
//                  after the last bucket: signal DONE[me] = e into every rank's flag block
// Waits are bounded (globaltimer): on expiry the kernel raises an error flag instead of hanging the GPU.
#include <cuda_runtime.h>
#include <cstring>
#include "common.cuh"
#include "vlm_b200.h"

namespace vlm {

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

static constexpr unsigned long long P2P_TIMEOUT_NS = 90ull * 1000ull * 1000ull * 1000ull;   // 90 s: ranks may be seconds apart
                                                                                           // right after each has captured its CUDA graph

// thread w < world waits for flags[w] >= want; returns false on timeout (and raises *err)
__device__ __forceinline__ bool p2p_poll(const int* flags, int want, int* err) {
  if (ld_acquire_sys(flags) >= want) return true;
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (ld_acquire_sys(flags) < want) {
    __nanosleep(200);
    if ((++spins & 1023u) == 0u && globaltimer_ns() - t0 > P2P_TIMEOUT_NS) {
      if (err) atomicExch(err, 1);
      return false;
    }
  }
  return true;
}

static constexpr int P2P_MAX_WORLD = 16;
struct P2pFlagPtrs {
  int* f[P2P_MAX_WORLD];               // every rank's flag block (mine included), as mapped into this process
};

__global__ void p2p_signal_kernel(P2pFlagPtrs peer_flags, int world, int rank, int slot, const int* __restrict__ epoch_ptr) {
  const int w = threadIdx.x;
  if (w < world) {
    __threadfence_system();
    st_release_sys(peer_flags.f[w] + (long long)slot * world + rank, *epoch_ptr);
  }
}

__global__ void p2p_wait_kernel(const int* __restrict__ my_flags, int world, int slot, const int* __restrict__ epoch_ptr, int delta, int* err) {
  const int w = threadIdx.x;
  if (w < world) p2p_poll(my_flags + (long long)slot * world + w, *epoch_ptr + delta, err);
}

__global__ void p2p_epoch_inc_kernel(int* epoch) { *epoch += 1; }

struct P2pOptimScalars {
  float lr, beta1, beta2, eps, weight_decay, grad_scale;
};

struct P2pPeers {
  const bf16* g16[P2P_MAX_WORLD];      // every rank's bf16 gradient buffer, already offset to the first element of the span
};
struct P2pReduced {
  const float* r32[P2P_MAX_WORLD];     // every rank's fp32 buffer of reduced slices, offset to the first element of the span
};

// Two-shot exchange, first half (reduce-scatter through peer memory): this rank sums ITS slice of a bucket over all ranks' bf16
// buffers (rank order, fp32) into its own fp32 buffer; the optimizer kernels of all ranks then read each slice from its owner.
// NVLink bytes per GPU and parameter: (N-1)/N * (2 + 4) instead of (N-1) * 2 of the one-shot kernel — the choice for N >= 4.
__global__ void __launch_bounds__(256) p2p_reduce_slice_kernel(P2pPeers peers, float* __restrict__ out, long long n4, int world,
                                                               const int* __restrict__ ready_flags, const int* __restrict__ epoch_ptr,
                                                               int* err) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < world) {
    if (!p2p_poll(ready_flags + threadIdx.x, *epoch_ptr, err)) ok = 0;
  }
  __syncthreads();
  if (!ok) return;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = i0; i < n4; i += stride) {
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w0 = 0; w0 < world; w0 += 8) {
      uint2 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (w0 + k < world) u[k] = __ldcv(reinterpret_cast<const uint2*>(peers.g16[w0 + k]) + i);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (w0 + k < world) {
          const float2 a = unpack_bf16x2(u[k].x), b = unpack_bf16x2(u[k].y);
          gv.x += a.x; gv.y += a.y; gv.z += b.x; gv.w += b.y;
        }
    }
    reinterpret_cast<float4*>(out)[i] = gv;
  }
}

// KIND as in optim.cu (0 AdamW, 1 Adam, 2 RAdam); update formulas identical to optim_kernel (kept in sync by tests/test_ddp_gpu.py:
// the 2-rank step must equal the full-batch single-GPU step).
// TWO_SHOT: the gradient of 4-element unit i is read (fp32, already summed) from the rank that owns it:
// owner = (unit_base + i) / units_per_rank; `ready_flags` are then the REDUCED flags of the bucket.
// (<= 64 registers: a 128-thread background CTA must fit into the 8 K registers a resident persistent GEMM CTA leaves free)
template <int KIND, bool TWO_SHOT>
__global__ void __launch_bounds__(256, 4) optim_p2p_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                        bf16* __restrict__ p_bf16, long long n, P2pOptimScalars sc, P2pPeers peers,
                                                        P2pReduced red, long long unit_base, long long units_per_rank, int world,
                                                        const int* __restrict__ ready_flags, const int* __restrict__ epoch_ptr,
                                                        const int* __restrict__ step_ptr, const float* __restrict__ lr_scale_ptr, int* err) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < world) {
    if (!p2p_poll(ready_flags + threadIdx.x, *epoch_ptr, err)) ok = 0;
  }
  __syncthreads();
  if (!ok) return;                       // a peer never published this bucket: leave the parameters untouched (err is raised)
  const long long n4 = n / 4;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const int step = step_ptr ? *step_ptr : 1;
  const float beta1 = sc.beta1, beta2 = sc.beta2, eps = sc.eps, wd = sc.weight_decay;
  const float b1t = powf(beta1, (float)step), b2t = powf(beta2, (float)step);
  const float bc1 = 1.f - b1t, bc2 = 1.f - b2t;
  const float lr_eff = sc.lr * (lr_scale_ptr ? *lr_scale_ptr : 1.f);
  const float gs = sc.grad_scale;
  const float step_size = lr_eff / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  float rect = 0.f;
  bool rectified = false;
  const float sqrt_bc2 = sqrtf(bc2);
  if (KIND == 2) {
    const float rho_inf = 2.f / (1.f - beta2) - 1.f;
    const float rho_t = rho_inf - 2.f * (float)step * b2t / bc2;
    if (rho_t > 5.f) {
      rectified = true;
      rect = sqrtf((rho_t - 4.f) * (rho_t - 2.f) * rho_inf / ((rho_inf - 4.f) * (rho_inf - 2.f) * rho_t));
    }
  }
  // update of the 4-element unit i with the (unscaled) gradient sum gv
  auto apply = [&](long long i, float4 gv) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gj = gp[j] * gs;
      if (KIND != 0) gj = fmaf(wd, pp[j], gj);
      mp[j] = beta1 * mp[j] + (1.f - beta1) * gj;
      vp[j] = beta2 * vp[j] + (1.f - beta2) * gj * gj;
      if (KIND == 0) {
        const float denom = sqrtf(vp[j]) * inv_sqrt_bc2 + eps;
        pp[j] = pp[j] * (1.f - lr_eff * wd) - step_size * mp[j] / denom;
      } else if (KIND == 1) {
        const float denom = sqrtf(vp[j]) * inv_sqrt_bc2 + eps;
        pp[j] = pp[j] - step_size * mp[j] / denom;
      } else {
        const float mhat = mp[j] / bc1;
        if (rectified) pp[j] = pp[j] - mhat * lr_eff * (sqrt_bc2 / (sqrtf(vp[j]) + eps)) * rect;
        else pp[j] = pp[j] - mhat * lr_eff;
      }
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 o;
    o.x = pack_bf16x2(pv.x, pv.y);
    o.y = pack_bf16x2(pv.z, pv.w);
    reinterpret_cast<uint2*>(p_bf16)[i] = o;
  };
  if (TWO_SHOT) {
    // four peer loads in flight per thread: as a background CTA (128 threads next to a persistent GEMM CTA) the kernel is bound by
    // NVLink latency x bytes in flight — one 16-byte load per thread gave ~100 GB/s per GPU on 8 GPUs (tools/jobs/r2z.sh)
    const uint32_t per = (uint32_t)units_per_rank, base = (uint32_t)unit_base;
    for (long long i = i0; i < n4; i += 4 * stride) {
      float4 gq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long j = i + u * stride;
        if (j < n4) gq[u] = __ldcv(reinterpret_cast<const float4*>(red.r32[(base + (uint32_t)j) / per]) + j);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long j = i + u * stride;
        if (j < n4) apply(j, gq[u]);
      }
    }
  } else {
    for (long long i = i0; i < n4; i += stride) {
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int w0 = 0; w0 < world; w0 += 8) {      // up to 8 peer loads in flight; summation order = rank order on every rank
        uint2 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (w0 + k < world) u[k] = __ldcv(reinterpret_cast<const uint2*>(peers.g16[w0 + k]) + i);   // never from a stale L1 line
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (w0 + k < world) {
            const float2 a = unpack_bf16x2(u[k].x), b = unpack_bf16x2(u[k].y);
            gv.x += a.x; gv.y += a.y; gv.z += b.x; gv.w += b.y;
          }
      }
      apply(i, gv);
    }
  }
}

}  // namespace vlm

using namespace vlm;

extern "C" int vlm_ipc_alloc(long long bytes, void** dev_ptr, void* handle64) {
  VLM_REQUIRE(bytes > 0 && dev_ptr && handle64, "vlm_ipc_alloc: bad args");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    cudaGetLastError();
    set_error("vlm_ipc_alloc(%lld bytes): %s", bytes, cudaGetErrorString(e));
    return -1;
  }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return 0;
}

extern "C" int vlm_ipc_open(const void* handle64, void** dev_ptr) {
  VLM_REQUIRE(handle64 && dev_ptr, "vlm_ipc_open: bad args");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("vlm_ipc_open: %s", cudaGetErrorString(e));
    return -1;
  }
  *dev_ptr = p;
  return 0;
}

extern "C" int vlm_ipc_close(void* dev_ptr) {
  if (dev_ptr && cudaIpcCloseMemHandle(dev_ptr) != cudaSuccess) {
    cudaGetLastError();
    set_error("vlm_ipc_close failed");
    return -1;
  }
  return 0;
}

extern "C" int vlm_ipc_free(void* dev_ptr) {
  if (dev_ptr && cudaFree(dev_ptr) != cudaSuccess) {
    cudaGetLastError();
    set_error("vlm_ipc_free failed");
    return -1;
  }
  return 0;
}

extern "C" int vlm_p2p_epoch_inc(int* epoch_ptr, void* stream) {
  VLM_REQUIRE(epoch_ptr, "vlm_p2p_epoch_inc: null pointer");
  p2p_epoch_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(epoch_ptr);
  return check_launch("p2p_epoch_inc");
}

extern "C" int vlm_p2p_signal(void* const* peer_flags, int world, int rank, int slot, const int* epoch_ptr, void* stream) {
  VLM_REQUIRE(peer_flags && epoch_ptr && world >= 1 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world && slot >= 0,
              "vlm_p2p_signal: bad args (world=%d rank=%d slot=%d)", world, rank, slot);
  P2pFlagPtrs fp;
  for (int w = 0; w < P2P_MAX_WORLD; ++w) fp.f[w] = w < world ? reinterpret_cast<int*>(peer_flags[w]) : nullptr;
  for (int w = 0; w < world; ++w) VLM_REQUIRE(fp.f[w], "vlm_p2p_signal: flag block of rank %d missing", w);
  p2p_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(fp, world, rank, slot, epoch_ptr);
  return check_launch("p2p_signal");
}

extern "C" int vlm_p2p_wait(const int* my_flags, int world, int slot, const int* epoch_ptr, int epoch_delta, int* err_flag, void* stream) {
  VLM_REQUIRE(my_flags && epoch_ptr && world >= 1 && world <= P2P_MAX_WORLD && slot >= 0, "vlm_p2p_wait: bad args");
  p2p_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(my_flags, world, slot, epoch_ptr, epoch_delta, err_flag);
  return check_launch("p2p_wait");
}

extern "C" int vlm_p2p_reduce_slice(const void* const* peer_g16, long long elem_offset, float* out, long long n, int world,
                                    const int* my_flags, int slot, const int* epoch_ptr, int* err_flag, void* stream) {
  VLM_REQUIRE(peer_g16 && out && n > 0 && n % 4 == 0 && elem_offset % 4 == 0 && ((uintptr_t)out % 16 == 0), "vlm_p2p_reduce_slice: bad span");
  VLM_REQUIRE(my_flags && epoch_ptr && world >= 1 && world <= P2P_MAX_WORLD && slot >= 0, "vlm_p2p_reduce_slice: bad exchange args");
  P2pPeers peers;
  for (int w = 0; w < P2P_MAX_WORLD; ++w) peers.g16[w] = nullptr;
  for (int w = 0; w < world; ++w) {
    VLM_REQUIRE(peer_g16[w] && ((uintptr_t)peer_g16[w] % 8 == 0), "vlm_p2p_reduce_slice: peer buffer %d missing / misaligned", w);
    peers.g16[w] = reinterpret_cast<const bf16*>(peer_g16[w]) + elem_offset;
  }
  const int threads = background_mode() ? 128 : 256;
  long long blocks = background_mode() ? (n / 4 + 511) / 512 : (n / 4 + 255) / 256;
  const long long cap = background_mode() ? (long long)num_sms_all() * 16 : (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  p2p_reduce_slice_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(peers, out, n / 4, world, my_flags + (long long)slot * world,
                                                                            epoch_ptr, err_flag);
  return check_launch("p2p_reduce_slice");
}

extern "C" int vlm_optim_step_p2p(int kind, float* p, float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                                  float beta2, float eps, float weight_decay, const int* step_ptr, const float* lr_scale_ptr,
                                  float grad_scale, const void* const* peer_g16, const void* const* peer_r32, long long elem_offset,
                                  long long unit_base, long long units_per_rank, int world,
                                  const int* my_flags, int slot, const int* epoch_ptr, int* err_flag, void* stream) {
  VLM_REQUIRE(kind >= 0 && kind <= 2, "vlm_optim_step_p2p: kind must be 0 (AdamW) | 1 (Adam) | 2 (RAdam)");
  VLM_REQUIRE(p && g && m && v && p_bf16 && n > 0 && n % 4 == 0 && elem_offset % 4 == 0, "vlm_optim_step_p2p: flat buffers must be non-null, n %% 4 == 0");
  VLM_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) && ((uintptr_t)v % 16 == 0) &&
                  ((uintptr_t)p_bf16 % 8 == 0), "vlm_optim_step_p2p: buffers must be 16-byte aligned");
  VLM_REQUIRE((peer_g16 || peer_r32) && my_flags && epoch_ptr && world >= 1 && world <= P2P_MAX_WORLD && slot >= 0, "vlm_optim_step_p2p: bad exchange args");
  VLM_REQUIRE(!peer_r32 || (units_per_rank > 0 && unit_base >= 0), "vlm_optim_step_p2p: two-shot needs units_per_rank > 0");
  P2pPeers peers;
  P2pReduced red;
  for (int w = 0; w < P2P_MAX_WORLD; ++w) { peers.g16[w] = nullptr; red.r32[w] = nullptr; }
  for (int w = 0; w < world; ++w) {
    if (peer_r32) {
      VLM_REQUIRE(peer_r32[w] && ((uintptr_t)peer_r32[w] % 16 == 0), "vlm_optim_step_p2p: reduced buffer %d missing / misaligned", w);
      red.r32[w] = reinterpret_cast<const float*>(peer_r32[w]) + elem_offset;
    } else {
      VLM_REQUIRE(peer_g16[w] && ((uintptr_t)peer_g16[w] % 8 == 0), "vlm_optim_step_p2p: peer buffer %d missing / misaligned", w);
      peers.g16[w] = reinterpret_cast<const bf16*>(peer_g16[w]) + elem_offset;
    }
  }
  P2pOptimScalars sc{lr, beta1, beta2, eps, weight_decay, grad_scale};
  // launch shape as in optim.cu: small background CTAs under the backward pass, full-size ones for the tail
  const int threads = background_mode() ? 128 : 256;
  long long blocks = background_mode() ? (n / 4 + 511) / 512 : (n / 4 + 255) / 256;
  const long long cap = background_mode() ? (long long)num_sms_all() * 16 : (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaStream_t s = (cudaStream_t)stream;
  const int* ready = my_flags + (long long)slot * world;
#define VLM_P2P(K_, T_) optim_p2p_kernel<K_, T_><<<(int)blocks, threads, 0, s>>>(p, g, m, v, (bf16*)p_bf16, n, sc, peers, red, unit_base, units_per_rank, world, ready, epoch_ptr, step_ptr, lr_scale_ptr, err_flag)
  if (peer_r32) {
    if (kind == 0) VLM_P2P(0, true);
    else if (kind == 1) VLM_P2P(1, true);
    else VLM_P2P(2, true);
  } else {
    if (kind == 0) VLM_P2P(0, false);
    else if (kind == 1) VLM_P2P(1, false);
    else VLM_P2P(2, false);
  }
#undef VLM_P2P
  return check_launch("optim_step_p2p");
}
