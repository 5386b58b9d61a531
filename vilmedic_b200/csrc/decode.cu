// Device-side incremental decoding and beam search: the loop the reference drives from the host —
// vilmedic/blocks/huggingface/decoder/beam_search.py:243-320 (sum of the models' next-token logits :254, log_softmax :260-262,
// + running beam score :265, top-2k over k*V :289-294, BeamSearchScorer.process :297-304, cache reorder :317-319) and HF's cached
// generate step behind vilmedic/blocks/huggingface/decoder/evaluation.py:73-78 — as kernels whose step counter, tokens, scores,
// hypotheses and cache indirection all live in device memory, so that one decode step is a replayable CUDA graph.
//
//  * vlm_embed_step          word[tok] + pos[*t] for the current token of every row (position read from device memory)
//  * vlm_decode_attention    T_q = 1 attention over an INDIRECTED KV cache: row r's key/value at position j lives in physical row
//                            row_map[r][j]; the step's new K/V are appended at (r, *t).  Beam reordering therefore permutes a
//                            [rows, max_len] int32 table instead of copying the cache (beam_search.py:317-319 index_select).
//                            Cross-attention uses the same kernel with one K/V per IMAGE (row / beams), not per beam.
//  * vlm_beam_rows           per row: sum_m logits_m, log-softmax, + beam score, top-2k candidates (score desc, token asc)
//  * vlm_beam_select         per batch element: merge to the global top-2k, hypothesis / EOS bookkeeping, next tokens / parents
//  * vlm_beam_advance        ids / row_map follow their parents (double-buffered), the step counter advances
// HBM-bound integer/byte work around the weight-streaming GEMMs of the step; nothing here touches tensor cores.
#include "common.cuh"
#include "vlm_b200.h"

namespace vlm {

// ------------------------------------------------------------------------------------------------------------------ embedding
__global__ void embed_step_kernel(const long long* __restrict__ tok, const float* __restrict__ word, const float* __restrict__ pos,
                                  bf16* __restrict__ z, int R, int D, int V, const int* __restrict__ t_ptr, int max_pos, int pos_shift,
                                  const float* __restrict__ tt_row) {
  const int nvec = D / 8;
  const int t = min(*t_ptr + pos_shift, max_pos - 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * nvec; i += gridDim.x * blockDim.x) {
    const int v = i % nvec, r = i / nvec;
    long long id = tok[r];
    if (id < 0 || id >= V) id = 0;
    const float* w = word + (size_t)id * D + v * 8;
    const float* p = pos + (size_t)t * D + v * 8;
    float4 w0 = __ldg(reinterpret_cast<const float4*>(w)), w1 = __ldg(reinterpret_cast<const float4*>(w) + 1);
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(p)), p1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
    if (tt_row) {
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

// ------------------------------------------------------------------------------------------------------------------ attention
// One CTA (4 warps) per (row, head).  Phase 1: lane-per-key dot products (each lane streams one 2*DH-byte key), scores to smem,
// block max.  Phase 2: p = exp(s - max), block sum.  Phase 3: lane-per-dimension accumulation of p_j v_j (one coalesced DH*2-byte
// read per key), 4 warps over interleaved keys, combined through smem.  fp32 throughout, one rounding to bf16 at the end.
static constexpr int DA_THREADS = 128;

template <int DH>
__global__ void __launch_bounds__(DA_THREADS) decode_attn_kernel(
    const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ kv_new, long long ld_new, bf16* __restrict__ cache,
    long long cache_row_stride, int two_d, int* __restrict__ row_map, int map_ld, const int* __restrict__ t_ptr, int fixed_len,
    int row_div, const uint8_t* __restrict__ kmask, int kmask_ld, bf16* __restrict__ out, long long ldo, float scale) {
  extern __shared__ float sm[];
  const int r = blockIdx.x, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = two_d >> 1;
  const bool self = kv_new != nullptr;
  const int t = self ? *t_ptr : 0;
  const int len = self ? t + 1 : fixed_len;
  float* s = sm;                    // [len]
  float* red = sm + ((len + 3) & ~3);   // [4 * DH + 8]
  const int crow = self ? 0 : r / row_div;
  const bf16* qp = q + (long long)r * ldq + h * DH;
  // q in registers, fp32 (every lane needs all DH values for its keys)
  float qf[DH];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(qp) + i);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    qf[8 * i + 0] = a.x; qf[8 * i + 1] = a.y; qf[8 * i + 2] = b.x; qf[8 * i + 3] = b.y;
    qf[8 * i + 4] = c.x; qf[8 * i + 5] = c.y; qf[8 * i + 6] = d.x; qf[8 * i + 7] = d.y;
  }
  const bf16* new_k = self ? kv_new + (long long)r * ld_new + h * DH : nullptr;
  const bf16* new_v = self ? new_k + D : nullptr;
  if (self) {                        // append this head's slice of the new key / value at (r, t); record the indirection
    bf16* dst = cache + (long long)r * cache_row_stride + (long long)t * two_d + h * DH;
    for (int i = tid; i < DH / 8 * 2; i += DA_THREADS) {
      const int which = i / (DH / 8), c = i % (DH / 8);
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(which ? new_v : new_k) + c);
      reinterpret_cast<uint4*>(dst + which * D)[c] = u;
    }
    if (h == 0 && tid == 0) row_map[(long long)r * map_ld + t] = r;
  }
  // ---- phase 1: scores
  float lmax = -INFINITY;
  for (int j = tid; j < len; j += DA_THREADS) {
    const bf16* kp;
    if (self) kp = (j == t) ? new_k : cache + (long long)row_map[(long long)r * map_ld + j] * cache_row_stride + (long long)j * two_d + h * DH;
    else kp = cache + (long long)crow * cache_row_stride + (long long)j * two_d + h * DH;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(kp) + i);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      acc = fmaf(qf[8 * i + 0], a.x, acc); acc = fmaf(qf[8 * i + 1], a.y, acc);
      acc = fmaf(qf[8 * i + 2], b.x, acc); acc = fmaf(qf[8 * i + 3], b.y, acc);
      acc = fmaf(qf[8 * i + 4], c.x, acc); acc = fmaf(qf[8 * i + 5], c.y, acc);
      acc = fmaf(qf[8 * i + 6], d.x, acc); acc = fmaf(qf[8 * i + 7], d.y, acc);
    }
    acc *= scale;
    if (kmask && kmask[(long long)crow * kmask_ld + j] == 0) acc = -INFINITY;
    s[j] = acc;
    lmax = fmaxf(lmax, acc);
  }
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  const float m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  // ---- phase 2: probabilities
  float lsum = 0.f;
  for (int j = tid; j < len; j += DA_THREADS) {
    const float p = (m == -INFINITY) ? 0.f : __expf(s[j] - m);
    s[j] = p;
    lsum += p;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[4 + warp] = lsum;
  __syncthreads();
  const float denom = red[4] + red[5] + red[6] + red[7];
  // ---- phase 3: o[d] = sum_j p_j v_j[d]; warp w takes keys j = w, w + 4, ...
  constexpr int DPL = (DH + 31) / 32;
  float o[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) o[i] = 0.f;
#pragma unroll 4
  for (int j = warp; j < len; j += 4) {
    const float p = s[j];
    const bf16* vp;
    if (self) vp = (j == t) ? new_v : cache + (long long)row_map[(long long)r * map_ld + j] * cache_row_stride + (long long)j * two_d + D + h * DH;
    else vp = cache + (long long)crow * cache_row_stride + (long long)j * two_d + D + h * DH;
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      const int d = lane + 32 * i;
      if (d < DH) o[i] = fmaf(p, __bfloat162float(vp[d]), o[i]);
    }
  }
  float* part = red + 8;            // [4][DH]
#pragma unroll
  for (int i = 0; i < DPL; ++i) {
    const int d = lane + 32 * i;
    if (d < DH) part[warp * DH + d] = o[i];
  }
  __syncthreads();
  for (int d = tid; d < DH; d += DA_THREADS) {
    const float v = part[d] + part[DH + d] + part[2 * DH + d] + part[3 * DH + d];
    out[(long long)r * ldo + h * DH + d] = __float2bfloat16(denom > 0.f ? v / denom : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------------------------ beam: rows
struct LogitPtrs {
  const float* p[8];
};

// (score desc, index asc) ordering shared by all selection code
__device__ __forceinline__ bool cand_better(float sa, long long ia, float sb, long long ib) { return sa > sb || (sa == sb && ia < ib); }

template <int K>
__global__ void __launch_bounds__(256) beam_rows_kernel(LogitPtrs lp, int n_models, long long ld, int V, const float* __restrict__ beam_scores,
                                                       float* __restrict__ cand_score, int* __restrict__ cand_tok, int k2) {
  __shared__ float red_f[16];
  __shared__ float ws[8];
  __shared__ int wi[8], wt[8];
  __shared__ float lse_m, lse_l;
  const int r = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto val = [&](int v) {
    float x = lp.p[0][(long long)r * ld + v];
    for (int m = 1; m < n_models; ++m) x += lp.p[m][(long long)r * ld + v];      // sum of next-token logits (beam_search.py:254)
    return x;
  };
  // pass A: max
  float mx = -INFINITY;
  for (int v = tid; v < V; v += 256) mx = fmaxf(mx, val(v));
  mx = warp_max(mx);
  if (lane == 0) red_f[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = red_f[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red_f[i]);
    lse_m = m;
  }
  __syncthreads();
  const float m = lse_m;
  // pass B: sum exp
  float sm_ = 0.f;
  for (int v = tid; v < V; v += 256) sm_ += expf(val(v) - m);
  sm_ = warp_sum(sm_);
  if (lane == 0) red_f[8 + warp] = sm_;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red_f[8 + i];
    lse_l = logf(t);
  }
  __syncthreads();
  const float lg = lse_l, bs = beam_scores[r];
  // pass C: thread-local top-K of score = ((x - max) - log sum) + beam_score   (log_softmax then + beam score, :260-265)
  float ts[K];
  int tt[K];
#pragma unroll
  for (int i = 0; i < K; ++i) { ts[i] = -INFINITY; tt[i] = 0x7fffffff; }
  for (int v = tid; v < V; v += 256) {
    const float sc = ((val(v) - m) - lg) + bs;
    if (cand_better(sc, v, ts[K - 1], tt[K - 1])) {
      ts[K - 1] = sc; tt[K - 1] = v;
#pragma unroll
      for (int i = K - 1; i > 0; --i) {
        if (cand_better(ts[i], tt[i], ts[i - 1], tt[i - 1])) {
          const float fs = ts[i]; ts[i] = ts[i - 1]; ts[i - 1] = fs;
          const int ft = tt[i]; tt[i] = tt[i - 1]; tt[i - 1] = ft;
        }
      }
    }
  }
  // merge: k2 rounds of block arg-best over the threads' list heads
  int head = 0;
  for (int round = 0; round < k2; ++round) {
    float cs = -INFINITY;
    int ct = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < K; ++i)
      if (i == head) { cs = ts[i]; ct = tt[i]; }
    int owner = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, cs, o);
      const int ot = __shfl_xor_sync(0xffffffffu, ct, o);
      const int oo = __shfl_xor_sync(0xffffffffu, owner, o);
      if (cand_better(os, ot, cs, ct)) { cs = os; ct = ot; owner = oo; }
    }
    if (lane == 0) { ws[warp] = cs; wt[warp] = ct; wi[warp] = owner; }
    __syncthreads();
    float bs_ = ws[0];
    int bt = wt[0], bo = wi[0];
    for (int i = 1; i < 8; ++i)
      if (cand_better(ws[i], wt[i], bs_, bt)) { bs_ = ws[i]; bt = wt[i]; bo = wi[i]; }
    if (tid == bo && head < K) ++head;
    if (tid == 0) {
      cand_score[(long long)r * k2 + round] = bs_;
      cand_tok[(long long)r * k2 + round] = bt;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------ beam: select
// State of one search (all device memory, owned by the caller):
//   ids [R, max_len] int64, beam_scores [R] f32, done [B] u8, next_tok [R] int64, parent [R] int32,
//   hyp_score [B, k] f64, hyp_len [B, k] i32, hyp_tok [B, k, max_len] int64, hyp_count [B] i32, hyp_worst [B] f64,
//   counters[4] i32: {t (tokens already consumed = position of the current token), n_done, final_len, unused}
// One warp per batch element; lane 0 runs the control flow of BeamSearchScorer.process (beam_search.py:297-304 / the legacy
// transformers BeamHypotheses), the other lanes help with the copies.  Scores of finished hypotheses are computed in double like
// the host implementation (Python floats).
__global__ void beam_select_kernel(const float* __restrict__ cand_score, const int* __restrict__ cand_tok, int k, int V, int B, int max_len,
                                   const long long* __restrict__ ids, float* __restrict__ beam_scores, uint8_t* __restrict__ done,
                                   long long* __restrict__ next_tok, int* __restrict__ parent, double* __restrict__ hyp_score,
                                   int* __restrict__ hyp_len, long long* __restrict__ hyp_tok, int* __restrict__ hyp_count,
                                   double* __restrict__ hyp_worst, int* __restrict__ counters, int eos, int pad, double length_penalty,
                                   int forced_last) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int t = counters[0];
  const int cur_len = t + 1;                    // tokens in ids[r] so far
  const int k2 = 2 * k;
  if (cur_len >= max_len) return;               // nothing to append (the host stops the loop; replays beyond are no-ops)
  if (k == 1) {
    // greedy (HF greedy search: argmax, finished rows emit pad)
    if (lane == 0) {
      const bool was_done = done[b] != 0;
      int tok = cand_tok[(long long)b * k2];
      if (forced_last >= 0 && cur_len + 1 == max_len) tok = forced_last;   // HF ForcedEOSTokenLogitsProcessor at the last position
      if (was_done) tok = pad;
      next_tok[b] = tok;
      parent[b] = b;
      beam_scores[b] = cand_score[(long long)b * k2];
      if (!was_done && tok == eos) {
        done[b] = 1;
        const int n = atomicAdd(&counters[1], 1) + 1;
        if (n == B) counters[2] = cur_len + 1;
      }
    }
    return;
  }
  if (done[b]) {
    if (lane < k) {
      next_tok[b * k + lane] = pad;
      parent[b * k + lane] = b * k + lane;
      beam_scores[b * k + lane] = 0.f;
    }
    return;
  }
  // ---- global top-2k of the k rows' candidate lists, ordered by (score desc, beam * V + token asc)
  extern __shared__ unsigned char sel_raw[];
  float* rs = reinterpret_cast<float*>(sel_raw) + (threadIdx.x >> 5) * 64;           // ranked scores  [<= 32]
  int* rflat_b = reinterpret_cast<int*>(rs + 32);                                    // ranked beam index
  int* rtok = reinterpret_cast<int*>(sel_raw + (blockDim.x >> 5) * 256) + (threadIdx.x >> 5) * 32;
  const int n_c = k * k2;                       // <= 8 * 16 = 128 -> 4 per lane
  float cs[4];
  long long cf[4];
  bool used[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = lane + 32 * i;
    used[i] = true;
    cs[i] = -INFINITY;
    cf[i] = 0x7fffffffffffffffLL;
    if (c < n_c) {
      const int beam = c / k2;
      cs[i] = cand_score[(long long)(b * k + beam) * k2 + (c % k2)];
      cf[i] = (long long)beam * V + cand_tok[(long long)(b * k + beam) * k2 + (c % k2)];
      used[i] = false;
    }
  }
  for (int rank = 0; rank < k2; ++rank) {
    float bs_ = -INFINITY;
    long long bf_ = 0x7fffffffffffffffLL;
    int bi = -1;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (!used[i] && (bi < 0 || cand_better(cs[i], cf[i], bs_, bf_))) { bs_ = cs[i]; bf_ = cf[i]; bi = i; }
    int owner = lane;
    float ws_ = bi < 0 ? -INFINITY : bs_;
    long long wf = bi < 0 ? 0x7fffffffffffffffLL : bf_;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, ws_, o);
      const long long of = __shfl_xor_sync(0xffffffffu, wf, o);
      const int oo = __shfl_xor_sync(0xffffffffu, owner, o);
      if (cand_better(os, of, ws_, wf)) { ws_ = os; wf = of; owner = oo; }
    }
    if (lane == owner && bi >= 0) used[bi] = true;
    if (lane == 0) {
      rs[rank] = ws_;
      rflat_b[rank] = (int)(wf / V);
      rtok[rank] = (int)(wf % V);
    }
  }
  __syncwarp();
  // ---- BeamSearchScorer.process
  const double len_pow = pow((double)cur_len, length_penalty);
  int j = 0;
  for (int rank = 0; rank < k2 && j < k; ++rank) {
    const int src = b * k + rflat_b[rank];
    const int tk = rtok[rank];
    const float sc = rs[rank];
    if (tk == eos) {
      if (rank >= k) continue;
      // hyps[b].add(ids[src], sc): score = sum_logprobs / len ** length_penalty
      const double score = (double)sc / len_pow;
      int cnt = hyp_count[b];
      const double worst = hyp_worst[b];
      if (cnt < k || score > worst) {
        // append at slot cnt (slots 0..k-1 live, slot k is scratch for the overflow case handled by shifting)
        int slot = cnt;
        if (cnt == k) {
          // remove the worst (lowest score, lowest index on ties), shifting the list left to keep insertion order
          int wi_ = 0;
          double ws2 = hyp_score[(long long)b * k];
          for (int i = 1; i < k; ++i)
            if (hyp_score[(long long)b * k + i] < ws2) { ws2 = hyp_score[(long long)b * k + i]; wi_ = i; }
          // (the new entry has index k and score > worst = the old minimum, so the entry removed by the host's
          //  `sorted((s, i))[0]` is always this old minimum)
          for (int i = wi_; i < k - 1; ++i) {
            if (lane == 0) {
              hyp_score[(long long)b * k + i] = hyp_score[(long long)b * k + i + 1];
              hyp_len[(long long)b * k + i] = hyp_len[(long long)b * k + i + 1];
            }
            for (int p = lane; p < max_len; p += 32)
              hyp_tok[((long long)b * k + i) * max_len + p] = hyp_tok[((long long)b * k + i + 1) * max_len + p];
            __syncwarp();
          }
          slot = k - 1;
          cnt = k - 1;
        }
        for (int p = lane; p < cur_len; p += 32) hyp_tok[((long long)b * k + slot) * max_len + p] = ids[(long long)src * max_len + p];
        __syncwarp();
        if (lane == 0) {
          hyp_score[(long long)b * k + slot] = score;
          hyp_len[(long long)b * k + slot] = cur_len;
          hyp_count[b] = cnt + 1;
        }
        __syncwarp();
        // worst = min over the live entries (host: min(score, worst) while filling, sorted(...)[1] after an overflow)
        if (lane == 0) {
          double w = hyp_score[(long long)b * k];
          for (int i = 1; i < cnt + 1; ++i) w = fmin(w, hyp_score[(long long)b * k + i]);
          hyp_worst[b] = w;
        }
        __syncwarp();
      }
    } else {
      if (lane == 0) {
        beam_scores[b * k + j] = sc;
        next_tok[b * k + j] = tk;
        parent[b * k + j] = src;
      }
      ++j;
    }
  }
  // is_done: len(hyps) >= k and worst >= best_sum_logprobs / cur_len ** length_penalty   (best = max over the ranked candidates)
  if (lane == 0) {
    const bool full = hyp_count[b] >= k;
    if (full && hyp_worst[b] >= (double)rs[0] / len_pow) {
      done[b] = 1;
      const int n = atomicAdd(&counters[1], 1) + 1;
      if (n == B) counters[2] = cur_len + 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------ sampling
// Logit filters of the reference's SCST rollouts (vilmedic/blocks/rl/SCST.py:139-153: generate(do_sample=True, top_k=...,
// bad_words_ids=[[pad], [bos]])): HF NoBadWordsLogitsProcessor sets the listed single-token ids to -inf, TopKLogitsWarper removes
// every score below the k-th largest (ties with the k-th are kept: `scores < topk(scores, k)[..., -1]`).  In place, one CTA per row;
// the k-th largest is found by a 32-step binary search over the order-preserving integer image of the fp32 values cached in smem.
__device__ __forceinline__ uint32_t f32_key(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
struct BadIds {
  int id[8];
};

template <typename TL>
__global__ void __launch_bounds__(256) logits_filter_kernel(TL* __restrict__ logits, long long ld, int V, BadIds bad, int n_bad, int top_k) {
  extern __shared__ uint32_t keys[];
  __shared__ int cnt_s[8];
  __shared__ uint32_t thr_s;
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  TL* row = logits + (long long)r * ld;
  const TL ninf = sizeof(TL) == 4 ? (TL)(-INFINITY) : (TL)__float2bfloat16(-INFINITY);
  for (int v = tid; v < V; v += 256) {
    float x = (float)row[v];
    for (int i = 0; i < n_bad; ++i)
      if (v == bad.id[i]) x = -INFINITY;
    keys[v] = f32_key(x);
  }
  __syncthreads();
  uint32_t thr = 0u;
  if (top_k > 0 && top_k < V) {
    // largest key value T such that count(key >= T) >= top_k  == the key of the k-th largest element
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = thr | (1u << bit);
      int c = 0;
      for (int v = tid; v < V; v += 256) c += keys[v] >= cand;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0) cnt_s[warp] = c;
      __syncthreads();
      if (tid == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += cnt_s[i];
        thr_s = (t >= top_k) ? cand : thr;
      }
      __syncthreads();
      thr = thr_s;
    }
  }
  const uint32_t ninf_key = f32_key(-INFINITY);
  for (int v = tid; v < V; v += 256) {
    const uint32_t kv = keys[v];
    if (kv == ninf_key || kv < thr) row[v] = ninf;
  }
}

// One token per row from softmax(logits / temperature) by the Gumbel-max trick: argmax_v (logit_v / T - log(-log u_v)), u from
// Philox(seed, offset + *t, row * V + v) — one pass, no normalisation needed for the draw; the log-probability of the drawn token
// (log-softmax of the same scaled logits) is written next to it.  Writes slot 0 of the row's candidate list (cand_* [R, 2]), so
// that vlm_beam_select's k = 1 path does the EOS / pad / finished bookkeeping.
__global__ void __launch_bounds__(256) sample_rows_kernel(const float* __restrict__ logits, long long ld, int V, float inv_temp,
                                                         unsigned long long seed, unsigned long long offset, const int* __restrict__ t_ptr,
                                                         float* __restrict__ cand_score, int* __restrict__ cand_tok) {
  __shared__ float red_f[8], red_g[8];
  __shared__ int red_i[8];
  __shared__ float bc[2];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = logits + (long long)r * ld;
  const Philox rng(seed);
  // counters[0] = step, counters[3] = per-rollout nonce (set by the host before the first step): both live in device memory so that
  // one captured graph serves every step of every rollout
  const unsigned long long off = offset + (t_ptr ? (unsigned long long)t_ptr[0] + ((unsigned long long)(unsigned)t_ptr[3] << 16) : 0ull);
  float mx = -INFINITY, best = -INFINITY;
  int best_v = 0x7fffffff;
  float best_x = -INFINITY;
  const int nq = (V + 3) / 4;
  for (int q4 = tid; q4 < nq; q4 += 256) {
    const uint4 u = rng((unsigned long long)r * (unsigned long long)nq + (unsigned long long)q4, off);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int v = q4 * 4 + j;
      if (v >= V) break;
      const float x = row[v] * inv_temp;
      mx = fmaxf(mx, x);
      if (x == -INFINITY) continue;
      const float uu = ((float)(w[j] >> 8) + 0.5f) * (1.0f / 16777216.0f);        // (0, 1)
      const float g = x - __logf(-__logf(uu));
      if (g > best || (g == best && v < best_v)) { best = g; best_v = v; best_x = x; }
    }
  }
  // block arg-max of the perturbed scores, block max of the scaled logits
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float og = __shfl_xor_sync(0xffffffffu, best, o), ox = __shfl_xor_sync(0xffffffffu, best_x, o);
    const int ov = __shfl_xor_sync(0xffffffffu, best_v, o);
    if (og > best || (og == best && ov < best_v)) { best = og; best_v = ov; best_x = ox; }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { red_g[warp] = best; red_i[warp] = best_v; red_f[warp] = mx; }
  __syncthreads();
  if (tid == 0) {
    float m = red_f[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red_f[i]);
    bc[0] = m;
  }
  __syncthreads();
  const float m = bc[0];
  float se = 0.f;
  for (int v = tid; v < V; v += 256) se += __expf(row[v] * inv_temp - m);
  se = warp_sum(se);
  __syncthreads();
  if (lane == 0) red_f[warp] = se;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f, bg = red_g[0];
    int bv = red_i[0];
    for (int i = 0; i < 8; ++i) t += red_f[i];
    for (int i = 1; i < 8; ++i)
      if (red_g[i] > bg || (red_g[i] == bg && red_i[i] < bv)) { bg = red_g[i]; bv = red_i[i]; }
    if (bv >= V) bv = 0;                                       // whole row filtered out: cannot happen with a sane filter
    cand_tok[(long long)r * 2] = bv;
    cand_tok[(long long)r * 2 + 1] = bv;
    const float lp = row[bv] * inv_temp - m - logf(t);
    cand_score[(long long)r * 2] = lp;
    cand_score[(long long)r * 2 + 1] = lp;
  }
}

// ids / row_map of every row follow the row's parent; the new token is appended; one CTA per row.
__global__ void beam_advance_kernel(const long long* __restrict__ ids_in, long long* __restrict__ ids_out, const int* __restrict__ map_in,
                                    int* __restrict__ map_out, const int* __restrict__ parent, const long long* __restrict__ next_tok,
                                    int max_len, const int* __restrict__ counters) {
  const int r = blockIdx.x;
  const int t = counters[0];
  const int cur_len = t + 1;
  const int p = parent[r];
  const bool append = cur_len < max_len;
  for (int i = threadIdx.x; i < max_len; i += blockDim.x) {
    long long v = ids_in[(long long)p * max_len + i];
    if (append && i == cur_len) v = next_tok[r];
    ids_out[(long long)r * max_len + i] = v;
    map_out[(long long)r * max_len + i] = map_in[(long long)p * max_len + i];
  }
}

__global__ void beam_commit_kernel(const long long* __restrict__ ids_tmp, long long* __restrict__ ids, const int* __restrict__ map_tmp,
                                   int* __restrict__ row_map, long long n, int* __restrict__ counters, int max_len) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    ids[i] = ids_tmp[i];
    row_map[i] = map_tmp[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && counters[0] + 1 < max_len) counters[0] += 1;
}

}  // namespace vlm

using namespace vlm;

extern "C" int vlm_embed_step(const long long* tok, const float* word, const float* pos, void* z, int R, int D, int V, const int* t_ptr,
                              int max_pos, int pos_shift, const float* tt_row, void* stream) {
  VLM_REQUIRE(tok && word && pos && z && t_ptr && R > 0 && D % 8 == 0 && V > 0 && max_pos > 0, "vlm_embed_step: bad args");
  const int n = R * (D / 8);
  embed_step_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(tok, word, pos, (bf16*)z, R, D, V, t_ptr, max_pos, pos_shift, tt_row);
  return check_launch("embed_step");
}

extern "C" int vlm_decode_attention(const void* q, long long ldq, const void* kv_new, long long ld_new, void* cache, long long cache_row_stride,
                                    int two_d, int* row_map, int map_ld, const int* t_ptr, int fixed_len, int row_div,
                                    const uint8_t* kmask, int kmask_ld, void* out, long long ldo, int R, int H, int DH, float scale,
                                    int max_len, void* stream) {
  VLM_REQUIRE(q && cache && out && R > 0 && H > 0, "vlm_decode_attention: null argument");
  VLM_REQUIRE(DH == 48 || DH == 64 || DH == 96, "vlm_decode_attention: head dim %d has no kernel (48, 64, 96)", DH);
  VLM_REQUIRE(two_d == 2 * H * DH, "vlm_decode_attention: cache rows must hold [K | V] of H * DH each");
  VLM_REQUIRE(ldq % 8 == 0 && ldo % 8 == 0 && cache_row_stride % 8 == 0 && (!kv_new || ld_new % 8 == 0), "vlm_decode_attention: 16-byte row pitches");
  VLM_REQUIRE(kv_new ? (row_map && t_ptr && map_ld >= max_len) : (fixed_len > 0 && row_div > 0), "vlm_decode_attention: self-attention needs row_map + t_ptr, cross-attention fixed_len + row_div");
  VLM_REQUIRE(max_len > 0 && max_len <= 8192, "vlm_decode_attention: max_len out of range");
  const int cap = kv_new ? max_len : fixed_len;
  const size_t smem = (size_t)(((cap + 3) & ~3) + 8 + 4 * DH) * sizeof(float);
  dim3 grid(R, H);
  cudaStream_t s = (cudaStream_t)stream;
#define VLM_DA(DH_)                                                                                                              \
  decode_attn_kernel<DH_><<<grid, DA_THREADS, smem, s>>>((const bf16*)q, ldq, (const bf16*)kv_new, ld_new, (bf16*)cache,         \
                                                        cache_row_stride, two_d, row_map, map_ld, t_ptr, fixed_len, row_div, kmask, \
                                                        kmask_ld, (bf16*)out, ldo, scale)
  if (DH == 48) VLM_DA(48);
  else if (DH == 64) VLM_DA(64);
  else VLM_DA(96);
#undef VLM_DA
  return check_launch("decode_attention");
}

extern "C" int vlm_beam_rows(const float* const* logits, int n_models, long long ld, int V, const float* beam_scores, float* cand_score,
                             int* cand_tok, int R, int k, void* stream) {
  VLM_REQUIRE(logits && n_models >= 1 && n_models <= 8 && beam_scores && cand_score && cand_tok && R > 0 && V > 0, "vlm_beam_rows: bad args");
  VLM_REQUIRE(k >= 1 && k <= 8 && 2 * k <= V, "vlm_beam_rows: beam width must be 1..8 and 2k <= V");
  LogitPtrs lp;
  for (int i = 0; i < 8; ++i) lp.p[i] = i < n_models ? logits[i] : nullptr;
  for (int i = 0; i < n_models; ++i) VLM_REQUIRE(lp.p[i], "vlm_beam_rows: null logits pointer");
  const int k2 = 2 * k;
  cudaStream_t s = (cudaStream_t)stream;
  if (k2 <= 2) beam_rows_kernel<2><<<R, 256, 0, s>>>(lp, n_models, ld, V, beam_scores, cand_score, cand_tok, k2);
  else if (k2 <= 4) beam_rows_kernel<4><<<R, 256, 0, s>>>(lp, n_models, ld, V, beam_scores, cand_score, cand_tok, k2);
  else if (k2 <= 8) beam_rows_kernel<8><<<R, 256, 0, s>>>(lp, n_models, ld, V, beam_scores, cand_score, cand_tok, k2);
  else beam_rows_kernel<16><<<R, 256, 0, s>>>(lp, n_models, ld, V, beam_scores, cand_score, cand_tok, k2);
  return check_launch("beam_rows");
}

extern "C" int vlm_beam_select(const float* cand_score, const int* cand_tok, int k, int V, int B, int max_len, const long long* ids,
                               float* beam_scores, uint8_t* done, long long* next_tok, int* parent, double* hyp_score, int* hyp_len,
                               long long* hyp_tok, int* hyp_count, double* hyp_worst, int* counters, int eos, int pad,
                               double length_penalty, int forced_last_token, void* stream) {
  VLM_REQUIRE(cand_score && cand_tok && ids && beam_scores && done && next_tok && parent && counters, "vlm_beam_select: null argument");
  VLM_REQUIRE(k >= 1 && k <= 8 && B > 0 && max_len > 1, "vlm_beam_select: bad sizes");
  VLM_REQUIRE(k == 1 || (hyp_score && hyp_len && hyp_tok && hyp_count && hyp_worst), "vlm_beam_select: beam search needs the hypothesis buffers");
  VLM_REQUIRE(k == 1 || forced_last_token < 0, "vlm_beam_select: a forced last token is only supported for greedy / sampling (k = 1)");
  const int warps = 4;
  const size_t smem = (size_t)warps * (256 + 128);
  beam_select_kernel<<<(B + warps - 1) / warps, warps * 32, smem, (cudaStream_t)stream>>>(cand_score, cand_tok, k, V, B, max_len, ids, beam_scores,
                                                                                         done, next_tok, parent, hyp_score, hyp_len, hyp_tok,
                                                                                         hyp_count, hyp_worst, counters, eos, pad, length_penalty, forced_last_token);
  return check_launch("beam_select");
}

extern "C" int vlm_beam_advance(long long* ids, long long* ids_tmp, int* row_map, int* map_tmp, const int* parent, const long long* next_tok,
                                int R, int max_len, int* counters, void* stream) {
  VLM_REQUIRE(ids && ids_tmp && row_map && map_tmp && parent && next_tok && counters && R > 0 && max_len > 1, "vlm_beam_advance: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  beam_advance_kernel<<<R, 128, 0, s>>>(ids, ids_tmp, row_map, map_tmp, parent, next_tok, max_len, counters);
  if (check_launch("beam_advance")) return -1;
  const long long n = (long long)R * max_len;
  beam_commit_kernel<<<(int)((n + 255) / 256 < 296 ? (n + 255) / 256 : 296), 256, 0, s>>>(ids_tmp, ids, map_tmp, row_map, n, counters, max_len);
  return check_launch("beam_commit");
}

extern "C" int vlm_logits_filter(void* logits, int logits_fp32, long long ld, int R, int V, const int* bad_ids, int n_bad, int top_k,
                                 void* stream) {
  VLM_REQUIRE(logits && R > 0 && V > 0 && ld >= V, "vlm_logits_filter: bad args");
  VLM_REQUIRE(n_bad >= 0 && n_bad <= 8 && (n_bad == 0 || bad_ids), "vlm_logits_filter: at most 8 single-token bad ids (host array)");
  VLM_REQUIRE(top_k >= 0, "vlm_logits_filter: top_k must be >= 0 (0 = off)");
  VLM_REQUIRE((size_t)V * 4 <= 200 * 1024, "vlm_logits_filter: V=%d too large for the row cache", V);
  BadIds bad;
  for (int i = 0; i < 8; ++i) bad.id[i] = i < n_bad ? bad_ids[i] : -1;
  const size_t smem = (size_t)V * 4;
  cudaStream_t s = (cudaStream_t)stream;
  if (logits_fp32) {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(logits_filter_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set = true; }
    logits_filter_kernel<float><<<R, 256, smem, s>>>((float*)logits, ld, V, bad, n_bad, top_k);
  } else {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(logits_filter_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set = true; }
    logits_filter_kernel<bf16><<<R, 256, smem, s>>>((bf16*)logits, ld, V, bad, n_bad, top_k);
  }
  return check_launch("logits_filter");
}

extern "C" int vlm_sample_rows(const float* logits, long long ld, int V, float temperature, unsigned long long seed,
                               unsigned long long offset, const int* t_ptr, float* cand_score, int* cand_tok, int R, void* stream) {
  VLM_REQUIRE(logits && cand_score && cand_tok && R > 0 && V > 0 && ld >= V && temperature > 0.f, "vlm_sample_rows: bad args");
  sample_rows_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(logits, ld, V, 1.f / temperature, seed, offset, t_ptr, cand_score, cand_tok);
  return check_launch("sample_rows");
}
