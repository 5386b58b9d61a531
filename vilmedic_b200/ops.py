"""Thin, autograd-free Python wrappers over the C ABI (one per entry point of include/vlm_b200.h).

These only validate dtypes/devices, allocate outputs with torch (device memory + stream plumbing) and forward raw
pointers.  All arithmetic happens in libvlmb200.so.
"""
import ctypes

import torch

from . import _lib
from ._lib import c_float, c_int, c_ll, c_u64, c_void_p, check, ptr, stream_ptr

ACT_NONE, ACT_GELU, ACT_GELU_GRAD = 0, 1, 2

# Number of OUR kernels launched through this module (bench.py reports it as `gpu_launches`).
LAUNCHES = [0]
# (attention backward: one kernel on the tcgen05 path since the delta pre-pass moved into it; the mma.sync fallback launches more —
#  counted as one, i.e. never over-claimed)
_KERNELS_PER_CALL = {"vlm_adamw_step": 2, "vlm_beam_advance": 2}
# Optional recording of the GEMM launches (bench.py roofline pass): list of (M, N, K, batch, algorithmic HBM bytes =
# operands read once + outputs written once, replay closure, operands kept alive)
GEMM_TIMING = None
# Optional device uint64 added to every dropout offset (see vlm_rng_advance): set by GraphedTrainStep.
RNG_COUNTER = [None]

_orig_check = check


def check(rc, what):  # noqa: F811
    _orig_check(rc, what)
    LAUNCHES[0] += _KERNELS_PER_CALL.get(what, 1)


def _req(cond, msg):
    if not cond:
        raise ValueError(msg)


# ------------------------------------------------------------------------------------------------ side stream / background mode
class background:
    """with ops.background(): helper kernels (optimizer update, column sums) launch as small co-resident CTAs (vlm_set_background)."""

    def __enter__(self):
        self.prev = _L().vlm_set_background(c_int(1))

    def __exit__(self, *exc):
        _L().vlm_set_background(c_int(self.prev))


class SideWork:
    """Kernels of the backward pass that nothing downstream waits for until the layer's gradients are announced run on a SIDE
    stream, under the dependent chain of the main stream:
      * SIDE      — the bias-gradient column sums (73 launches / 0.87 ms of the RRG step when serialised), in background mode
                    (small CTAs that fit next to a resident persistent GEMM CTA).          VLM_SIDE_COLSUM=0 turns it off;
      * SIDE_GEMM — the weight-gradient GEMMs (a third of the backward GEMM time): a persistent 148-CTA GEMM whose tile count is
                    1.3-1.7 waves leaves a quarter of the SMs idle in its last wave, and the dgrad -> LayerNorm -> attention chain
                    of the main stream is full of such tails; CTAs of the wgrad kernel fill them (measured: -0.2 ms of 22.4 ms on one
                    B200, tools/jobs/r2w.sh).                                                   VLM_SIDE_WGRAD=0 turns it off.
    `run(fn, *tensors)` forks from the current stream and keeps the tensors alive; `join()` makes the current stream wait for
    everything issued so far (nn.notify_grad_ready calls it before a span's gradients are declared final).  Works inside
    CUDA-graph capture (fork / join become graph edges)."""

    def __init__(self, env, default, bg):
        import os
        self.enabled = os.environ.get(env, default) != "0"
        self.bg = bg
        self.streams = {}
        self.pending = []
        self.dirty = False

    def _stream(self, dev):
        st = self.streams.get(dev)
        if st is None:
            st = self.streams[dev] = torch.cuda.Stream(device=dev)     # default (= lowest) priority
        return st

    def run(self, fn, *keep):
        if not self.enabled:
            fn()
            return
        cur = torch.cuda.current_stream()
        st = self._stream(cur.device)
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            if self.bg:
                with background():
                    fn()
            else:
                fn()
        self.pending.append(keep)
        self.dirty = True

    def join(self):
        if self.dirty:
            cur = torch.cuda.current_stream()
            cur.wait_stream(self._stream(cur.device))
            self.pending.clear()
            self.dirty = False


SIDE = SideWork("VLM_SIDE_COLSUM", "1", True)
SIDE_GEMM = SideWork("VLM_SIDE_WGRAD", "1", False)


def _is_bf16_cuda(t):
    return t.is_cuda and t.dtype == torch.bfloat16


def gemm(a, b, *, a_mn_major=False, b_mn_major=False, out=None, out_dtype=torch.bfloat16, bias=None, residual=None,
         act=ACT_NONE, aux_in=None, aux_out=None, alpha=1.0, alpha_t=None, accumulate=False, p_drop=0.0, seed=0, offset=0,
         force_bn=0, max_ctas=0):
    """C[M,N] = epi(alpha * A' B'^T).

    a: [M,K] (K-major) or [K,M] (a_mn_major);  b: [N,K] (K-major, nn.Linear weight layout) or [K,N] (b_mn_major).
    3-D operands ([batch, rows, cols]) run the strided-batched path.
    """
    _req(_is_bf16_cuda(a) and _is_bf16_cuda(b), "gemm operands must be CUDA bf16")
    batched = a.dim() == 3
    _req(a.dim() == b.dim() and a.dim() in (2, 3), "gemm operands must both be 2-D or 3-D")
    _req(a.stride(-1) == 1 and b.stride(-1) == 1, "gemm operands must have a contiguous last dim")
    if a_mn_major:
        K, M = a.shape[-2], a.shape[-1]
    else:
        M, K = a.shape[-2], a.shape[-1]
    if b_mn_major:
        Kb, N = b.shape[-2], b.shape[-1]
    else:
        N, Kb = b.shape[-2], b.shape[-1]
    _req(K == Kb, "gemm inner dims differ: %d vs %d" % (K, Kb))
    batch = a.shape[0] if batched else 1
    if batched:
        _req(b.shape[0] == batch, "batch mismatch")
    if out is None:
        al = 4 if out_dtype == torch.float32 else 8          # rows must stay 16-byte aligned
        Np = (N + al - 1) // al * al
        shape = (batch, M, Np) if batched else (M, Np)
        out = torch.empty(shape, device=a.device, dtype=out_dtype)
        if Np != N:
            out = out[..., :N]
    _req(out.is_cuda and out.dtype in (torch.bfloat16, torch.float32) and out.stride(-1) == 1, "bad gemm out")
    _req(tuple(out.shape[-2:]) == (M, N), "gemm out shape %s != (%d,%d)" % (tuple(out.shape), M, N))
    c_fp32 = out.dtype == torch.float32
    if bias is not None:
        _req(bias.is_cuda and bias.dtype == torch.float32 and bias.numel() >= N and bias.is_contiguous(), "bias must be fp32 [N]")
    ldr = 0
    res_bs = 0
    if residual is not None:
        _req(residual.dtype == out.dtype and residual.stride(-1) == 1, "residual dtype must match out")
        ldr = residual.stride(-2)
        if batched:
            res_bs = residual.stride(0) if residual.dim() == 3 else 0
    ld_aux = 0
    aux_bs = 0
    for t in (aux_in, aux_out):
        if t is not None:
            _req(_is_bf16_cuda(t) and t.stride(-1) == 1, "aux must be CUDA bf16")
            ld_aux = t.stride(-2)
            if batched:
                aux_bs = t.stride(0)
    args = (
        ptr(a), c_ll(a.stride(-2)), c_int(int(a_mn_major)),
        ptr(b), c_ll(b.stride(-2)), c_int(int(b_mn_major)),
        ptr(out), c_ll(out.stride(-2)), c_int(int(c_fp32)),
        c_int(M), c_int(N), c_int(K),
        ptr(bias), ptr(residual), c_ll(ldr), c_int(act), ptr(aux_in), ptr(aux_out), c_ll(ld_aux),
        c_float(alpha), ptr(alpha_t), c_int(int(accumulate)), c_int(batch),
        c_ll(a.stride(0) if batched else 0), c_ll(b.stride(0) if batched else 0),
        c_ll(out.stride(0) if batched else 0), c_ll(aux_bs), c_ll(res_bs),
        c_float(p_drop), c_u64(seed), c_u64(offset), ptr(RNG_COUNTER[0]), c_int(force_bn), c_int(max_ctas))
    rc = _lib.lib().vlm_gemm_bf16(*args, stream_ptr())
    check(rc, "vlm_gemm_bf16")
    if GEMM_TIMING is not None:
        # bench.py roofline pass: keep the exact call (and its operands alive) so that it can be replayed under CUDA events
        esz = out.element_size()
        nbytes = batch * (2 * (M * K + N * K) + M * N * esz * (1 + int(residual is not None) + int(bool(accumulate)))
                          + 2 * M * N * (int(aux_in is not None) + int(aux_out is not None)))
        keep = (a, b, out, bias, residual, aux_in, aux_out, alpha_t)
        GEMM_TIMING.append((M, N, K, batch, nbytes, lambda: _lib.lib().vlm_gemm_bf16(*args, stream_ptr()), keep))
    return out


def _L():
    return _lib.lib()


def layernorm_fwd(x, gamma, beta, eps, save_stats=True):
    """x [M,D] bf16|fp32 -> (y bf16, mean, rstd)."""
    _req(x.is_cuda and x.dim() == 2 and x.is_contiguous(), "layernorm input must be contiguous [M,D]")
    M, D = x.shape
    y = torch.empty((M, D), device=x.device, dtype=torch.bfloat16)
    mean = torch.empty(M, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if save_stats else None
    check(_L().vlm_layernorm_fwd(ptr(x), c_int(int(x.dtype == torch.float32)), ptr(gamma), ptr(beta), ptr(y), ptr(mean),
                                 ptr(rstd), c_int(M), c_int(D), c_float(eps), stream_ptr()), "vlm_layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, dres=None, drop=None, colsum=None):
    """Returns dx (dtype of x) — or (dx, dropout(dx)) when drop=(p, seed, offset); ACCUMULATES into dgamma / dbeta and,
    if given, the column sums of the (dropped) dx into `colsum` (fp32 [D])."""
    M, D = x.shape
    _req(dy.is_contiguous() and dy.dtype == torch.bfloat16 and x.is_contiguous(), "layernorm_bwd: bad inputs")
    dx = torch.empty_like(x)
    if dres is not None:
        _req(dres.dtype == x.dtype and dres.is_contiguous(), "dres dtype must match x")
    dxd = torch.empty_like(x) if drop is not None else None
    p, seed, off = drop if drop is not None else (0.0, 0, 0)
    check(_L().vlm_layernorm_bwd(ptr(dy), ptr(x), c_int(int(x.dtype == torch.float32)), ptr(mean), ptr(rstd), ptr(gamma),
                                 ptr(dres), ptr(dx), ptr(dgamma), ptr(dbeta), c_int(M), c_int(D), ptr(dxd), c_float(p),
                                 c_u64(seed), c_u64(off), ptr(RNG_COUNTER[0]), ptr(colsum), stream_ptr()),
          "vlm_layernorm_bwd")
    if drop is not None:
        return dx, dxd
    return dx


def _bh_strides(t, H, DH):
    """t is a [B, T, >=H*DH] view (last dim contiguous) -> (batch stride, row stride)."""
    _req(t.dim() == 3 and t.stride(2) == 1 and t.dtype == torch.bfloat16, "attention operand must be [B,T,H*DH] bf16 view")
    return t.stride(0), t.stride(1)


def attention_fwd(q, k, v, H, DH, *, kmask=None, causal=False, scale=None, p_drop=0.0, seed=0, offset=0, force_tc=False):
    """q [B,Tq,H*DH], k/v [B,Sk,H*DH] (possibly strided views of packed projections) -> (o [B,Tq,H*DH], lse [B,H,Tq])."""
    B, Tq = q.shape[0], q.shape[1]
    Sk = k.shape[1]
    scale = (1.0 / DH ** 0.5) if scale is None else scale
    o = torch.empty((B, Tq, H * DH), device=q.device, dtype=torch.bfloat16)
    lse = torch.empty((B, H, Tq), device=q.device, dtype=torch.float32)
    qs, ks, vs = _bh_strides(q, H, DH), _bh_strides(k, H, DH), _bh_strides(v, H, DH)
    if kmask is not None:
        _req(kmask.dtype == torch.uint8 and kmask.is_contiguous() and tuple(kmask.shape) == (B, Sk), "kmask must be uint8 [B,Sk]")
    fn = _L().vlm_attention_fwd_tc if force_tc else _L().vlm_attention_fwd
    check(fn(ptr(q), c_ll(qs[0]), c_ll(qs[1]), ptr(k), c_ll(ks[0]), c_ll(ks[1]), ptr(v), c_ll(vs[0]),
                                 c_ll(vs[1]), ptr(o), c_ll(o.stride(0)), c_ll(o.stride(1)), ptr(lse), ptr(kmask),
                                 c_int(B), c_int(H), c_int(Tq), c_int(Sk), c_int(DH), c_int(int(causal)), c_float(scale),
                                 c_float(p_drop), c_u64(seed), c_u64(offset), ptr(RNG_COUNTER[0]), stream_ptr()), "vlm_attention_fwd")
    return o, lse


def attention_bwd(q, k, v, o, do, lse, dq, dk, dv, H, DH, *, kmask=None, causal=False, scale=None, p_drop=0.0, seed=0,
                  offset=0, force_tc=False):
    """Writes dq/dk/dv (pre-allocated bf16 views with the q/k/v addressing scheme).  force_tc: call the tcgen05 kernel
    explicitly (tests); otherwise libvlmb200 picks (VLM_ATTN_TC=1 routes supported shapes to tcgen05)."""
    B, Tq = q.shape[0], q.shape[1]
    Sk = k.shape[1]
    scale = (1.0 / DH ** 0.5) if scale is None else scale
    s = [_bh_strides(t, H, DH) for t in (q, k, v, o, do, dq, dk, dv)]
    delta = torch.empty((B, H, Tq), device=q.device, dtype=torch.float32)
    if force_tc:
        check(_L().vlm_attention_bwd_tc(ptr(q), c_ll(s[0][0]), c_ll(s[0][1]), ptr(k), c_ll(s[1][0]), c_ll(s[1][1]), ptr(v),
                                        c_ll(s[2][0]), c_ll(s[2][1]), ptr(o), c_ll(s[3][0]), c_ll(s[3][1]), ptr(do),
                                        c_ll(s[4][0]), c_ll(s[4][1]), ptr(lse), ptr(delta), ptr(dq), c_ll(s[5][0]), c_ll(s[5][1]),
                                        ptr(dk), c_ll(s[6][0]), c_ll(s[6][1]), ptr(dv), c_ll(s[7][0]), c_ll(s[7][1]),
                                        ptr(kmask), c_int(B), c_int(H), c_int(Tq), c_int(Sk), c_int(DH), c_int(int(causal)),
                                        c_float(scale), c_float(p_drop), c_u64(seed), c_u64(offset), ptr(RNG_COUNTER[0]),
                                        stream_ptr()), "vlm_attention_bwd_tc")
        return
    check(_L().vlm_attention_bwd(ptr(q), c_ll(s[0][0]), c_ll(s[0][1]), ptr(k), c_ll(s[1][0]), c_ll(s[1][1]), ptr(v),
                                 c_ll(s[2][0]), c_ll(s[2][1]), ptr(o), c_ll(s[3][0]), c_ll(s[3][1]), ptr(do),
                                 c_ll(s[4][0]), c_ll(s[4][1]), ptr(lse), ptr(delta), ptr(dq), c_ll(s[5][0]),
                                 c_ll(s[5][1]), ptr(dk), c_ll(s[6][0]), c_ll(s[6][1]), ptr(dv), c_ll(s[7][0]),
                                 c_ll(s[7][1]), ptr(kmask), c_int(B), c_int(H), c_int(Tq), c_int(Sk), c_int(DH),
                                 c_int(int(causal)), c_float(scale), c_float(p_drop), c_u64(seed), c_u64(offset),
                                 ptr(RNG_COUNTER[0]), stream_ptr()), "vlm_attention_bwd")


def softmax_ce(logits, ids, V, *, shift_T=0, smoothing=0.0, grad_scale=1.0, dlogits=None, want_lse=False, row_weight=None):
    """logits [R, ld>=V] (bf16|fp32).  Returns (loss_rows fp32 [R], lse_rows|None).  dlogits may alias logits."""
    _req(logits.dim() == 2 and logits.stride(1) == 1, "logits must be [R, ld]")
    R = logits.shape[0]
    loss_rows = torch.empty(R, device=logits.device, dtype=torch.float32)
    lse_rows = torch.empty(R, device=logits.device, dtype=torch.float32) if want_lse else None
    _req(ids.dtype == torch.int64 and ids.is_contiguous() and ids.numel() == R, "ids must be int64 [R]")
    if dlogits is not None:
        _req(dlogits.dtype == logits.dtype and dlogits.stride(1) == 1, "dlogits dtype must match logits")
    check(_L().vlm_softmax_ce(ptr(logits), c_int(int(logits.dtype == torch.float32)), c_ll(logits.stride(0)), ptr(ids),
                              c_int(shift_T), c_int(R), c_int(V), c_float(smoothing), c_float(grad_scale), ptr(dlogits),
                              c_ll(dlogits.stride(0) if dlogits is not None else 0), ptr(loss_rows), ptr(lse_rows), ptr(row_weight),
                              stream_ptr()), "vlm_softmax_ce")
    return loss_rows, lse_rows


def cast_bf16(src, dst=None):
    _req(src.is_cuda and src.dtype == torch.float32 and src.is_contiguous(), "cast_bf16: need contiguous fp32")
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    check(_L().vlm_cast_f32_to_bf16(ptr(src), ptr(dst), c_ll(src.numel()), stream_ptr()), "vlm_cast_f32_to_bf16")
    return dst


def patchify(images, P, n_prefix=1):
    _req(images.is_cuda and images.dtype == torch.float32 and images.is_contiguous() and images.dim() == 4, "patchify: fp32 NCHW")
    B, C, Hh, W = images.shape
    S = n_prefix + (Hh // P) * (W // P)
    out = torch.empty((B, S, C * P * P), device=images.device, dtype=torch.bfloat16)
    check(_L().vlm_patchify(ptr(images), ptr(out), c_int(B), c_int(C), c_int(Hh), c_int(W), c_int(P), c_int(n_prefix), stream_ptr()),
          "vlm_patchify")
    return out


def vit_cls_pos(x, tok, pos, row=0):
    B, S, D = x.shape
    check(_L().vlm_vit_cls_pos(ptr(x), c_int(int(x.dtype == torch.float32)), ptr(tok), ptr(pos), c_int(B), c_int(S), c_int(D), c_int(row),
                               stream_ptr()), "vlm_vit_cls_pos")


def vit_embed_bwd(dx, dpos, dcls, dbias, ddist=None):
    B, S, D = dx.shape
    _req(dx.is_contiguous(), "vit_embed_bwd: dx must be contiguous")
    check(_L().vlm_vit_embed_bwd(ptr(dx), c_int(int(dx.dtype == torch.float32)), ptr(dpos), ptr(dcls), ptr(ddist), ptr(dbias), c_int(B),
                                 c_int(S), c_int(D), c_int(2 if ddist is not None else 1), stream_ptr()), "vlm_vit_embed_bwd")


def colsum(x, out, scale_t=None):
    """out[n] += scale * sum_m x[m,n] (x bf16 [M,N] view with contiguous last dim)."""
    _req(x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.bfloat16, "colsum: bf16 [M,N]")
    check(_L().vlm_colsum_bf16(ptr(x), c_ll(x.stride(0)), ptr(out), c_int(x.shape[0]), c_int(x.shape[1]), ptr(scale_t),
                               stream_ptr()), "vlm_colsum_bf16")


def features_mask(feats):
    """feats bf16 [B,S,D] contiguous -> uint8 [B,S]."""
    _req(feats.is_contiguous() and feats.dtype == torch.bfloat16, "features_mask: contiguous bf16")
    B, S, D = feats.shape
    mask = torch.empty((B, S), device=feats.device, dtype=torch.uint8)
    check(_L().vlm_features_mask(ptr(feats), ptr(mask), c_int(B * S), c_int(D), stream_ptr()), "vlm_features_mask")
    return mask


def embed_fwd(ids, word, pos, T, pos_offset=0, pos_ids=None, tt_row=None):
    _req(ids.dtype == torch.int64 and ids.is_contiguous(), "ids must be contiguous int64")
    R = ids.numel()
    V, D = word.shape
    z = torch.empty((R, D), device=word.device, dtype=torch.bfloat16)
    if pos_ids is not None:
        _req(pos_ids.dtype == torch.int32 and pos_ids.is_contiguous() and pos_ids.numel() == R, "embed_fwd: pos_ids int32 [R]")
    check(_L().vlm_embed_fwd(ptr(ids), ptr(word), ptr(pos), ptr(z), c_int(R), c_int(T), c_int(D), c_int(V), c_int(pos_offset),
                             ptr(pos_ids), ptr(tt_row), stream_ptr()), "vlm_embed_fwd")
    return z


def embed_bwd(ids, dz, dword, dpos, T, V, pos_offset=0, padding_idx=-1, pos_ids=None):
    R, D = dz.shape
    _req(dz.is_contiguous() and dz.dtype == torch.bfloat16, "embed_bwd: dz contiguous bf16")
    check(_L().vlm_embed_bwd(ptr(ids), ptr(dz), ptr(dword), ptr(dpos), c_int(R), c_int(T), c_int(D), c_int(V),
                             c_int(pos_offset), c_int(padding_idx), ptr(pos_ids), stream_ptr()), "vlm_embed_bwd")


def dropout(x, p, seed, offset, out=None):
    _req(x.is_contiguous() and x.dtype == torch.bfloat16 and x.numel() % 8 == 0, "dropout: contiguous bf16, numel % 8 == 0")
    if out is None:
        out = torch.empty_like(x)
    check(_L().vlm_dropout_bf16(ptr(x), ptr(out), c_ll(x.numel()), c_float(p), c_u64(seed), c_u64(offset), ptr(RNG_COUNTER[0]),
                                stream_ptr()), "vlm_dropout_bf16")
    return out


def mask_rows(x, mask, rows_per_mask):
    _req(x.dim() == 2 and x.is_contiguous() and x.dtype == torch.bfloat16, "mask_rows: contiguous bf16 [R,D]")
    y = torch.empty_like(x)
    check(_L().vlm_mask_rows_bf16(ptr(x), ptr(y), ptr(mask), c_int(x.shape[0]), c_int(x.shape[1]), c_int(rows_per_mask),
                                  stream_ptr()), "vlm_mask_rows_bf16")
    return y


def act_fwd(x, kind):
    _req(x.is_cuda and x.dtype == torch.float32 and x.is_contiguous(), "act_fwd: contiguous fp32")
    y = torch.empty_like(x)
    check(_L().vlm_act_fwd_f32(ptr(x), ptr(y), c_ll(x.numel()), c_int(kind), stream_ptr()), "vlm_act_fwd_f32")
    return y


def act_bwd(dy, y, kind):
    dx = torch.empty_like(y)
    check(_L().vlm_act_bwd_f32(ptr(dy), ptr(y), ptr(dx), c_ll(y.numel()), c_int(kind), stream_ptr()), "vlm_act_bwd_f32")
    return dx


def sum_scale(x, scale):
    out = torch.empty((), device=x.device, dtype=torch.float32)
    check(_L().vlm_sum_scale_f32(ptr(x), c_int(x.numel()), c_float(scale), ptr(out), stream_ptr()), "vlm_sum_scale_f32")
    return out


def sumsq(g, out):
    if g.dtype == torch.bfloat16:
        check(_L().vlm_sumsq_bf16(ptr(g), c_ll(g.numel()), ptr(out), stream_ptr()), "vlm_sumsq_bf16")
    else:
        check(_L().vlm_sumsq_f32(ptr(g), c_ll(g.numel()), ptr(out), stream_ptr()), "vlm_sumsq_f32")


def adamw_step(p, g, m, v, p_bf16, *, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, step_t=None,
               increment_step=True, lr_scale_t=None, grad_scale=1.0, gnorm_sq_t=None, max_norm=0.0, zero_grad=True):
    n = p.numel()
    check(_L().vlm_adamw_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_bf16), c_ll(n), c_float(lr), c_float(betas[0]),
                              c_float(betas[1]), c_float(eps), c_float(weight_decay), ptr(step_t), c_int(int(increment_step)),
                              ptr(lr_scale_t), c_float(grad_scale), ptr(gnorm_sq_t), c_float(max_norm),
                              c_int(int(zero_grad)), stream_ptr()), "vlm_adamw_step")


def rownorm_split(x, normalize, eps=1e-8, want_a=True, want_b=True, want_h=True):
    _req(x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2, "rownorm_split: contiguous fp32 [N,D]")
    N, D = x.shape
    mk = lambda w: torch.empty((N, w), device=x.device, dtype=torch.bfloat16)
    xa = mk(3 * D) if want_a else None
    xb = mk(3 * D) if want_b else None
    xh = mk(D) if want_h else None
    inv = torch.empty(N, device=x.device, dtype=torch.float32)
    check(_L().vlm_rownorm_split(ptr(x), ptr(xa), ptr(xb), ptr(xh), ptr(inv), c_int(N), c_int(D), c_int(int(normalize)),
                                 c_float(eps), stream_ptr()), "vlm_rownorm_split")
    return xa, xb, xh, inv


def rownorm_bwd(x, inv, dxh, normalize):
    dx = torch.empty_like(x)
    check(_L().vlm_rownorm_bwd(ptr(x), ptr(inv), ptr(dxh), ptr(dx), c_int(x.shape[0]), c_int(x.shape[1]), c_int(int(normalize)),
                               stream_ptr()), "vlm_rownorm_bwd")
    return dx


def sym_lse(S, scale):
    N = S.shape[0]
    outs = [torch.empty(N, device=S.device, dtype=torch.float32) for _ in range(4)]
    check(_L().vlm_sym_lse(ptr(S), c_int(N), c_ll(S.stride(0)), c_float(scale), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]),
                           ptr(outs[3]), stream_ptr()), "vlm_sym_lse")
    return outs  # lse_row, lse_col, loss_row, loss_col


def sym_lse_bwd(S, scale, lse_row, lse_col, w_row, w_col, g_t):
    N = S.shape[0]
    ldd = (N + 7) // 8 * 8
    dS = torch.empty((N, ldd), device=S.device, dtype=torch.bfloat16)
    check(_L().vlm_sym_lse_bwd(ptr(S), c_int(N), c_ll(S.stride(0)), c_float(scale), ptr(lse_row), ptr(lse_col), c_float(w_row),
                               c_float(w_col), ptr(g_t), ptr(dS), c_ll(ldd), stream_ptr()), "vlm_sym_lse_bwd")
    return dS[:, :N]


# ---- GLoRIA local loss pieces (csrc/gloria.cu) ---------------------------------------------------------------------------
def transpose_cast(x, out_dtype, c_out=None, row_limit=None, want_lo=False):
    """x fp32 [B, R, C] -> [B, c_out or C, R] (bf16 or fp32); rows >= C or >= row_limit[b] are zero.
    want_lo (bf16 only): returns (hi, lo) with x ~= hi + lo."""
    _req(x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.stride(-1) == 1, "transpose_cast: x must be CUDA fp32 [B,R,C]")
    B, R, C = x.shape
    Co = C if c_out is None else c_out
    out = torch.empty((B, Co, R), device=x.device, dtype=out_dtype)
    lo = torch.empty_like(out) if (want_lo and out_dtype == torch.bfloat16) else None
    check(_L().vlm_transpose_cast(ptr(x), ptr(out), ptr(lo), c_int(int(out_dtype == torch.bfloat16)), c_int(B), c_int(R), c_int(C),
                                  c_ll(x.stride(1)), c_ll(x.stride(0)), c_ll(R), c_ll(Co * R), c_int(Co), ptr(row_limit),
                                  stream_ptr()), "vlm_transpose_cast")
    return (out, lo) if want_lo else out


def gloria_word_softmax(A, cap_lens, NB, L):
    P1 = torch.empty_like(A)
    check(_L().vlm_gloria_word_softmax(ptr(A), ptr(P1), ptr(cap_lens), c_int(A.shape[0]), c_int(NB), c_int(L), c_ll(A.stride(0)),
                                       stream_ptr()), "vlm_gloria_word_softmax")
    return P1


def gloria_region_softmax(P1, cap_lens, NI, S, NB, L, temp1):
    P2 = torch.empty_like(P1)
    P2h = torch.empty(P1.shape, device=P1.device, dtype=torch.bfloat16)
    P2l = torch.empty_like(P2h)
    check(_L().vlm_gloria_region_softmax(ptr(P1), ptr(P2), ptr(P2h), ptr(P2l), ptr(cap_lens), c_int(NI), c_int(S), c_int(NB), c_int(L),
                                         c_float(temp1), stream_ptr()), "vlm_gloria_region_softmax")
    return P2, P2h, P2l


def gloria_cos(WC, Q, cap_lens, NB, L, eps=1e-8):
    NI, NL, D = WC.shape
    cosv = torch.empty((NI, NL), device=WC.device, dtype=torch.float32)
    wnorm = torch.empty_like(cosv)
    qnorm = torch.empty(NL, device=WC.device, dtype=torch.float32)
    check(_L().vlm_gloria_cos(ptr(WC), ptr(Q), ptr(cap_lens), ptr(cosv), ptr(wnorm), ptr(qnorm), c_int(NI), c_int(NB), c_int(L),
                              c_int(D), c_float(eps), stream_ptr()), "vlm_gloria_cos")
    return cosv, wnorm, qnorm


def gloria_sims(cosv, cap_lens, NB, L, temp2, temp3):
    sims = torch.empty((NB, NB), device=cosv.device, dtype=torch.float32)
    check(_L().vlm_gloria_sims(ptr(cosv), ptr(cap_lens), ptr(sims), c_int(NB), c_int(L), c_float(temp2), c_float(temp3),
                               stream_ptr()), "vlm_gloria_sims")
    return sims


def gloria_cos_bwd(WC, Q, cap_lens, cosv, wnorm, qnorm, sims, lse_row, lse_col, g0, g1, NB, L, temp2, temp3, eps=1e-8):
    NI, NL, D = WC.shape
    dWC = torch.empty((NI, NL, D), device=WC.device, dtype=torch.bfloat16)
    dQ = torch.empty((NL, D), device=WC.device, dtype=torch.float32)
    check(_L().vlm_gloria_cos_bwd(ptr(WC), ptr(Q), ptr(cap_lens), ptr(cosv), ptr(wnorm), ptr(qnorm), ptr(sims), ptr(lse_row),
                                  ptr(lse_col), ptr(g0), ptr(g1), ptr(dWC), ptr(dQ), c_int(NB), c_int(L), c_int(D), c_float(temp2),
                                  c_float(temp3), c_float(eps), stream_ptr()), "vlm_gloria_cos_bwd")
    return dWC, dQ


def gloria_region_softmax_bwd(P2, G, NI, S, temp1):
    check(_L().vlm_gloria_region_softmax_bwd(ptr(P2), ptr(G), c_int(NI), c_int(S), c_ll(P2.shape[-1]), c_float(temp1),
                                             stream_ptr()), "vlm_gloria_region_softmax_bwd")
    return G


def gloria_word_softmax_bwd(P1, G, cap_lens, NB, L):
    dA = torch.empty(P1.shape, device=P1.device, dtype=torch.bfloat16)
    check(_L().vlm_gloria_word_softmax_bwd(ptr(P1), ptr(G), ptr(dA), ptr(cap_lens), c_int(P1.shape[0]), c_int(NB), c_int(L),
                                           c_ll(P1.stride(0)), stream_ptr()), "vlm_gloria_word_softmax_bwd")
    return dA


def image_crop_flip_normalize(images_u8, top, left, flip, crop, mean, std):
    """images_u8 uint8 [B,H,W,3] (CUDA) + per-image crop origin / flip flag -> fp32 [B,3,crop,crop] (ImageDataset.py:97-104)."""
    _req(images_u8.is_cuda and images_u8.dtype == torch.uint8 and images_u8.dim() == 4 and images_u8.shape[3] == 3 and
         images_u8.is_contiguous(), "image_crop_flip_normalize: contiguous CUDA uint8 [B,H,W,3]")
    B, H, W, _ = images_u8.shape
    _req(top.dtype == torch.int32 and left.dtype == torch.int32 and flip.dtype == torch.uint8 and top.numel() == B and
         left.numel() == B and flip.numel() == B and top.is_cuda and left.is_cuda and flip.is_cuda,
         "image_crop_flip_normalize: top/left int32 [B], flip uint8 [B] on the device")
    out = torch.empty((B, 3, crop, crop), device=images_u8.device, dtype=torch.float32)
    m = (ctypes.c_float * 3)(*[float(x) for x in mean])
    sd = (ctypes.c_float * 3)(*[float(x) for x in std])
    check(_L().vlm_image_crop_flip_normalize(ptr(images_u8), ptr(out), ptr(top), ptr(left), ptr(flip), c_int(B), c_int(H), c_int(W),
                                             c_int(crop), m, sd, stream_ptr()), "vlm_image_crop_flip_normalize")
    return out


# ---- CNN backbone pieces (csrc/conv.cu) -----------------------------------------------------------------------------------
def conv_out_size(H, k, stride, pad):
    return (H + 2 * pad - k) // stride + 1


def conv_weight_pack(w, Kp):
    """w fp32 OIHW -> bf16 [Cout, Kp] with columns ordered (kh, kw, ci), zero padded to Kp."""
    Cout, Cin, KH, KW = w.shape
    _req(w.is_cuda and w.dtype == torch.float32 and w.is_contiguous(), "conv_weight_pack: contiguous CUDA fp32 OIHW")
    wm = torch.empty((Cout, Kp), device=w.device, dtype=torch.bfloat16)
    check(_L().vlm_conv_weight_pack(ptr(w), ptr(wm), c_int(Cout), c_int(Cin), c_int(KH), c_int(KW), c_int(Kp), stream_ptr()),
          "vlm_conv_weight_pack")
    return wm


def conv_wgrad_unpack(dwm, gw):
    """gw (fp32 OIHW view of the gradient arena) += dwm fp32 [Cout, Kp]."""
    Cout, Cin, KH, KW = gw.shape
    _req(dwm.dtype == torch.float32 and dwm.is_contiguous() and gw.dtype == torch.float32 and gw.is_contiguous(), "conv_wgrad_unpack: fp32")
    check(_L().vlm_conv_wgrad_unpack(ptr(dwm), ptr(gw), c_int(Cout), c_int(Cin), c_int(KH), c_int(KW), c_int(dwm.shape[1]), stream_ptr()),
          "vlm_conv_wgrad_unpack")


def im2col_nhwc(x, B, H, W, C, KH, KW, stride, pad):
    """x bf16 [B*H*W, C] -> col bf16 [B*Ho*Wo, KH*KW*C]."""
    _req(_is_bf16_cuda(x) and x.is_contiguous() and x.numel() == B * H * W * C, "im2col_nhwc: contiguous bf16 [B*H*W, C]")
    Ho, Wo = conv_out_size(H, KH, stride, pad), conv_out_size(W, KW, stride, pad)
    col = torch.empty((B * Ho * Wo, KH * KW * C), device=x.device, dtype=torch.bfloat16)
    check(_L().vlm_im2col_nhwc(ptr(x), ptr(col), c_int(B), c_int(H), c_int(W), c_int(C), c_int(KH), c_int(KW), c_int(stride), c_int(pad),
                               stream_ptr()), "vlm_im2col_nhwc")
    return col


def im2col_nchw_f32(img, KH, KW, stride, pad, Kp):
    """img fp32 [B,Cin,H,W] -> col bf16 [B*Ho*Wo, Kp] (columns (kh, kw, ci), zero padded)."""
    _req(img.is_cuda and img.dtype == torch.float32 and img.is_contiguous() and img.dim() == 4, "im2col_nchw_f32: contiguous fp32 NCHW")
    B, Cin, H, W = img.shape
    Ho, Wo = conv_out_size(H, KH, stride, pad), conv_out_size(W, KW, stride, pad)
    col = torch.empty((B * Ho * Wo, Kp), device=img.device, dtype=torch.bfloat16)
    check(_L().vlm_im2col_nchw_f32(ptr(img), ptr(col), c_int(B), c_int(Cin), c_int(H), c_int(W), c_int(KH), c_int(KW), c_int(stride),
                                   c_int(pad), c_int(Kp), stream_ptr()), "vlm_im2col_nchw_f32")
    return col


def col2im_nhwc(dcol, B, H, W, C, KH, KW, stride, pad, add=None):
    """dcol bf16 [B*Ho*Wo, KH*KW*C] -> dx bf16 [B*H*W, C] (+ add)."""
    _req(_is_bf16_cuda(dcol) and dcol.is_contiguous(), "col2im_nhwc: contiguous bf16")
    if add is not None:
        _req(_is_bf16_cuda(add) and add.is_contiguous() and add.numel() == B * H * W * C, "col2im_nhwc: add must be bf16 [B*H*W, C]")
    dx = torch.empty((B * H * W, C), device=dcol.device, dtype=torch.bfloat16)
    check(_L().vlm_col2im_nhwc(ptr(dcol), ptr(add), ptr(dx), c_int(B), c_int(H), c_int(W), c_int(C), c_int(KH), c_int(KW), c_int(stride),
                               c_int(pad), stream_ptr()), "vlm_col2im_nhwc")
    return dx


def bn_train_fwd(x, gamma, beta, running_mean, running_var, num_batches, eps, momentum, relu, res=None):
    """x bf16 [M,C] -> (y bf16, mean, rstd) with torch.nn.BatchNorm2d training semantics (running buffers updated in place)."""
    M, C = x.shape
    _req(_is_bf16_cuda(x) and x.is_contiguous(), "bn_train_fwd: contiguous bf16 [M,C]")
    y = torch.empty_like(x)
    mean, rstd, scale, shift = (torch.empty(C, device=x.device, dtype=torch.float32) for _ in range(4))
    ws = torch.empty(2 * C, device=x.device, dtype=torch.float32)
    check(_L().vlm_bn_train_fwd(ptr(x), ptr(res), ptr(y), ptr(gamma), ptr(beta), ptr(mean), ptr(rstd), ptr(scale), ptr(shift), ptr(ws),
                                ptr(running_mean), ptr(running_var), ptr(num_batches), c_int(M), c_int(C), c_float(eps), c_float(momentum),
                                c_int(int(relu)), stream_ptr()), "vlm_bn_train_fwd")
    return y, mean, rstd


def bn_eval_fwd(x, gamma, beta, running_mean, running_var, eps, relu, res=None):
    M, C = x.shape
    _req(_is_bf16_cuda(x) and x.is_contiguous(), "bn_eval_fwd: contiguous bf16 [M,C]")
    y = torch.empty_like(x)
    scale, shift = (torch.empty(C, device=x.device, dtype=torch.float32) for _ in range(2))
    check(_L().vlm_bn_eval_fwd(ptr(x), ptr(res), ptr(y), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(scale), ptr(shift),
                               c_int(M), c_int(C), c_float(eps), c_int(int(relu)), stream_ptr()), "vlm_bn_eval_fwd")
    return y


def bn_train_bwd(dy, y, x, mean, rstd, gamma, dgamma, dbeta, relu, want_dres):
    """-> (dx bf16, dres bf16|None); ACCUMULATES into dgamma / dbeta."""
    M, C = x.shape
    _req(_is_bf16_cuda(dy) and dy.is_contiguous() and tuple(dy.shape) == (M, C), "bn_train_bwd: dy must be contiguous bf16 [M,C]")
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dres else None
    ws = torch.empty(2 * C, device=x.device, dtype=torch.float32)
    check(_L().vlm_bn_train_bwd(ptr(dy), ptr(y), ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(dgamma), ptr(dbeta), ptr(ws), ptr(dx),
                                ptr(dres), c_int(M), c_int(C), c_int(int(relu)), stream_ptr()), "vlm_bn_train_bwd")
    return dx, dres


def maxpool3x3s2_fwd(x, B, H, W, C):
    Ho, Wo = conv_out_size(H, 3, 2, 1), conv_out_size(W, 3, 2, 1)
    _req(_is_bf16_cuda(x) and x.is_contiguous() and x.numel() == B * H * W * C, "maxpool3x3s2_fwd: contiguous bf16 [B*H*W, C]")
    y = torch.empty((B * Ho * Wo, C), device=x.device, dtype=torch.bfloat16)
    idx = torch.empty((B * Ho * Wo, C), device=x.device, dtype=torch.uint8)
    check(_L().vlm_maxpool3x3s2_fwd(ptr(x), ptr(y), ptr(idx), c_int(B), c_int(H), c_int(W), c_int(C), stream_ptr()), "vlm_maxpool3x3s2_fwd")
    return y, idx


def maxpool3x3s2_bwd(dy, idx, B, H, W, C):
    _req(_is_bf16_cuda(dy) and dy.is_contiguous(), "maxpool3x3s2_bwd: contiguous bf16")
    dx = torch.empty((B * H * W, C), device=dy.device, dtype=torch.bfloat16)
    check(_L().vlm_maxpool3x3s2_bwd(ptr(dy), ptr(idx), ptr(dx), c_int(B), c_int(H), c_int(W), c_int(C), stream_ptr()), "vlm_maxpool3x3s2_bwd")
    return dx


def avgpool_fwd(x, B, HW, C):
    _req(_is_bf16_cuda(x) and x.is_contiguous() and x.numel() == B * HW * C, "avgpool_fwd: contiguous bf16 [B*HW, C]")
    y = torch.empty((B, C), device=x.device, dtype=torch.bfloat16)
    check(_L().vlm_avgpool_fwd(ptr(x), ptr(y), c_int(B), c_int(HW), c_int(C), stream_ptr()), "vlm_avgpool_fwd")
    return y


def avgpool_bwd(dy, B, HW, C):
    _req(_is_bf16_cuda(dy) and dy.is_contiguous() and dy.numel() == B * C, "avgpool_bwd: contiguous bf16 [B, C]")
    dx = torch.empty((B * HW, C), device=dy.device, dtype=torch.bfloat16)
    check(_L().vlm_avgpool_bwd(ptr(dy), ptr(dx), c_int(B), c_int(HW), c_int(C), stream_ptr()), "vlm_avgpool_bwd")
    return dx


def rng_advance(counter, delta):
    check(_L().vlm_rng_advance(ptr(counter), c_u64(delta), stream_ptr()), "vlm_rng_advance")


# ---- incremental decoding / device-side beam search (csrc/decode.cu) ------------------------------------------------------------
def embed_step(tok, word, pos, t_ptr, pos_shift=0, tt_row=None):
    """tok int64 [R] (device), position read from the device counter t_ptr -> bf16 [R, D]."""
    _req(tok.dtype == torch.int64 and tok.is_contiguous() and t_ptr.dtype == torch.int32, "embed_step: tok int64, t_ptr int32")
    R = tok.numel()
    V, D = word.shape
    z = torch.empty((R, D), device=word.device, dtype=torch.bfloat16)
    check(_L().vlm_embed_step(ptr(tok), ptr(word), ptr(pos), ptr(z), c_int(R), c_int(D), c_int(V), ptr(t_ptr), c_int(pos.shape[0]),
                              c_int(pos_shift), ptr(tt_row), stream_ptr()), "vlm_embed_step")
    return z


def decode_attention(q, H, DH, cache, *, kv_new=None, row_map=None, t_ptr=None, fixed_len=0, row_div=1, kmask=None, max_len=0):
    """q bf16 [R, >=H*DH] view; cache bf16 [rows, L, 2*H*DH].  Self-attention: kv_new [R, 2*H*DH] view + row_map int32 [R, max_len] +
    t_ptr; cross-attention: fixed_len / row_div (+ kmask uint8 [rows, fixed_len]).  -> bf16 [R, H*DH]."""
    R = q.shape[0]
    D = H * DH
    _req(q.dtype == torch.bfloat16 and q.stride(1) == 1 and cache.dtype == torch.bfloat16 and cache.is_contiguous() and cache.shape[2] == 2 * D,
         "decode_attention: q bf16 rows, cache contiguous bf16 [rows, L, 2D]")
    out = torch.empty((R, D), device=q.device, dtype=torch.bfloat16)
    if kv_new is not None:
        _req(kv_new.dtype == torch.bfloat16 and kv_new.stride(1) == 1 and row_map.dtype == torch.int32 and row_map.is_contiguous(),
             "decode_attention: kv_new bf16 rows, row_map contiguous int32")
    if kmask is not None:
        _req(kmask.dtype == torch.uint8 and kmask.is_contiguous(), "decode_attention: kmask uint8 contiguous")
    check(_L().vlm_decode_attention(ptr(q), c_ll(q.stride(0)), ptr(kv_new), c_ll(kv_new.stride(0) if kv_new is not None else 0), ptr(cache),
                                    c_ll(cache.stride(0)), c_int(2 * D), ptr(row_map), c_int(row_map.shape[1] if row_map is not None else 0),
                                    ptr(t_ptr), c_int(fixed_len), c_int(row_div), ptr(kmask), c_int(kmask.shape[1] if kmask is not None else 0),
                                    ptr(out), c_ll(out.stride(0)), c_int(R), c_int(H), c_int(DH), c_float(1.0 / DH ** 0.5),
                                    c_int(max_len if kv_new is not None else max(fixed_len, 1)), stream_ptr()), "vlm_decode_attention")
    return out


def beam_rows(logits_list, V, beam_scores, cand_score, cand_tok, k):
    """logits_list: fp32 [R, ld] tensors (one per ensemble member, same pitch)."""
    n = len(logits_list)
    ld = logits_list[0].stride(0)
    for l in logits_list:
        _req(l.dtype == torch.float32 and l.stride(1) == 1 and l.stride(0) == ld, "beam_rows: fp32 logits with equal row pitch")
    arr = (ctypes.c_void_p * n)(*[l.data_ptr() for l in logits_list])
    check(_L().vlm_beam_rows(arr, c_int(n), c_ll(ld), c_int(V), ptr(beam_scores), ptr(cand_score), ptr(cand_tok),
                             c_int(beam_scores.numel()), c_int(k), stream_ptr()), "vlm_beam_rows")


def beam_select(st, k, V, B, max_len, eos, pad, length_penalty, forced_last=-1):
    check(_L().vlm_beam_select(ptr(st["cand_score"]), ptr(st["cand_tok"]), c_int(k), c_int(V), c_int(B), c_int(max_len), ptr(st["ids"]),
                               ptr(st["beam_scores"]), ptr(st["done"]), ptr(st["next_tok"]), ptr(st["parent"]), ptr(st.get("hyp_score")),
                               ptr(st.get("hyp_len")), ptr(st.get("hyp_tok")), ptr(st.get("hyp_count")), ptr(st.get("hyp_worst")),
                               ptr(st["counters"]), c_int(eos), c_int(pad), ctypes.c_double(length_penalty), c_int(forced_last), stream_ptr()), "vlm_beam_select")


def beam_advance(st, R, max_len):
    check(_L().vlm_beam_advance(ptr(st["ids"]), ptr(st["ids_tmp"]), ptr(st["row_map"]), ptr(st["map_tmp"]), ptr(st["parent"]),
                                ptr(st["next_tok"]), c_int(R), c_int(max_len), ptr(st["counters"]), stream_ptr()), "vlm_beam_advance")


def logits_filter(logits, V, bad_ids=(), top_k=0):
    """In place: bad single-token ids -> -inf, then keep the top_k scores of every row (0 = off).  logits bf16 | fp32 [R, ld]."""
    _req(logits.is_cuda and logits.dim() == 2 and logits.stride(1) == 1 and logits.dtype in (torch.float32, torch.bfloat16), "logits_filter: [R, ld] rows")
    bad = [int(b) for b in bad_ids]
    arr = (ctypes.c_int * max(len(bad), 1))(*(bad or [0]))
    check(_L().vlm_logits_filter(ptr(logits), c_int(int(logits.dtype == torch.float32)), c_ll(logits.stride(0)), c_int(logits.shape[0]),
                                 c_int(V), arr, c_int(len(bad)), c_int(int(top_k or 0)), stream_ptr()), "vlm_logits_filter")
    return logits


def sample_rows(logits, V, cand_score, cand_tok, seed, offset, t_ptr, temperature=1.0):
    _req(logits.dtype == torch.float32 and logits.stride(1) == 1, "sample_rows: fp32 logits rows")
    check(_L().vlm_sample_rows(ptr(logits), c_ll(logits.stride(0)), c_int(V), c_float(temperature), c_u64(seed), c_u64(offset), ptr(t_ptr),
                               ptr(cand_score), ptr(cand_tok), c_int(logits.shape[0]), stream_ptr()), "vlm_sample_rows")


def image_resample_u8(x, bounds, coefs, out_h, out_w, axis):
    """One Pillow-compatible resampling pass over uint8 [B,H,W,3] (axis 0: columns -> out_w, axis 1: rows -> out_h)."""
    _req(x.is_cuda and x.dtype == torch.uint8 and x.dim() == 4 and x.shape[3] == 3 and x.is_contiguous(), "image_resample_u8: uint8 [B,H,W,3]")
    B, H, W, _ = x.shape
    out = torch.empty((B, out_h, out_w, 3), device=x.device, dtype=torch.uint8)
    check(_L().vlm_image_resample_u8(ptr(x), ptr(out), ptr(bounds), ptr(coefs), c_int(coefs.shape[1]), c_int(B), c_int(H), c_int(W),
                                     c_int(out_h), c_int(out_w), c_int(axis), stream_ptr()), "vlm_image_resample_u8")
    return out
