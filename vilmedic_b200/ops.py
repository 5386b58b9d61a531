"""Thin, autograd-free Python wrappers over the C ABI (one per entry point of include/vlm_b200.h).

These only validate dtypes/devices, allocate outputs with torch (device memory + stream plumbing) and forward raw
pointers.  All arithmetic happens in libvlmb200.so.
"""
import ctypes

import torch

from . import _lib
from ._lib import c_float, c_int, c_ll, c_u64, c_void_p, check, ptr, stream_ptr

ACT_NONE, ACT_GELU, ACT_GELU_GRAD = 0, 1, 2


def _req(cond, msg):
    if not cond:
        raise ValueError(msg)


def _is_bf16_cuda(t):
    return t.is_cuda and t.dtype == torch.bfloat16


def gemm(a, b, *, a_mn_major=False, b_mn_major=False, out=None, out_dtype=torch.bfloat16, bias=None, residual=None,
         act=ACT_NONE, aux_in=None, aux_out=None, alpha=1.0, accumulate=False, force_bn=0, max_ctas=0):
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
        shape = (batch, M, N) if batched else (M, N)
        out = torch.empty(shape, device=a.device, dtype=out_dtype)
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
    rc = _lib.lib().vlm_gemm_bf16(
        ptr(a), c_ll(a.stride(-2)), c_int(int(a_mn_major)),
        ptr(b), c_ll(b.stride(-2)), c_int(int(b_mn_major)),
        ptr(out), c_ll(out.stride(-2)), c_int(int(c_fp32)),
        c_int(M), c_int(N), c_int(K),
        ptr(bias), ptr(residual), c_ll(ldr), c_int(act), ptr(aux_in), ptr(aux_out), c_ll(ld_aux),
        c_float(alpha), c_int(int(accumulate)), c_int(batch),
        c_ll(a.stride(0) if batched else 0), c_ll(b.stride(0) if batched else 0),
        c_ll(out.stride(0) if batched else 0), c_ll(aux_bs), c_ll(res_bs),
        c_int(force_bn), c_int(max_ctas), stream_ptr())
    check(rc, "vlm_gemm_bf16")
    return out
