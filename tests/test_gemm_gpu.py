"""Parity of the tcgen05 GEMM (through the C ABI) against fp32 torch.matmul on the same bf16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, a_mn, b_mn):
    A = a.float().transpose(-1, -2) if a_mn else a.float()
    B = b.float() if b_mn else b.float().transpose(-1, -2)
    return A @ B


def _rand(shape, dev, seed):
    """bf16 randn whose row pitch is padded to a multiple of 8 elements (TMA needs 16-byte row strides)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = torch.randn(shape, generator=g).to(dev).to(torch.bfloat16)
    pad = (-shape[-1]) % 8
    if pad:
        buf = torch.zeros(shape[:-1] + (shape[-1] + pad,), device=dev, dtype=torch.bfloat16)
        buf[..., :shape[-1]] = t
        t = buf[..., :shape[-1]]
    return t


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 256), (394, 768, 768), (1000, 2304, 768),
                                   (197, 3072, 768), (640, 30522 // 8 * 8, 128), (8, 16, 24), (300, 520, 200)])
@pytest.mark.parametrize("bn", [0, 64, 128, 192, 256])
def test_gemm_plain(cuda_dev, a_mn, b_mn, M, N, K, bn):
    from vilmedic_b200 import ops
    if bn != 0 and (M, N, K) not in [(256, 256, 256), (394, 768, 768), (300, 520, 200)]:
        pytest.skip("forced tile sizes only on a subset")
    a = _rand((K, M) if a_mn else (M, K), cuda_dev, 1)
    b = _rand((K, N) if b_mn else (N, K), cuda_dev, 2)
    ref = _ref(a, b, a_mn, b_mn)
    out = ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, out_dtype=torch.float32, force_bn=bn)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, "max err %g (scale %g)" % (err, scale)
    out16 = ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, force_bn=bn)
    torch.cuda.synchronize()
    assert (out16.float() - ref).abs().max().item() <= 1e-2 * scale + 1e-2


def test_gemm_epilogue(cuda_dev):
    from vilmedic_b200 import ops
    M, N, K = 394, 3072, 768
    a = _rand((M, K), cuda_dev, 3)
    w = _rand((N, K), cuda_dev, 4) * 0.05
    bias = torch.randn(N, device=cuda_dev)
    res = _rand((M, N), cuda_dev, 5)
    pre_ref = a.float() @ w.float().t() + bias
    # bias + GELU; the epilogue stashes GELU'(pre-activation) (bf16) for the backward multiply
    xr = pre_ref.clone().requires_grad_(True)
    torch.nn.functional.gelu(xr).sum().backward()
    dgelu_ref = xr.grad
    pre = torch.empty(M, N, device=cuda_dev, dtype=torch.bfloat16)
    h = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, aux_out=pre)
    torch.cuda.synchronize()
    assert (pre.float() - dgelu_ref).abs().max().item() < 2 ** -8 * 1.2 + 1e-3        # |GELU'| <= 1.13: one bf16 ulp
    g_ref = torch.nn.functional.gelu(pre_ref)
    assert ((h.float() - g_ref).abs() - 2 ** -8 * g_ref.abs()).max().item() < 2e-3        # one bf16 ulp of the value + erf approximation
    h_nostash = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU)
    assert torch.equal(h, h_nostash)
    # bias + residual, bf16 and fp32 outputs
    y = ops.gemm(a, w, bias=bias, residual=res)
    assert (y.float() - (pre_ref + res.float())).abs().max().item() < 5e-2
    y32 = ops.gemm(a, w, bias=bias, residual=res.float(), out_dtype=torch.float32)
    assert (y32 - (pre_ref + res.float())).abs().max().item() < 2e-3
    # accumulate
    acc = torch.ones(M, N, device=cuda_dev)
    ops.gemm(a, w, out=acc, accumulate=True, alpha=0.5)
    assert (acc - (1 + 0.5 * (pre_ref - bias))).abs().max().item() < 2e-3
    # GELU' epilogue (backward of the FFN activation)
    g = _rand((M, K), cuda_dev, 6)
    dpre = ops.gemm(g, w, act=ops.ACT_GELU_GRAD, aux_in=pre, out_dtype=torch.float32)
    up = g.float() @ w.float().t()
    assert (dpre - up * pre.float()).abs().max().item() < 2e-3 * up.abs().max().item()       # exact multiply by the stash
    assert (dpre - up * dgelu_ref).abs().max().item() < 2 ** -7 * up.abs().max().item() + 1e-3   # vs the fp32 derivative
    dpre16 = ops.gemm(g, w, act=ops.ACT_GELU_GRAD, aux_in=pre)                                # staged fast path (bf16 C)
    assert (dpre16.float() - up * pre.float()).abs().max().item() < 1e-2 * up.abs().max().item()
    torch.cuda.synchronize()


def test_gemm_batched(cuda_dev):
    from vilmedic_b200 import ops
    Bt, M, N, K = 5, 197, 768, 768
    a = _rand((Bt, M, K), cuda_dev, 7)
    w = _rand((1, N, K), cuda_dev, 8).expand(Bt, N, K)
    # shared weight through a zero batch stride is expressed by passing stride 0
    w0 = w[0].contiguous()
    outs = torch.empty(Bt, M, N, device=cuda_dev, dtype=torch.float32)
    wb = w0.unsqueeze(0).expand(Bt, N, K)
    ops.gemm(a, wb, out=outs)
    torch.cuda.synchronize()
    ref = a.float() @ w0.float().t()
    assert (outs - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3
    # true batched (attention-like): per-batch B, MN-major B
    k = _rand((Bt, 200, 64), cuda_dev, 9)
    q = _rand((Bt, 128, 64), cuda_dev, 10)
    s = ops.gemm(q, k, out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref = q.float() @ k.float().transpose(1, 2)
    assert (s - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3
    v = _rand((Bt, 200, 64), cuda_dev, 11)
    p = _rand((Bt, 128, 200), cuda_dev, 12)
    o = ops.gemm(p, v, b_mn_major=True, out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref = p.float() @ v.float()
    assert (o - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 768, 768), (1000, 2304, 768), (12608, 768, 768), (640, 30522, 128),
                                   (300, 520, 200)])
@pytest.mark.parametrize("bn", [1128, 1256])
def test_gemm_2cta(cuda_dev, a_mn, b_mn, M, N, K, bn):
    """cta_group::2 kernel (CTA pair per 256 x BN tile), forced through force_bn = 1000 + BN."""
    from vilmedic_b200 import ops
    if (M, N, K) == (12608, 768, 768) and (a_mn or b_mn):
        pytest.skip("big shape only in the forward layout")
    a = _rand((K, M) if a_mn else (M, K), cuda_dev, 1)
    b = _rand((K, N) if b_mn else (N, K), cuda_dev, 2)
    ref = _ref(a, b, a_mn, b_mn)
    out = ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, out_dtype=torch.float32, force_bn=bn)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, "max err %g (scale %g)" % (err, scale)


def test_gemm_2cta_epilogue_and_repeat(cuda_dev):
    from vilmedic_b200 import ops
    M, N, K = 1576, 3072, 768
    a = _rand((M, K), cuda_dev, 3)
    w = _rand((N, K), cuda_dev, 4) * 0.05
    bias = torch.randn(N, device=cuda_dev)
    pre_ref = a.float() @ w.float().t() + bias
    pre = torch.empty(M, N, device=cuda_dev, dtype=torch.bfloat16)
    for _ in range(3):   # back-to-back launches exercise TMEM alloc/dealloc + barrier re-init across kernels
        h = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, aux_out=pre, force_bn=1256)
    torch.cuda.synchronize()
    xr = pre_ref.clone().requires_grad_(True)
    torch.nn.functional.gelu(xr).sum().backward()
    assert (pre.float() - xr.grad).abs().max().item() < 2 ** -8 * 1.2 + 1e-3
    gref = torch.nn.functional.gelu(pre_ref)
    assert ((h.float() - gref).abs() - 2 ** -8 * gref.abs()).max().item() < 2e-3
    acc = torch.ones(N, K, device=cuda_dev)
    g = _rand((M, N), cuda_dev, 5)
    ops.gemm(g, a, a_mn_major=True, b_mn_major=True, out=acc, accumulate=True, force_bn=1128)   # wgrad layout
    torch.cuda.synchronize()
    ref = 1 + g.float().t() @ a.float()
    assert (acc - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3


def test_gemm_split_k_wgrad(cuda_dev):
    """Weight-gradient shape (few output tiles, K = tokens): the split-K path accumulates with atomics into fp32 C."""
    from vilmedic_b200 import ops
    for (M, N, K) in [(768, 768, 8192), (2304, 768, 12608), (768, 3072, 8192), (330, 768, 1024)]:
        dy = _rand((K, M), cuda_dev, 1)        # [tokens, out]  -> A' MN-major
        x = _rand((K, N), cuda_dev, 2)         # [tokens, in]   -> B' MN-major
        acc = torch.full((M, N), 0.5, device=cuda_dev)
        ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out=acc, accumulate=True)
        torch.cuda.synchronize()
        ref = 0.5 + dy.float().t() @ x.float()
        assert (acc - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-2, (M, N, K)


def test_gemm_unaligned_n_output(cuda_dev):
    from vilmedic_b200 import ops
    a, w = _rand((37, 768), cuda_dev, 1), _rand((330, 768), cuda_dev, 2)
    bias = torch.randn(332, device=cuda_dev)
    out = ops.gemm(a, w, bias=bias, out_dtype=torch.float32)
    assert out.shape == (37, 330)
    ref = a.float() @ w.float().t() + bias[:330]
    assert (out - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3
