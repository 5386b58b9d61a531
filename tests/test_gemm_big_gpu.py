"""tcgen05 GEMM parity AT THE BENCHMARKED SHAPES (BASELINE configs[1]: B=64 -> M = 12608 ViT tokens / 8192 decoder tokens,
N in {768, 2304, 3072, 30522}, K in {768, 3072}) for every epilogue mode the training step uses.  At these sizes every persistent
CTA processes 2.7 .. 13 tiles, i.e. the code that small tests never reach runs: TMEM accumulator ping-pong, double-buffered
staging tiles, the one-span-ahead row-input requests that cross tile boundaries, split-K and 2-CTA scheduling.
Reference: fp32 torch.matmul (TF32 off) on the same bf16-rounded operands, on the GPU.

Error model (stated, then asserted element-wise): the kernel accumulates in fp32 and rounds ONCE to bf16, so
|out - ref| <= 2^-8 |ref| (one bf16 ulp; round-to-nearest gives half of it) + a small absolute term for fp32 summation order.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

ULP = 2.0 ** -8


@pytest.fixture(autouse=True)
def _no_tf32():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = old


def _rand(shape, seed, scale=1.0, pitch=None):
    """bf16 randn [rows, cols] on the GPU with a row pitch that is a multiple of 8 elements."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    rows, cols = shape
    pitch = pitch or (cols + 7) // 8 * 8
    buf = torch.zeros(rows, pitch, device="cuda", dtype=torch.bfloat16)
    buf[:, :cols] = (torch.randn(rows, cols, device="cuda", generator=g) * scale).to(torch.bfloat16)
    return buf[:, :cols]


def _check(out, ref, what, abs_tol, rel=ULP):
    out = out.float()
    bound = rel * ref.abs() + abs_tol
    excess = ((out - ref).abs() / bound).max().item()
    assert excess <= 1.0, "%s: max |err| / bound = %.3f (abs_tol %.2e)" % (what, excess, abs_tol)
    return excess


SHAPES_FWD = [  # M, N, K
    (12608, 768, 768), (12608, 2304, 768), (12608, 3072, 768), (12608, 768, 3072),
    (8192, 768, 768), (8192, 3072, 768), (8192, 768, 3072), (8192, 30522, 768),
]


@pytest.mark.parametrize("M,N,K", SHAPES_FWD)
@pytest.mark.parametrize("bn", [0, 1256])
def test_forward_bias(cuda_dev, M, N, K, bn):
    from vilmedic_b200 import ops
    a, w = _rand((M, K), 1), _rand((N, K), 2, 0.05)
    bias = torch.randn((N + 7) // 8 * 8, device="cuda")
    ref = a.float() @ w.float().t() + bias[:N]
    out = ops.gemm(a, w, bias=bias, force_bn=bn)
    torch.cuda.synchronize()
    assert out.shape == (M, N)
    _check(out, ref, "bias", 2e-5 * K ** 0.5)


@pytest.mark.parametrize("M,N,K", [(12608, 3072, 768), (8192, 3072, 768)])
@pytest.mark.parametrize("bn", [0, 1256])
def test_forward_gelu_with_stash_then_gelugrad(cuda_dev, M, N, K, bn):
    """FFN-up forward (bias + GELU, GELU' stashed) and the matching FFN-down dgrad (x stash) on the stash the kernel wrote."""
    from vilmedic_b200 import ops
    a, w = _rand((M, K), 3), _rand((N, K), 4, 0.05)
    bias = torch.randn(N, device="cuda") * 0.5
    pre_ref = a.float() @ w.float().t() + bias
    stash = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    h = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, aux_out=stash, force_bn=bn)
    torch.cuda.synchronize()
    # erf-GELU and its derivative in fp32 (HF ACT2FN["gelu"] = nn.functional.gelu, exact form)
    g_ref = torch.nn.functional.gelu(pre_ref)
    cdf = 0.5 * (1 + torch.erf(pre_ref * 0.7071067811865476))
    dg_ref = cdf + pre_ref * torch.exp(-0.5 * pre_ref * pre_ref) * 0.3989422804014327
    _check(h, g_ref, "gelu", 2e-3)
    _check(stash, dg_ref, "gelu' stash", 2e-3)
    # dgrad of the FFN-down Linear with the x GELU' epilogue: dpre = (dy W2) * stash  — [M, K2] x [K2, N] (B MN-major)
    K2 = 768
    dy, w2 = _rand((M, K2), 5), _rand((K2, N), 6, 0.05)
    up = dy.float() @ w2.float()
    dpre = ops.gemm(dy, w2, b_mn_major=True, act=ops.ACT_GELU_GRAD, aux_in=stash, force_bn=bn)
    torch.cuda.synchronize()
    _check(dpre, up * stash.float(), "x stash", 2e-5 * K2 ** 0.5)


@pytest.mark.parametrize("M,N,K", [(12608, 768, 768), (12608, 768, 3072), (8192, 768, 768), (8192, 768, 3072)])
@pytest.mark.parametrize("p", [0.0, 0.1])
@pytest.mark.parametrize("bn", [0, 1256])
def test_forward_dropout_residual(cuda_dev, M, N, K, p, bn):
    """out-projection / FFN-down: bias -> dropout -> + residual, same Philox stream as vlm_dropout_bf16 on the [M, N] tensor."""
    from vilmedic_b200 import ops
    a, w, res = _rand((M, K), 7), _rand((N, K), 8, 0.05), _rand((M, N), 9)
    bias = torch.randn(N, device="cuda")
    y = a.float() @ w.float().t() + bias
    if p > 0:
        keep = ops.dropout(torch.ones(M, N, device="cuda", dtype=torch.bfloat16), p, 77, 5) != 0      # the Philox keep mask
        frac = 1.0 - keep.float().mean().item()
        assert abs(frac - p) < 5e-3
        y = y * keep.float() * (1.0 / (1.0 - p))
    ref = y + res.float()
    out = ops.gemm(a, w, bias=bias, residual=res, p_drop=p, seed=77, offset=5, force_bn=bn)
    torch.cuda.synchronize()
    _check(out, ref, "dropout+residual", 2e-5 * K ** 0.5)
    # and the dgrad flavour (B MN-major) with a residual-gradient add
    w2 = _rand((K, N), 10, 0.05)
    dy = _rand((M, K), 11)
    ref2 = dy.float() @ w2.float() + res.float()
    out2 = ops.gemm(dy, w2, b_mn_major=True, residual=res, force_bn=bn)
    torch.cuda.synchronize()
    _check(out2, ref2, "dgrad+residual", 2e-5 * K ** 0.5)


@pytest.mark.parametrize("M,N,K", [(768, 768, 12608), (2304, 768, 12608), (3072, 768, 12608), (768, 3072, 8192),
                                   (30522, 768, 8192)])
def test_wgrad_accumulate_splitk(cuda_dev, M, N, K):
    """dW += alpha_t * dY^T X (both operands MN-major, fp32 C accumulated in place; split-K atomics for few-tile shapes)."""
    from vilmedic_b200 import ops
    dy, x = _rand((K, M), 12, 0.1), _rand((K, N), 13)
    acc = torch.full((M, N), 0.25, device="cuda")
    g = torch.tensor(0.5, device="cuda")
    ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out=acc, accumulate=True, alpha_t=g)
    torch.cuda.synchronize()
    ref = 0.25 + 0.5 * (dy.float().t() @ x.float())
    err = (acc - ref).abs().max().item()
    assert err <= 1e-5 * ref.abs().max().item() + 2e-6 * K ** 0.5, (M, N, K, err)


def test_lm_head_dgrad_ragged_k(cuda_dev):
    """dh = dlogits[:, :V] E (K = V = 30522 is not a multiple of the 64-wide k block; B MN-major)."""
    from vilmedic_b200 import ops
    M, V, D = 8192, 30522, 768
    dl = _rand((M, V), 14, 0.01)
    E = _rand((V, D), 15, 0.05)
    out = ops.gemm(dl, E, b_mn_major=True)
    torch.cuda.synchronize()
    ref = dl.float() @ E.float()
    _check(out, ref, "lm-head dgrad", 2e-5 * V ** 0.5 * 0.01)
