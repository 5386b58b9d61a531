"""Plain-torch stand-ins for the CNN kernels of csrc/conv.cu (and the few other ops the ResNet runner calls).

They are the executable SPECIFICATION of those kernels: tests/test_cpu.py swaps them in to check the host logic of
vilmedic_b200/cnn.py (tape, residual wiring, weight packing, BatchNorm bookkeeping) against torchvision autograd on the CPU,
and tests/test_cnn_gpu.py checks each CUDA kernel against the same functions.  TEST INFRASTRUCTURE ONLY."""
import torch
import torch.nn.functional as F


DT = torch.bfloat16        # activation dtype of the stand-ins (float32 turns them into exact references of the host logic)


def bf(x):
    return x.to(DT)


def conv_out_size(H, k, stride, pad):
    return (H + 2 * pad - k) // stride + 1


def cast_bf16(src, dst=None):
    if dst is None:
        return bf(src)
    dst.copy_(src)
    return dst


def gemm(a, b, *, a_mn_major=False, b_mn_major=False, out=None, out_dtype=torch.bfloat16, residual=None, accumulate=False, **kw):
    assert not kw or all(v in (None, 0, 0.0, 1.0, False) for v in kw.values()), kw
    A = a.float().t() if a_mn_major else a.float()
    Bm = b.float() if b_mn_major else b.float().t()
    c = A @ Bm
    if residual is not None:
        c = c + residual.float()
    if out is not None:
        out.copy_((out.float() + c) if accumulate else c)
        return out
    return c.to(DT if out_dtype == torch.bfloat16 else out_dtype)


def conv_weight_pack(w, Kp):
    Cout, Cin, KH, KW = w.shape
    wm = torch.zeros(Cout, Kp, dtype=DT, device=w.device)
    wm[:, :KH * KW * Cin] = bf(w.permute(0, 2, 3, 1).reshape(Cout, -1))
    return wm


def conv_wgrad_unpack(dwm, gw):
    Cout, Cin, KH, KW = gw.shape
    gw += dwm[:, :KH * KW * Cin].reshape(Cout, KH, KW, Cin).permute(0, 3, 1, 2)


def im2col_nhwc(x, B, H, W, C, KH, KW, stride, pad):
    u = F.unfold(x.view(B, H, W, C).permute(0, 3, 1, 2).float(), (KH, KW), padding=pad, stride=stride)   # [B, C*KH*KW, L]
    L = u.shape[-1]
    return bf(u.view(B, C, KH * KW, L).permute(0, 3, 2, 1).reshape(B * L, KH * KW * C))


def im2col_nchw_f32(img, KH, KW, stride, pad, Kp):
    B, C, H, W = img.shape
    u = F.unfold(img, (KH, KW), padding=pad, stride=stride)
    L = u.shape[-1]
    col = torch.zeros(B * L, Kp, dtype=DT, device=img.device)
    col[:, :KH * KW * C] = bf(u.view(B, C, KH * KW, L).permute(0, 3, 2, 1).reshape(B * L, KH * KW * C))
    return col


def col2im_nhwc(dcol, B, H, W, C, KH, KW, stride, pad, add=None):
    L = dcol.shape[0] // B
    u = dcol.float().view(B, L, KH * KW, C).permute(0, 3, 2, 1).reshape(B, C * KH * KW, L)
    dx = F.fold(u, (H, W), (KH, KW), padding=pad, stride=stride).permute(0, 2, 3, 1).reshape(B * H * W, C)
    if add is not None:
        dx = dx + add.float()
    return bf(dx)


def bn_train_fwd(x, gamma, beta, running_mean, running_var, num_batches, eps, momentum, relu, res=None):
    M, C = x.shape
    xf = x.float()
    mean = xf.mean(0)
    var = (xf * xf).mean(0) - mean * mean
    var = var.clamp_min(0)
    rstd = torch.rsqrt(var + eps)
    scale = gamma * rstd
    y = xf * scale + (beta - mean * scale)
    if res is not None:
        y = y + res.float()
    if relu:
        y = y.clamp_min(0)
    running_mean.mul_(1 - momentum).add_(momentum * mean)
    running_var.mul_(1 - momentum).add_(momentum * var * (M / max(M - 1, 1)))
    num_batches += 1
    return bf(y), mean, rstd


def bn_eval_fwd(x, gamma, beta, running_mean, running_var, eps, relu, res=None):
    scale = gamma * torch.rsqrt(running_var + eps)
    y = x.float() * scale + (beta - running_mean * scale)
    if res is not None:
        y = y + res.float()
    return bf(y.clamp_min(0) if relu else y)


def bn_train_bwd(dy, y, x, mean, rstd, gamma, dgamma, dbeta, relu, want_dres):
    M, C = x.shape
    g = dy.float()
    if relu:
        g = torch.where(y.float() > 0, g, torch.zeros_like(g))
    xh = (x.float() - mean) * rstd
    sg, sgx = g.sum(0), (g * xh).sum(0)
    dgamma += sgx
    dbeta += sg
    dx = gamma * rstd * (g - sg / M - xh * sgx / M)
    return bf(dx), (bf(g) if want_dres else None)


def maxpool3x3s2_fwd(x, B, H, W, C):
    xin = x.view(B, H, W, C).permute(0, 3, 1, 2).float()
    y = F.max_pool2d(xin, 3, 2, 1)
    return bf(y.permute(0, 2, 3, 1).reshape(-1, C)), ("standin-idx", x)       # the stand-in "index" is the input itself


def maxpool3x3s2_bwd(dy, idx, B, H, W, C):
    x = idx[1].view(B, H, W, C).permute(0, 3, 1, 2).float().requires_grad_(True)
    with torch.enable_grad():
        y = F.max_pool2d(x, 3, 2, 1)
    Ho, Wo = y.shape[2], y.shape[3]
    y.backward(dy.float().view(B, Ho, Wo, C).permute(0, 3, 1, 2))
    return bf(x.grad.permute(0, 2, 3, 1).reshape(B * H * W, C))


def avgpool_fwd(x, B, HW, C):
    return bf(x.float().view(B, HW, C).mean(1))


def avgpool_bwd(dy, B, HW, C):
    return bf((dy.float() / HW).view(B, 1, C).expand(B, HW, C).reshape(B * HW, C))


def dropout(x, p, seed, offset, out=None):
    raise AssertionError("not used by the CNN host-logic test")


ALL = ["conv_out_size", "cast_bf16", "gemm", "conv_weight_pack", "conv_wgrad_unpack", "im2col_nhwc", "im2col_nchw_f32", "col2im_nhwc",
       "bn_train_fwd", "bn_eval_fwd", "bn_train_bwd", "maxpool3x3s2_fwd", "maxpool3x3s2_bwd", "avgpool_fwd", "avgpool_bwd"]


def install(monkeypatch, ops_module):
    import sys
    me = sys.modules[__name__]
    for n in ALL:
        monkeypatch.setattr(ops_module, n, getattr(me, n))
