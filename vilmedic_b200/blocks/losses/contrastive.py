"""ConVIRTLoss / InfoNCELoss (+ the global half of GLoRIALoss) on the sm_100a kernels.

Same constructor kwargs and return tuples as the reference:
  ConVIRTLoss(tau, lambda_)(linguistic, visual) -> (loss, loss_l, loss_v)     vilmedic/blocks/losses/selfsup/ConVIRTLoss.py:5-23
  InfoNCELoss(tau)(linguistic, visual)          -> (loss, loss_t, loss_i)     vilmedic/blocks/losses/selfsup/InfoNCELoss.py:5-19
    (tau is accepted and unused, exactly like the reference — SURVEY.md defects #8)
  GLoRIAGlobalLoss(temp3)(img, txt)             -> (loss0, loss1)             vilmedic/blocks/losses/selfsup/GLoRIALoss.py:54-75
Similarity matrix on tcgen05 with the bf16 hi/lo split (3-term product, fp32 accumulate); the N x N matrix is consumed by
fused row/column log-sum-exp kernels.  Negatives are rank-local, as in the reference.
"""
import torch
import torch.nn as nn

from ... import ops


class _SymNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, normalize, scale, w_row, w_col):
        a = a.float().contiguous()
        b = b.float().contiguous()
        N = a.shape[0]
        xa, _, ah, inv_a = ops.rownorm_split(a, normalize, want_b=False)
        _, xb, bh, inv_b = ops.rownorm_split(b, normalize, want_a=False)
        S = ops.gemm(xa, xb, out_dtype=torch.float32)                      # [N, N] = a_hat b_hat^T (hi*hi + hi*lo + lo*hi)
        lse_row, lse_col, loss_row, loss_col = ops.sym_lse(S, scale)
        ctx.saved = (a, b, ah, bh, inv_a, inv_b, S, lse_row, lse_col, normalize, scale, w_row, w_col)
        ctx.mark_non_differentiable(loss_row, loss_col)
        loss = ops.sum_scale(_axpby(loss_row, loss_col, w_row, w_col), 1.0)
        return loss, loss_row, loss_col

    @staticmethod
    def backward(ctx, g, _gr, _gc):
        a, b, ah, bh, inv_a, inv_b, S, lse_row, lse_col, normalize, scale, w_row, w_col = ctx.saved
        dS = ops.sym_lse_bwd(S, scale, lse_row, lse_col, w_row, w_col, g.contiguous().float())
        dah = ops.gemm(dS, bh, b_mn_major=True, out_dtype=torch.float32)                 # dS   b_hat
        dbh = ops.gemm(dS, ah, a_mn_major=True, b_mn_major=True, out_dtype=torch.float32)  # dS^T a_hat
        da = ops.rownorm_bwd(a, inv_a, dah, normalize)
        db = ops.rownorm_bwd(b, inv_b, dbh, normalize)
        return da, db, None, None, None, None


def _axpby(x, y, wx, wy):
    """wx * x + wy * y on small fp32 vectors (per-row losses); one fused GEMV-free kernel is not worth a launch: the
    vectors have N <= a few thousand elements, so this reuses the deterministic reduction kernel on a stacked view."""
    out = torch.empty(2 * x.numel(), device=x.device, dtype=torch.float32)
    out[:x.numel()] = x * wx
    out[x.numel():] = y * wy
    return out


class ConVIRTLoss(nn.Module):
    def __init__(self, tau, lambda_, **kwargs):
        super().__init__()
        self.tau = tau
        self.lambda_ = lambda_

    def forward(self, linguistic, visual):
        n = linguistic.shape[0]
        # rows of S = linguistic -> loss_l (denominator over visuals); columns -> loss_v
        loss, loss_l, loss_v = _SymNCEFn.apply(linguistic.cuda(), visual.cuda(), True, 1.0 / self.tau,
                                                (1.0 - self.lambda_) / n, self.lambda_ / n)
        return loss, loss_l, loss_v

    def __repr__(self):
        return "ConVIRTLoss(\n\t(cos_loss): CosineSimilarity()\n\t(tau): {}\n\t(lambda_): {}\n)".format(self.tau, self.lambda_)


class InfoNCELoss(nn.Module):
    def __init__(self, tau, **kwargs):
        super().__init__()
        self.tau = tau

    def forward(self, linguistic, visual):
        n = linguistic.shape[0]
        loss, loss_t, loss_i = _SymNCEFn.apply(linguistic.cuda(), visual.cuda(), False, 1.0, 0.5 / n, 0.5 / n)
        return loss, loss_t, loss_i

    def __repr__(self):
        return "InfoNCELoss(\n\t(tau): {}\n)".format(self.tau)


class GLoRIAGlobalLoss(nn.Module):
    """global_loss(cnn_code, rnn_code, temp3) of GLoRIALoss.py:54-75 -> (loss0, loss1)."""

    def __init__(self, temp3=10.0, **kwargs):
        super().__init__()
        self.temp3 = temp3

    def forward(self, cnn_code, rnn_code):
        n = cnn_code.shape[0]
        loss0, _, _ = _SymNCEFn.apply(cnn_code.cuda(), rnn_code.cuda(), True, self.temp3, 1.0 / n, 0.0)
        loss1, _, _ = _SymNCEFn.apply(cnn_code.cuda(), rnn_code.cuda(), True, self.temp3, 0.0, 1.0 / n)
        return loss0, loss1
