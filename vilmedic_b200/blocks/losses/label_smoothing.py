"""LabelSmoothingCrossEntropy — mirror of vilmedic/blocks/losses/mvqa/LabelSmoothingCrossEntropyLoss.py:32-48 on the
fused softmax-CE kernel (loss and dlogits in one pass)."""
import torch
import torch.nn as nn

from ... import ops


class _LSCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, smoothing, reduction):
        x = output.float().contiguous()
        B, C = x.shape
        ld = (C + 3) // 4 * 4
        buf = x if ld == C else torch.nn.functional.pad(x, (0, ld - C))
        dl = torch.empty_like(buf)
        scale = 1.0 / B if reduction == "mean" else 1.0
        rows, _ = ops.softmax_ce(buf[:, :C] if ld != C else buf, target.long().contiguous(), C, smoothing=smoothing,
                                 grad_scale=scale, dlogits=dl)
        ctx.saved = (dl, C)
        if reduction == "none":
            return rows
        return ops.sum_scale(rows, scale)

    @staticmethod
    def backward(ctx, g):
        dl, C = ctx.saved
        return dl[:, :C] * g, None, None, None


class LabelSmoothingCrossEntropy(nn.Module):
    def __init__(self, smoothing=0.1, reduction="mean", **kwargs):
        super().__init__()
        self.smoothing = smoothing
        self.reduction = reduction

    def forward(self, output, target):
        if self.reduction not in ("mean", "sum"):
            raise NotImplementedError("reduction=%r" % self.reduction)
        return _LSCEFn.apply(output.cuda(), target.cuda(), self.smoothing, self.reduction)
