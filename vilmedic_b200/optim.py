"""Fused AdamW over the flat parameter arena (SURVEY.md §8f rank 1 — vilmedic/executors/trainor.py:119-124:
unscale + clip_grad_norm_ + optimizer.step + zero_grad, optimizer picked by name in executors/utils.py:81-86).

One kernel pass updates p/m/v, writes the bf16 mirror the GEMMs read, applies the global-norm clip and zeroes the
gradient buffer; nothing syncs with the host (step count, grad-norm and lr scale live in device memory), so the whole
training step can be captured in a CUDA graph.  Under data parallelism `grad_scale` carries the 1/world_size of the
summed all-reduce.
"""
import torch

from . import ops
from .arena import get_arena


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=0.0):
        self.arena = get_arena(model)
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        super().__init__([p for p in model.parameters() if p.requires_grad], defaults)
        a = self.arena
        self.m = torch.zeros_like(a.flat)
        self.v = torch.zeros_like(a.flat)
        self.step_t = torch.zeros(1, device=a.device, dtype=torch.int32)
        self.gnorm_sq = torch.zeros(1, device=a.device, dtype=torch.float32)
        self.lr_scale = torch.ones(1, device=a.device, dtype=torch.float32)
        a.mirror_owner = self
        a.refresh_mirror(force=True)
        a.mirror_clean = True

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        g = self.param_groups[0]
        a = self.arena
        max_norm = g["max_grad_norm"] or 0.0
        if max_norm > 0:
            self.gnorm_sq.zero_()
            ops.sumsq(a.flat_grad, self.gnorm_sq)
        ops.adamw_step(a.flat, a.flat_grad, self.m, self.v, a.flat_bf16, lr=g["lr"], betas=g["betas"], eps=g["eps"],
                       weight_decay=g["weight_decay"], step_t=self.step_t, increment_step=True, lr_scale_t=self.lr_scale,
                       grad_scale=grad_scale, gnorm_sq_t=self.gnorm_sq if max_norm > 0 else None, max_norm=max_norm,
                       zero_grad=True)
        a.mirror_clean = True

    def zero_grad(self, set_to_none=False):
        # gradients are zeroed inside the fused step; keep p.grad bound to the flat buffer
        self.arena.bind_grads()
