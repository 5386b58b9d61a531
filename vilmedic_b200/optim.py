"""Fused optimizers over the flat parameter arena (SURVEY.md §8f rank 1 — vilmedic/executors/trainor.py:119-124:
unscale + clip_grad_norm_ + optimizer.step + zero_grad; optimizer picked by name in executors/utils.py:81-86).

One kernel pass per trainable span updates p/m/v, writes the bf16 mirror the GEMMs read, applies the global-norm clip
and zeroes the gradient buffer; nothing syncs with the host (step count, grad-norm, lr scale and the NaN/Inf skip
decision live in device memory), so the whole training step can be captured in a CUDA graph.  Under data parallelism
`grad_scale` carries the 1/world_size of the summed all-reduce.

* `FusedAdamW` / `FusedAdam` / `FusedRAdam` — torch.optim.{AdamW,Adam,RAdam} semantics (the three names the reference's
  configs and the bench use); `create_optimizer(name, model, **optim_params)` mirrors `getattr(torch.optim, name)`.
* Frozen parameters (`requires_grad=False`, e.g. `VisualEncoder(freeze=True)`) are never decayed or updated: the kernel
  only runs over the contiguous spans of trainable parameters; the gradient slots of frozen spans are just cleared.
* `step(loss=...)`: device-side replacement of the reference's `isnan(loss) or isinf(loss)` host check
  (trainor.py:109-112) and of GradScaler's found-inf skip — a non-finite loss or gradient norm leaves the weights, the
  moments and the step count untouched, zeroes the gradients and bumps `skipped_steps`.
* `state_dict()` carries the moments, the step count and the lr scale (reference checkpoints store
  `optimizer.state_dict()`, trainor.py:194-199, and reload it in create_optimizer, utils.py:90-92).
"""
import torch

from . import _lib, ops
from ._lib import c_float, c_int, c_ll, c_void_p, ptr, stream_ptr
from .arena import get_arena

KIND_ADAMW, KIND_ADAM, KIND_RADAM = 0, 1, 2
_ALIGN = 32


def trainable_spans(arena):
    """Contiguous [lo, hi) element spans of the arena that hold trainable parameters (slot-aligned, adjacent spans merged),
    and the complementary frozen spans."""
    marks = []
    for p in arena._params:
        lo = arena.offsets[id(p)]
        hi = (lo + p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        marks.append((lo, min(hi, arena.numel), bool(p.requires_grad)))
    marks.sort()
    train, frozen = [], []
    for lo, hi, t in marks:
        dst = train if t else frozen
        if dst and dst[-1][1] == lo:
            dst[-1][1] = hi
        else:
            dst.append([lo, hi])
    return [tuple(s) for s in train], [tuple(s) for s in frozen]


class FusedOptimizer(torch.optim.Optimizer):
    kind = KIND_ADAMW

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: %r" % (lr,))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: %r" % (eps,))
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError("Invalid beta parameters: %r" % (betas,))
        if not 0.0 <= weight_decay:
            raise ValueError("Invalid weight_decay value: %r" % (weight_decay,))
        self.arena = get_arena(model)
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        super().__init__([p for p in model.parameters() if p.requires_grad], defaults)
        a = self.arena
        self.m = torch.zeros_like(a.flat)
        self.v = torch.zeros_like(a.flat)
        self.step_t = torch.zeros(1, device=a.device, dtype=torch.int32)
        self.skipped_steps = torch.zeros(1, device=a.device, dtype=torch.int32)
        self.gnorm_sq = torch.zeros(1, device=a.device, dtype=torch.float32)
        self.lr_scale = torch.ones(1, device=a.device, dtype=torch.float32)
        self.spans, self.frozen_spans = trainable_spans(a)
        a.mirror_owner = self
        a.refresh_mirror(force=True)
        a.mirror_clean = True

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0, loss=None, grad16=None):
        """loss: optional 0-dim fp32 CUDA tensor — a NaN/Inf value skips the step on the device (no host sync).
        grad16: optional bf16 tensor [arena.numel] holding the (all-reduced) gradient values — the exchange payload of
        ddp.GradSync; the fp32 gradient buffer is then only zeroed."""
        a = self.arena
        self.begin_step(loss=loss, grad16=grad16)
        self.step_range(0, a.numel, grad_scale=grad_scale, grad16=grad16)

    @torch.no_grad()
    def begin_step(self, loss=None, grad16=None):
        """First half of step(): global gradient norm (when clipping / skipping is on) and the step counter.  The second half,
        step_range(), may then be issued span by span — ddp.GradSync does that from the backward pass, bucket by bucket."""
        g = self.param_groups[0]
        a = self.arena
        L = _lib.lib()
        if grad16 is not None and not (grad16.is_cuda and grad16.dtype == torch.bfloat16 and grad16.numel() == a.numel):
            raise ValueError("step(grad16=...) expects a bf16 CUDA tensor with one element per arena slot")
        max_norm = g["max_grad_norm"] or 0.0
        loss_p = None
        if loss is not None:
            if not (loss.is_cuda and loss.dtype == torch.float32 and loss.numel() == 1):
                raise ValueError("step(loss=...) expects a 1-element fp32 CUDA tensor")
            loss_p = ptr(loss.detach())
        # the gradient norm doubles as the found-inf detector (GradScaler semantics), so it is always computed when a
        # skip decision is wanted
        want_norm = max_norm > 0 or loss is not None
        if want_norm:
            self.gnorm_sq.zero_()
            for lo, hi in self.spans:
                ops.sumsq((grad16 if grad16 is not None else a.flat_grad)[lo:hi], self.gnorm_sq)
        self._gn_p = ptr(self.gnorm_sq) if want_norm else None
        self._loss_p = loss_p
        self._loss_ref = loss                  # keep the tensor alive until the span launches have been issued
        ops.check(L.vlm_optim_step_begin(ptr(self.step_t), self._gn_p, loss_p, ptr(self.skipped_steps), stream_ptr()), "vlm_optim_step_begin")

    def needs_global_norm(self):
        """True when step() has to see ALL gradients before it may touch any parameter (global-norm clipping)."""
        return (self.param_groups[0]["max_grad_norm"] or 0.0) > 0

    @torch.no_grad()
    def step_range(self, lo, hi, grad_scale=1.0, grad16=None, peer=None):
        """Update the trainable parameters inside arena elements [lo, hi) (begin_step() must have run for this step).
        peer=(p2p.PeerExchange, bucket index) with [lo, hi) = the whole bucket: the gradient values are the sum over all ranks' bf16 buckets, read through peer memory by
        the update kernel itself (csrc/p2p.cu)."""
        g = self.param_groups[0]
        a = self.arena
        L = _lib.lib()
        max_norm = g["max_grad_norm"] or 0.0
        for slo, shi in self.spans:
            x, y = max(lo, slo), min(hi, shi)
            if y <= x:
                continue
            if peer is not None:
                peer[0].optim_span(self, x, y, peer[1], grad_scale, bucket_range=(lo, hi))
                continue
            ops.check(L.vlm_optim_step(c_int(self.kind), ptr(a.flat[x:y]), ptr(a.flat_grad[x:y]), ptr(self.m[x:y]),
                                       ptr(self.v[x:y]), ptr(a.flat_bf16[x:y]), c_ll(y - x), c_float(g["lr"]),
                                       c_float(g["betas"][0]), c_float(g["betas"][1]), c_float(g["eps"]), c_float(g["weight_decay"]),
                                       ptr(self.step_t), ptr(self.lr_scale), c_float(grad_scale), self._gn_p, c_float(max_norm),
                                       self._loss_p, c_int(1), ptr(grad16[x:y]) if grad16 is not None else None, stream_ptr()),
                      "vlm_optim_step")
        for slo, shi in self.frozen_spans:      # backward kernels may have accumulated into frozen slots: clear, never apply
            x, y = max(lo, slo), min(hi, shi)
            if y > x:
                a.flat_grad[x:y].zero_()
        a.mirror_clean = True

    def zero_grad(self, set_to_none=False):
        # gradients are zeroed inside the fused step; keep p.grad bound to the flat buffer
        self.arena.bind_grads()

    # ---- checkpointing (vilmedic/executors/trainor.py:194-199 saves optimizer.state_dict(); utils.py:90-92 reloads it) ----
    def state_dict(self):
        sd = super().state_dict()
        sd["fused"] = {"kind": self.kind, "numel": self.arena.numel, "m": self.m.detach().cpu().clone(),
                       "v": self.v.detach().cpu().clone(), "step": self.step_t.detach().cpu().clone(),
                       "lr_scale": self.lr_scale.detach().cpu().clone(),
                       "skipped_steps": self.skipped_steps.detach().cpu().clone()}
        return sd

    def load_state_dict(self, state_dict):
        fused = state_dict.get("fused")
        if fused is None:
            raise KeyError("optimizer state has no 'fused' entry: it was not written by a vilmedic_b200 fused optimizer")
        if fused["numel"] != self.arena.numel or fused["kind"] != self.kind:
            raise ValueError("fused optimizer state does not match this model/optimizer (numel %d vs %d, kind %d vs %d)" % (
                fused["numel"], self.arena.numel, fused["kind"], self.kind))
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "fused"})
        self.m.copy_(fused["m"])
        self.v.copy_(fused["v"])
        self.step_t.copy_(fused["step"])
        self.lr_scale.copy_(fused["lr_scale"])
        self.skipped_steps.copy_(fused["skipped_steps"])


class FusedAdamW(FusedOptimizer):
    """torch.optim.AdamW (decoupled weight decay, default 1e-2)."""
    kind = KIND_ADAMW

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=0.0):
        super().__init__(model, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)


class FusedAdam(FusedOptimizer):
    """torch.optim.Adam (L2 weight decay folded into the gradient, default 0)."""
    kind = KIND_ADAM


class FusedRAdam(FusedOptimizer):
    """torch.optim.RAdam (decoupled_weight_decay=False) — the optimizer of config/RRG/biomed-roberta-baseline-mimic.yml:37-40."""
    kind = KIND_RADAM


_BY_NAME = {"AdamW": FusedAdamW, "Adam": FusedAdam, "RAdam": FusedRAdam}
_UNSUPPORTED_KW = ("amsgrad", "maximize", "foreach", "capturable", "differentiable", "fused", "decoupled_weight_decay")


def create_optimizer(name, model, state_dict=None, **optim_params):
    """`getattr(torch.optim, name)(model.parameters(), **optim_params)` of vilmedic/executors/utils.py:65-95 on the fused
    kernels.  Unknown names raise NotImplementedError like the reference; torch options that would change the arithmetic and
    have no kernel raise instead of being ignored."""
    if "lr" not in optim_params:
        raise ValueError("config.optim_params.lr is required")
    if name not in _BY_NAME:
        raise NotImplementedError(name)
    for k in _UNSUPPORTED_KW:
        if optim_params.get(k):
            raise NotImplementedError("%s(%s=%r) has no fused kernel" % (name, k, optim_params[k]))
    kw = {k: v for k, v in optim_params.items() if k not in _UNSUPPORTED_KW}
    if isinstance(kw.get("betas"), list):
        kw["betas"] = tuple(kw["betas"])
    opt = _BY_NAME[name](model, **kw)
    if state_dict is not None and "optimizer" in state_dict:
        opt.load_state_dict(state_dict["optimizer"])
    return opt
