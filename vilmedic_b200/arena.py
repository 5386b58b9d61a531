"""Flat parameter arena: one fp32 master buffer, one bf16 mirror and one fp32 gradient buffer per model.

HBM layout (B200, 180 GB): every nn.Parameter of the model is a *view* into `flat` (fp32, 128-byte aligned slots);
the kernels read weights from `flat_bf16` (refreshed by one cast kernel, or written directly by the fused AdamW step)
and accumulate weight gradients straight into `flat_grad` (p.grad is a view of it).  Parameters that must be
consumed as one GEMM operand (query/key/value weights -> fused [3D, D] projection) are allocated back to back so the
fused operand is just a wider view — state_dict names/shapes stay HF-compatible (SURVEY.md §5 checkpoint row).
The gradient all-reduce of the data-parallel step is a single NCCL call on `flat_grad` chunks.
"""
import torch
import torch.nn as nn

from . import ops

_ALIGN = 32  # fp32 elements = 128 bytes


class ParamArena:
    def __init__(self, root: nn.Module):
        self.root = root
        # allocation order: one contiguous span per top-level child (decoder / encoder / heads), inside it the fused
        # groups first, then the remaining parameters -> the gradient all-reduce can be issued span by span while the
        # backward of the next span is still running.
        order = []
        placed = set()
        self.child_spans = {}
        owners = []                     # (module, first group index, one-past-last group index): parameters owned by the subtree
        children = list(root.named_children()) or [("", root)]
        blocks = [(name, child) for name, child in children]
        direct = [p for p in root.parameters(recurse=False)]

        def place_subtree(mod):
            """Depth-first: a module's fused groups, then its own parameters, then its children — so that the parameters of one
            transformer layer occupy ONE contiguous span (the gradient exchange is bucketed per layer, ddp.py)."""
            start = len(order)
            fn = getattr(mod, "_fused_param_groups", None)
            if fn is not None:
                for g in fn():
                    g = list(g)
                    if any(id(p) in placed for p in g):
                        continue
                    for p in g:
                        placed.add(id(p))
                    order.append(g)
            for p in mod.parameters(recurse=False):
                if id(p) not in placed:
                    placed.add(id(p))
                    order.append([p])
            for ch in mod.children():
                place_subtree(ch)
            owners.append((mod, start, len(order)))

        for name, child in blocks:
            start = len(order)
            place_subtree(child)
            self.child_spans[name] = (start, len(order))
        for p in direct:
            if id(p) not in placed:
                placed.add(id(p))
                order.append([p])
        dev = next(root.parameters()).device
        self.device = dev
        offsets = {}
        total = 0
        for g in order:
            total = (total + _ALIGN - 1) // _ALIGN * _ALIGN
            for p in g:
                assert p.dtype == torch.float32, "arena parameters must be fp32 masters"
                offsets[id(p)] = total
                total += p.numel()
        total = (total + _ALIGN - 1) // _ALIGN * _ALIGN
        # element spans [lo, hi) of each top-level child in the flat buffers
        spans = {}
        for name, (g0, g1) in self.child_spans.items():
            if g1 > g0:
                lo = offsets[id(order[g0][0])]
                hi = offsets[id(order[g1][0])] if g1 < len(order) else total
                spans[name] = (lo, hi)
        self.child_spans = spans
        # element span [lo, hi) of every module that owns at least one parameter slot (layers -> per-layer gradient buckets)
        self.module_spans = {}
        for mod, g0, g1 in owners:
            if g1 > g0:
                lo = offsets[id(order[g0][0])]
                hi = offsets[id(order[g1][0])] if g1 < len(order) else total
                self.module_spans[id(mod)] = (lo, hi)
        self.numel = total
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_bf16 = torch.zeros(total, device=dev, dtype=torch.bfloat16)
        self.offsets = offsets
        self._params = [p for g in order for p in g]
        with torch.no_grad():
            for p in self._params:
                off = offsets[id(p)]
                view = self.flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self.bind_grads()
        self.mirror_clean = False
        self.mirror_owner = None  # an optimizer that keeps flat_bf16 in sync (FusedAdamW)

    # ------------------------------------------------------------------ validity
    def valid(self):
        p = self._params[0]
        return p.device == self.device and p.data_ptr() == self.flat.data_ptr() + 4 * self.offsets[id(p)]

    # ------------------------------------------------------------------ views
    def _span(self, params):
        first = params[0]
        off = self.offsets[id(first)]
        n = 0
        for p in params:
            assert self.offsets[id(p)] == off + n, "parameters are not adjacent in the arena"
            n += p.numel()
        return off, n

    def bf16(self, *params, shape=None):
        off, n = self._span(params)
        v = self.flat_bf16[off:off + n]
        return v.view(shape if shape is not None else params[0].shape)

    def fp32(self, *params, shape=None):
        off, n = self._span(params)
        v = self.flat[off:off + n]
        return v.view(shape if shape is not None else params[0].shape)

    def grad(self, *params, shape=None):
        off, n = self._span(params)
        v = self.flat_grad[off:off + n]
        return v.view(shape if shape is not None else params[0].shape)

    # ------------------------------------------------------------------ per-step maintenance
    def bind_grads(self):
        """Make p.grad a view of flat_grad (after optimizer.zero_grad(set_to_none=True) dropped it)."""
        rebound = False
        for p in self._params:
            if p.requires_grad and (p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * self.offsets[id(p)]):
                off = self.offsets[id(p)]
                p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)
                rebound = True
        return rebound

    def prepare_step(self):
        """Called at the start of a training forward: grads bound (zeroed if they had been dropped), mirror fresh."""
        if any(p.grad is None for p in self._params if p.requires_grad):
            self.flat_grad.zero_()
            self.bind_grads()
        self.refresh_mirror()

    def refresh_mirror(self, force=False):
        if force or self.mirror_owner is None or not self.mirror_clean:
            ops.cast_bf16(self.flat, self.flat_bf16)
            self.mirror_clean = self.mirror_owner is not None


def get_arena(module: nn.Module) -> ParamArena:
    """Arena of the outermost module that has been flattened; built lazily on first use."""
    arena = getattr(module, "_vlm_arena", None)
    if arena is not None and arena.valid():
        return arena
    arena = ParamArena(module)
    for m in module.modules():
        object.__setattr__(m, "_vlm_arena", arena)
    return arena
