"""Data-parallel gradient exchange for the flat arena: one process per GPU, NCCL all-reduce (SUM) on
`flat_grad` spans, issued asynchronously per top-level block so the decoder span travels over NVLink while the ViT
backward is still computing; the 1/world factor is folded into the fused optimizer step (grad_scale).
No activation collectives: every pair is independent, contrastive negatives are rank-local as in the reference
(vilmedic/executors/trainor_accelerate.py:122,132)."""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, arena, group=None):
        self.arena = arena
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pending = []
        self._done = set()

    def launch_span(self, name):
        """Asynchronously all-reduce the gradient span of top-level child `name` (idempotent within a step)."""
        if self.world == 1 or name in self._done or name not in self.arena.child_spans:
            return
        lo, hi = self.arena.child_spans[name]
        self._done.add(name)
        self.pending.append(dist.all_reduce(self.arena.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """All-reduce whatever has not been sent yet, wait for everything; returns the grad scale for the optimizer."""
        if self.world > 1:
            a = self.arena
            todo = []
            covered = sorted(a.child_spans[n] for n in self._done)
            cur = 0
            for lo, hi in covered:
                if lo > cur:
                    todo.append((cur, lo))
                cur = max(cur, hi)
            if cur < a.numel:
                todo.append((cur, a.numel))
            for lo, hi in todo:
                self.pending.append(dist.all_reduce(a.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for w in self.pending:
                w.wait()
        self.pending = []
        self._done = set()
        return 1.0 / self.world
