"""Data-parallel gradient exchange + optimizer for the flat arena: one process per GPU, rank-local batches,
no activation collectives (every pair is independent, contrastive negatives are rank-local as in the reference).

Bucketing: every transformer layer's parameters occupy one contiguous span of the arena (arena.module_spans); the hand-written
backward of a layer announces "the gradients of this span are final" (nn.notify_grad_ready) and the exchange of that span starts
while the backward of the layers below is still computing.  Spans arrive in descending address order, adjacent ones are merged
until a bucket reaches `bucket_bytes`; only the last bucket (patch embedding + whatever nobody announced) is exposed.

Transport "p2p" (default, csrc/p2p.cu + p2p.PeerExchange): NO collective in the step.  Every rank casts a bucket to bf16 into a
buffer its peers have mapped through CUDA IPC and raises a READY flag in every rank's memory; the fused optimizer kernel of
every rank then reads the bucket of ALL ranks through NVLink peer loads, sums in fp32 in rank order (replicas stay bit-identical)
and applies Adam / AdamW / RAdam in the same pass — exchange and update are one kernel, launched bucket by bucket from the
backward pass as small background CTAs (128 threads, <= 64 registers) that fit on an SM next to the resident persistent GEMM CTA.
From 4 ranks on the exchange is two-shot: each rank first sums ITS slice of the bucket over all ranks into an fp32 buffer
(reduce-scatter through peer memory), the update kernels read every slice from its owner ((N-1)/N * 6 B per parameter over NVLink
instead of (N-1) * 2 B).  The epoch / READY / REDUCED / DONE flag protocol lives in device memory, so the whole step still replays
as ONE CUDA graph; every wait is bounded (90 s) and raises an error flag instead of hanging.  Needs `optimizer=` (the update is
part of the exchange) and no global-norm clipping; otherwise, or when peer memory is unavailable, the NCCL path below is used.

Transport "nccl" (VLM_DDP_TRANSPORT=nccl): async NCCL all-reduce (SUM) per bucket.  Payload bf16 by default — a cast kernel
writes the bucket into `grad16` right before its all-reduce and the fused optimizer reads the reduced values from there
(optim.FusedOptimizer.step(grad16=...)): 2 B per parameter, no widening pass.  `payload="fp32"` (VLM_DDP_PAYLOAD=fp32) keeps the
reference's fp32 DDP all-reduce bit for bit (vilmedic/executors/trainor_accelerate.py:122,132).  The 1/world factor is folded
into the optimizer kernel (grad_scale).  VLM_PIPELINE_OPTIMIZER=1 additionally issues the optimizer update of a bucket on a side
stream right behind that bucket's all-reduce (also on one GPU); measured gain on one B200: < 0.1 ms of 22.4 ms, off by default.

Measured (round 2, RRG B=64/GPU, ms per step; tools/jobs/r2m|r2t|r2u|r2x|r2y|r2z|r2z2.sh):
  2 GPUs: 22.41 on one GPU -> NCCL bf16 23.33, p2p one-shot 22.90 (two-shot 23.34);  NCCL history: 22.45 without any exchange,
          fp32 payload 23.87, bf16 23.69; fewer channels are worse (4: 26.3, 2: 33.8), the bucket size does not matter (8 / 32 /
          128 MB within 0.1 ms), 4 SMs left free for 4 high-priority channels 23.42 (VLM_SM_MARGIN, NCCL_MAX_NCHANNELS) — the NCCL
          kernels (32 CTAs) cannot share an SM with the persistent one-CTA-per-SM GEMM / attention kernels;
  8 GPUs: NCCL bf16 24.08, p2p one-shot 26.09 (7 x 446 MB of peer reads per GPU and step, latency-bound as background CTAs),
          p2p two-shot 24.12 -> 23.82 with four peer loads in flight per thread.
Round 1 sent two unbucketed fp32 spans and launched the encoder one after the backward had ended."""
import os

import torch
import torch.distributed as dist

from . import nn as _nn
from . import ops as _ops

_DONE_SLOT = 255        # p2p.DONE_SLOT


class GradSync:
    def __init__(self, arena, group=None, bucket_bytes=None, payload=None, optimizer=None, transport=None):
        if payload is None:
            payload = os.environ.get("VLM_DDP_PAYLOAD", "bf16")
        if payload not in ("bf16", "fp32"):
            raise ValueError("GradSync payload must be 'bf16' or 'fp32', got %r" % (payload,))
        self.payload = payload
        if bucket_bytes is None:        # VLM_DDP_BUCKET_MB: tuning knob for tools/jobs (default 32 MB)
            bucket_bytes = int(float(os.environ.get("VLM_DDP_BUCKET_MB", "32")) * (1 << 20))
        self.arena = arena
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.pending = []
        self.sent = []                  # [lo, hi) element ranges already handed to NCCL / the optimizer this step
        self.run = None                 # the current run of adjacent announced spans [lo, hi)
        self.launches = 0
        # bf16 exchange buffer (one slot per arena element); None when nothing is exchanged or the payload is fp32
        self.grad16 = None
        # pipelined optimizer: needs every decision of the update to be local to a span (no global gradient norm)
        self.opt = None
        pipe_ok = optimizer is not None and not optimizer.needs_global_norm() and arena.flat_grad.is_cuda
        # transport "p2p" (VLM_DDP_TRANSPORT): no collective at all — the optimizer kernel of every rank reads the bf16 gradient
        # buckets of all ranks through peer memory (csrc/p2p.cu, p2p.PeerExchange); needs the pipelined optimizer and the bf16 payload
        self.px = None
        if transport is None:
            transport = os.environ.get("VLM_DDP_TRANSPORT", "p2p")      # falls back to NCCL when peer memory is unavailable
        if transport not in ("nccl", "p2p"):
            raise ValueError("GradSync transport must be 'nccl' or 'p2p', got %r" % (transport,))
        if transport == "p2p" and self.world > 1 and pipe_ok and payload == "bf16":
            try:
                from .p2p import PeerExchange
                self.px = PeerExchange(arena.numel, arena.flat_grad.device, group)
                self.grad16 = self.px.grad16
            except Exception as e:      # no peer access on this box / IPC refused: keep the NCCL path, say so
                import sys
                sys.stderr.write("GradSync: peer-memory transport unavailable (%s); using NCCL\n" % (str(e).splitlines()[0][:200],))
                self.px = None
            # all ranks take the same path
            ok = torch.tensor([1 if self.px is not None else 0], device=arena.flat_grad.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.px = None
        self.transport = "p2p" if self.px is not None else "nccl"
        if self.px is None:
            self.grad16 = None
            if self.world > 1 and payload == "bf16":
                self.grad16 = torch.zeros(arena.numel, device=arena.flat_grad.device, dtype=torch.bfloat16)
        self._slot = 0
        if pipe_ok and (self.px is not None or os.environ.get("VLM_PIPELINE_OPTIMIZER", "0") != "0"):
            self.opt = optimizer
            self.side = torch.cuda.Stream(device=arena.flat_grad.device)
            self.pub = torch.cuda.Stream(device=arena.flat_grad.device) if self.px is not None else None
            self._begun = False

    @property
    def pipelined(self):
        return self.opt is not None

    # ---- wiring: the backward functions call nn.notify_grad_ready(module) -> on_ready
    def attach(self):
        _nn.GRAD_READY_HOOK[0] = self.on_ready if (self.world > 1 or self.pipelined) else None
        return self

    def detach(self):
        if _nn.GRAD_READY_HOOK[0] == self.on_ready:
            _nn.GRAD_READY_HOOK[0] = None

    def _exchange(self, lo, hi):
        """cast (bf16 payload) + async all-reduce of [lo, hi) issued from the CURRENT stream; returns the work handle."""
        if self.grad16 is not None:
            buf = self.grad16[lo:hi]
            self._cast(self.arena.flat_grad[lo:hi], buf)
        else:
            buf = self.arena.flat_grad[lo:hi]
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _launch(self, lo, hi, tail=False):
        if hi <= lo:
            return
        self.sent.append((lo, hi))
        self.launches += 1
        if not self.pipelined:
            if self.world > 1:
                self.pending.append(self._exchange(lo, hi))
            return
        if self.px is not None:
            self._launch_p2p(lo, hi, tail)
            return
        # pipelined: the exchange is issued from the main stream (NCCL's stream waits for the kernels that produced these
        # gradients), the update of the bucket runs on the side stream behind it — neither blocks the backward pass
        work = self._exchange(lo, hi) if self.world > 1 else None
        if work is None:
            self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            if work is not None:
                work.wait()                             # side stream waits for NCCL's stream; the host does not
            if not self._begun:
                self.opt.begin_step()
                self._begun = True
            if tail:        # issued from finish(): the backward pass is over, nothing to hide under — full-size CTAs
                self.opt.step_range(lo, hi, grad_scale=1.0 / self.world, grad16=self.grad16)
            else:           # small CTAs that fit next to the resident persistent GEMM / attention CTAs (ops.background)
                with _ops.background():
                    self.opt.step_range(lo, hi, grad_scale=1.0 / self.world, grad16=self.grad16)

    def _launch_p2p(self, lo, hi, tail):
        """Peer-memory transport.  main stream: (first bucket of the step: advance the epoch, wait until the peers are done with my
        buffers of the previous step).  publish stream: cast the bucket into my bf16 buffer, raise READY[bucket] in every rank's
        flag block — short kernels that never wait, so a bucket is visible to the peers as soon as its layer's backward is done.
        side stream: (two-shot: sum my slice of the bucket over all ranks, raise REDUCED[bucket]) the fused exchange + optimizer
        kernel of the bucket; both poll the flags themselves.  Everything but the tail of the step runs as background CTAs."""
        px = self.px
        b = self._slot
        self._slot += 1
        if 2 * b + 1 >= _DONE_SLOT:
            raise RuntimeError("GradSync(p2p): more than 127 gradient buckets in one step")
        cur = torch.cuda.current_stream()
        if b == 0:
            px.begin_step()

        def body():
            self.pub.wait_stream(cur)
            with torch.cuda.stream(self.pub):
                self._cast(self.arena.flat_grad[lo:hi], self.grad16[lo:hi])
                px.signal(2 * b)
            self.side.wait_stream(self.pub)
            with torch.cuda.stream(self.side):
                if not self._begun:
                    self.opt.begin_step()
                    self._begun = True
                if px.two_shot:
                    px.reduce_slice(lo, hi, b)
                    px.signal(2 * b + 1)
                self.opt.step_range(lo, hi, grad_scale=1.0 / self.world, peer=(px, b))

        if tail:            # issued from finish(): the backward pass is over, nothing to hide under — full-size CTAs
            body()
        else:               # small CTAs that fit next to the resident persistent GEMM / attention CTAs (ops.background)
            with _ops.background():
                body()

    @staticmethod
    def _cast(src, dst):
        if src.is_cuda:
            from . import ops
            ops.cast_bf16(src, dst)
        else:                       # gloo tests of the host logic
            dst.copy_(src)

    def on_ready(self, module):
        span = self.arena.module_spans.get(id(module))
        if span is None or not (self.world > 1 or self.pipelined):
            return
        lo, hi = span
        if self.run is not None and hi == self.run[0]:
            self.run = (lo, self.run[1])                      # adjacent below the current run: merge
        elif self.run is not None and lo == self.run[1]:
            self.run = (self.run[0], hi)
        else:
            if self.run is not None:
                self._launch(*self.run)
            self.run = (lo, hi)
        if self.run[1] - self.run[0] >= self.bucket_elems:
            self._launch(*self.run)
            self.run = None

    def launch_span(self, name):
        """Asynchronously all-reduce the whole gradient span of top-level child `name` (kept for callers that bucket by tower)."""
        if not (self.world > 1 or self.pipelined) or name not in self.arena.child_spans:
            return
        lo, hi = self.arena.child_spans[name]
        for a, b in self._uncovered(lo, hi):
            self._launch(a, b)

    def _uncovered(self, lo, hi):
        out, cur = [], lo
        for a, b in sorted(self.sent):
            if b <= cur or a >= hi:
                continue
            if a > cur:
                out.append((cur, min(a, hi)))
            cur = max(cur, b)
        if cur < hi:
            out.append((cur, hi))
        return out

    def finish(self):
        """All-reduce whatever has not been sent yet, wait for everything; returns the grad scale for the optimizer.
        With the bf16 payload the reduced gradients are in `self.grad16` (pass it on: opt.step(grad_scale=..., grad16=sync.grad16));
        `flat_grad` keeps the rank-local fp32 values until the optimizer zeroes it.  In pipelined mode this also issues the update
        of the leftover ranges and joins the side stream: the optimizer step is complete when it returns (use step())."""
        if self.world > 1 or self.pipelined:
            if self.run is not None:
                self._launch(*self.run, tail=True)
                self.run = None
            for lo, hi in self._uncovered(0, self.arena.numel):
                self._launch(lo, hi, tail=True)
            for w in self.pending:
                w.wait()
            if self.pipelined:
                if self.px is not None and self._slot > 0:
                    with torch.cuda.stream(self.side):
                        self.px.signal(_DONE_SLOT)      # I have read everything I need from the peers' buffers of this step
                    torch.cuda.current_stream().wait_stream(self.pub)
                torch.cuda.current_stream().wait_stream(self.side)
                self._begun = False
                self._slot = 0
        self.pending = []
        self.sent = []
        self.run = None
        return 1.0 / self.world

    def step(self, optimizer, loss=None):
        """finish() + the optimizer step, whichever way it is scheduled — the call a training loop makes after backward()."""
        if self.pipelined:
            if optimizer is not self.opt:
                raise ValueError("GradSync.step: this GradSync pipelines a different optimizer")
            if loss is not None:
                raise NotImplementedError("device-side loss skip needs the whole-arena step: build GradSync without optimizer=")
            self.finish()
            return
        scale = self.finish()
        optimizer.step(grad_scale=scale, loss=loss, grad16=self.grad16)
