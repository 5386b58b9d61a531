"""CUDA-graph capture of the whole training step (forward + hand-written backward + gradient exchange + fused AdamW).

The step issues ~1.1 k kernel launches through ctypes; replaying them as one graph removes the host from the loop
(SURVEY.md §7 "Host overhead", guide rule 9).  Everything in the step is capture-safe by construction: no host sync,
all scalars that change per step (AdamW step count, dropout stream position) live in device memory — the dropout
kernels add the device counter `ops.RNG_COUNTER` to their baked (seed, offset), and the graph advances it first.
"""
import torch

from . import ops


class GraphedTrainStep:
    def __init__(self, model, optimizer, example_batch, grad_sync=None, warmup=3, step_fn=None, stream=None):
        self.model = model
        self.opt = optimizer
        self.sync = grad_sync
        dev = next(model.parameters()).device
        self.static = {k: (v.to(dev).clone() if isinstance(v, torch.Tensor) else v) for k, v in example_batch.items()}
        self.counter = torch.zeros(1, device=dev, dtype=torch.int64)
        ops.RNG_COUNTER[0] = self.counter
        self._step_fn = step_fn or self._default_step
        # (high priority: kernels that the step forks onto side streams — ops.SideWork, the pipelined optimizer of ddp.GradSync — are
        #  low priority and only take the SMs the dependent chain of this stream leaves idle; captured kernel nodes keep it)
        s = stream if stream is not None else torch.cuda.Stream(priority=-1)      # pass the stream earlier eager steps ran on, if any
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step_fn(self.static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # capture on the SAME side stream the warm-up ran on: autograd ties every leaf's gradient accumulation to the stream on which
        # the leaf was first used, and waits for those streams at the end of backward — a leaf first used on a stream that is not
        # being captured makes the capture fail ("dependency created on uncaptured work in another stream")
        with torch.cuda.graph(self.graph, stream=s):
            self.loss = self._step_fn(self.static)
        torch.cuda.synchronize()

    def _default_step(self, batch):
        ops.rng_advance(self.counter, 4096)
        out = self.model(**batch)
        loss = out["loss"]
        loss.backward()
        if self.sync is not None:
            self.sync.step(self.opt)          # exchange + update, bucket by bucket under the backward pass when pipelined (ddp.py)
        else:
            self.opt.step()
        return loss

    # ---- input prefetch: the next step's batch travels host -> device on a copy stream while the current step computes ----
    def prefetch(self, batch):
        """Start copying `batch` (pinned host tensors) into device staging buffers on a side stream."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._staging = {k: torch.empty_like(v) for k, v in self.static.items() if isinstance(v, torch.Tensor)}
            self._ready = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed)      # the previous step has finished reading the staging buffers
            for k, v in batch.items():
                if isinstance(v, torch.Tensor):
                    self._staging[k].copy_(v, non_blocking=True)
            self._ready.record(self._copy_stream)

    def replay_prefetched(self):
        """Wait for the prefetched batch, move it into the graph's static inputs (device-to-device), replay."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready)
        for k, v in self._staging.items():
            self.static[k].copy_(v, non_blocking=True)
        self._consumed.record(cur)
        self.graph.replay()
        return self.loss

    # ---- loss read-back without stalling the GPU: the value of step n is copied device -> pinned host behind step n and
    #      collected by the host while step n + 1 is already running (the host stays one step ahead, the GPU never idles) ----
    def step_prefetched_async(self, next_batch=None):
        """replay_prefetched() + prefetch(next_batch) + an asynchronous D2H copy of this step's loss.
        Returns the loss VALUE (float) of the PREVIOUS call, None on the first; drain() returns the last one."""
        if not hasattr(self, "_loss_host"):
            self._loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._loss_ev = [torch.cuda.Event() for _ in range(2)]
            self._loss_pending = None
            self._loss_n = 0
        loss = self.replay_prefetched()
        if next_batch is not None:
            self.prefetch(next_batch)
        slot = self._loss_n & 1
        self._loss_n += 1
        self._loss_host[slot].copy_(loss.detach().reshape(1), non_blocking=True)
        self._loss_ev[slot].record(torch.cuda.current_stream())
        prev, self._loss_pending = self._loss_pending, slot
        if prev is None:
            return None
        self._loss_ev[prev].synchronize()
        return float(self._loss_host[prev][0])

    def drain(self):
        """Loss value of the last step_prefetched_async() call (waits for it)."""
        if getattr(self, "_loss_pending", None) is None:
            return None
        slot, self._loss_pending = self._loss_pending, None
        self._loss_ev[slot].synchronize()
        return float(self._loss_host[slot][0])

    def __call__(self, batch=None):
        """Copy `batch` (host or device tensors) into the static inputs, replay, return the (device) loss tensor."""
        if batch is not None:
            for k, v in batch.items():
                if isinstance(v, torch.Tensor):
                    self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.loss
