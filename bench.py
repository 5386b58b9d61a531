"""bench.py — image-report pairs/s for RRG ViT-B/16 -> BERT-base decoder TRAINING (fwd + bwd + grad all-reduce + AdamW)
on N B200s (BASELINE.json metric, configs[1]: B=64/GPU, 224^2 images, 128-token reports, bf16, V=30522).

  python bench.py --gpus N --steps K --warmup W                       (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                                 the reference's CPU PyTorch/HF path (oracle), bounded sample

One JSON line on rank 0.  `value`: inputs resident in HBM.  `e2e`: the same step driven through the plugin API
(model(**batch) with pinned HOST tensors as the reference's DataLoader hands them over, H2D copies + loss D2H inside the
timed region).  `roofline`: the tcgen05 GEMM family (dominant kernel), algorithmic FLOPs / CUDA-event time of its launches.
`cpu_baseline`: the oracle (HF modules composed as the reference composes them) on the host cores, bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "image-report pairs/sec (RRG ViT-B/16 -> BERT-base decoder train)"
UNIT = "pairs/s"
VOCAB = 30522
FLOP_PER_PAIR_TRAIN = 220.8e9  # 73.6 GFLOP forward (ViT 35.13 + decoder 38.48) x 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="rrg", choices=["rrg", "convirt", "mvqa"],
                    help="rrg = BASELINE configs[1] (the metric's config, default); convirt = configs[2]; mvqa = configs[3]")
    ap.add_argument("--no-decode", action="store_true", help="skip the secondary metric (configs[4] ensemble beam decode tokens/s)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the informational same-box GPU arm of the reference path")
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (BASELINE configs[1]: 64)")
    ap.add_argument("--seq-len", type=int, default=128)
    ap.add_argument("--dropout", type=float, default=0.1, help="decoder dropout (config/RRG/baseline-mimic.yml:14,19)")
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="resident-input leg only (for profiler runs)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    return ap.parse_args()


def model_cfgs(dropout):
    from vilmedic_b200 import synth
    dec = synth.bert_base_decoder(vocab=VOCAB, layers=12, dropout=dropout)
    cnn = dict(proto="VisualEncoder", backbone="vit", permute="no_permute", **synth.vit_b16())
    return dec, cnn


# ------------------------------------------------------------------------------------------------ CPU reference arm
def host_threads():
    """Use every host core this process may run on (torchrun pins OMP_NUM_THREADS=1 for its workers)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_reference_step_rate(batch, seq_len, steps, warmup, dropout):
    """pairs/s of the oracle training step (fp32, eager, AdamW) on the host cores."""
    import copy

    from vilmedic_b200 import synth
    from oracle.rrg import OracleRRG
    torch.manual_seed(0)
    dec, cnn = model_cfgs(dropout)
    model = OracleRRG(dec, cnn).train()
    opt = torch.optim.AdamW(model.parameters(), lr=5e-5)
    b = synth.rrg_batch(batch, seq_len, VOCAB)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = model(b["input_ids"], b["attention_mask"], b["images"])
        opt.zero_grad()
        out["loss"].backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return batch / mean, mean


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    steps = max(1, min(args.steps, 3))
    warmup = max(1, min(args.warmup, 1))
    v, mean = cpu_reference_step_rate(args.cpu_batch, args.seq_len, steps, warmup, args.dropout)
    sample = "oracle (HF ViTModel + BertGenerationDecoder as vilmedic composes them) fp32 eager + AdamW, B=%d of the B=%d step, T=%d, %d timed steps" % (
        args.cpu_batch, args.batch, args.seq_len, steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
        "data": "synthetic",
        "config": {"workload": "RRG train step, ViT-B/16 -> 12-layer BERT decoder, V=%d, T=%d (configs[1]); CPU sample B=%d" % (
            VOCAB, args.seq_len, args.cpu_batch)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))



# ------------------------------------------------------------------------------------------------ same-box GPU arm (informational)
def gpu_reference_step_rate(batch, seq_len, steps, warmup, dropout, dev):
    """pairs/s of the reference's own path on THIS GPU: the oracle composition (HF ViTModel + BertGenerationDecoder wired as
    vilmedic wires them) under torch bf16 autocast with SDPA attention and the fused AdamW — what a user of the reference gets on a
    B200 after switching AMP to bf16 (the reference itself trains fp16 + GradScaler, trainor.py:96).  cuBLAS / cuDNN / SDPA library
    kernels; none of this repo's kernels."""
    from vilmedic_b200 import synth
    from oracle.rrg import OracleRRG
    torch.manual_seed(0)
    dec, cnn = model_cfgs(dropout)
    model = OracleRRG(dec, cnn)
    model.enc.model.config._attn_implementation = "sdpa"
    model.dec.decoder.config._attn_implementation = "sdpa"
    model = model.to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=0.01, fused=True)
    b = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.rrg_batch(batch, seq_len, VOCAB).items()}

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(b["input_ids"], b["attention_mask"], b["images"])
        opt.zero_grad(set_to_none=True)
        out["loss"].float().backward()
        opt.step()
        return out["loss"]

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, opt
    torch.cuda.empty_cache()
    return batch / (ms / 1e3), ms, float(loss)


def run_reference_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    v, ms, loss = gpu_reference_step_rate(args.batch, args.seq_len, max(args.steps, 3), max(args.warmup, 3), args.dropout, dev)
    print(json.dumps({
        "impl": "reference-gpu", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": max(args.steps, 3),
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 autocast", "data": "synthetic",
        "config": {"workload": "RRG train step, ViT-B/16 -> 12-layer BERT decoder, V=%d, B=%d, T=%d (configs[1]); HF modules, SDPA, "
                               "torch bf16 autocast, fused AdamW — library kernels only, informational" % (VOCAB, args.batch, args.seq_len),
                   "loss_last": loss},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------ secondary metric: ensemble beam decode
def decode_metric(dev, batch=32, beams=4, n_models=2, max_len=128):
    """BASELINE configs[4]: tokens/s of the sum-of-logits ensemble beam search (two independently seeded RRG ViT-B/16 -> 12-layer
    decoders, B=32, beam 4, max_length 128, V=30522) through the public `generate` — image encoding by every model + the device-side
    search (one CUDA-graph replay per generated token).  EOS is made unreachable so that exactly B*(max_len-1) tokens are generated.
    HBM roofline: per step and model the decoder weights (275 MB bf16) + the K/V the step reads,
    2*12*(t + 197)*768*2 B per row at the mean t = (max_len-1)/2."""
    from vilmedic_b200 import synth
    from vilmedic_b200.models import RRG
    BOS, PAD, EOS = 0, 1, 2
    models = []
    for s in range(n_models):
        torch.manual_seed(100 + s)
        dec, cnn = model_cfgs(0.0)
        m = RRG(dec, cnn).to(dev).eval()
        with torch.no_grad():
            m.dec.decoder.lm_head.bias[EOS] = -1e4
        models.append(m)
    images = synth.rrg_batch(batch, 8, VOCAB)["images"].pin_memory()
    times = []
    out = None
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            img = images.to(dev, non_blocking=True)
            encs, masks = zip(*[m.encode(img) for m in models])
            out = models[0].dec.decoder.generate(input_ids=torch.full((batch, 1), BOS, dtype=torch.long, device=dev),
                                                 encoder_hidden_states=list(encs), encoder_attention_mask=list(masks),
                                                 ensemble=[m.dec.decoder for m in models], max_length=max_len, num_beams=beams,
                                                 bos_token_id=BOS, eos_token_id=EOS, pad_token_id=PAD)
            out = out.cpu()                                        # the result (token ids) is read back inside the timed region
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 1e3)
    assert out.shape[0] == batch and out.shape[1] == max_len, tuple(out.shape)
    best = min(times[1:])
    rows = batch * beams
    t_mean = (max_len - 1) / 2.0
    step_bytes = n_models * (275e6 + rows * 2 * 12 * (t_mean + 197) * 768 * 2)
    pk = peaks()
    ms_step = 1e3 * best / (max_len - 1)
    achieved = step_bytes / (ms_step / 1e3) / 1e9
    del models
    torch.cuda.empty_cache()
    from vilmedic_b200.blocks.huggingface.decoder.beam import DeviceSearch
    DeviceSearch._cache.clear()
    return {"metric": "generated tokens/s, RRG inference: %d-model ensemble, beam %d, batch %d, max_length %d (BASELINE configs[4])" % (
                n_models, beams, batch, max_len),
            "value": batch * (max_len - 1) / best, "unit": "tokens/s", "seconds": best, "ms_per_search_step": ms_step,
            "h2d_bytes": images.numel() * 4, "d2h_bytes": out.numel() * 8,
            "how": "public generate(): H2D of the images, ViT encode by every model, device-side search (one CUDA-graph replay per token, "
                   "ensemble members on parallel graph branches), D2H of the token ids — all inside the timed region; best of 2 after 1 warm-up",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                         "algorithmic_bytes_per_step": step_bytes}}

# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index):
        self.proc = None
        self.index = index
        self.path = "/tmp/vlm_clocks_%d_%d.csv" % (os.getpid(), index)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [l.strip().split(", ") for l in open(self.path) if l.strip()]
            sm = [float(r[0]) for r in rows if len(r) >= 7]
            mx = [float(r[1]) for r in rows if len(r) >= 7]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = set()
            for r in rows:
                if len(r) >= 7:
                    for n, val in zip(names, r[3:7]):
                        if val.strip().lower().startswith("active"):
                            reasons.add(n)
            if sm:
                out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
            os.remove(self.path)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist

    from vilmedic_b200 import synth
    from vilmedic_b200 import ops
    from vilmedic_b200.arena import get_arena
    from vilmedic_b200.models import RRG
    from vilmedic_b200.optim import FusedAdamW

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    t_start = time.time()

    def trace(msg):                      # progress markers on stderr (VLM_BENCH_TRACE=1): where a multi-rank run got to
        if os.environ.get("VLM_BENCH_TRACE"):
            sys.stderr.write("[bench rank %d +%.1fs] %s\n" % (rank, time.time() - t_start, msg))
            sys.stderr.flush()

    if os.environ.get("VLM_BENCH_TRACE"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("VLM_BENCH_TRACE_AFTER", "90")), repeat=False, file=sys.stderr)
    if world > 1:
        import datetime
        # a protocol bug must fail within minutes, not hang the box for the default 10-minute NCCL timeout
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    trace("process group up")
    torch.manual_seed(0)
    dec, cnn = model_cfgs(args.dropout)
    model = RRG(dec, cnn).cuda().train()
    arena = get_arena(model)
    opt = FusedAdamW(model, lr=5e-5, weight_decay=0.01)
    B, T = args.batch, args.seq_len
    host = synth.rrg_batch(B, T, VOCAB, seed=1234 + rank)
    host = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    devb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values() if isinstance(v, torch.Tensor))
    from vilmedic_b200.ddp import GradSync
    sync = GradSync(arena, optimizer=opt).attach()       # per-layer gradient buckets, launched from the backward pass (ddp.py)

    # DIAGNOSTIC ONLY (tools/jobs): VLM_BENCH_NO_EXCHANGE=1 times the N-rank step WITHOUT the gradient all-reduce, to split the
    # multi-GPU loss into "exchange" and "everything else"; the printed line is marked invalid.
    no_exchange = os.environ.get("VLM_BENCH_NO_EXCHANGE") == "1"

    def train_step(batch, read_loss, exchange=True):
        """exchange=False: rank-local step without the gradient all-reduce (instrumented passes that only one rank runs)."""
        exchange = exchange and not no_exchange
        # world > 1: every layer's backward announces its gradient span (nn.notify_grad_ready -> GradSync.on_ready) and the
        # all-reduce of that bucket runs on NCCL's stream / NVLink under the backward of the layers below it
        from vilmedic_b200 import nn as vnn
        hook = vnn.GRAD_READY_HOOK[0]
        if not exchange:
            vnn.GRAD_READY_HOOK[0] = None
        try:
            out = model(**batch)
            loss = out["loss"]
            loss.backward()
        finally:
            vnn.GRAD_READY_HOOK[0] = hook
        if exchange:
            sync.step(opt)              # gradient exchange + fused optimizer, bucket by bucket under the backward pass (ddp.py)
        else:
            opt.step(grad_scale=1.0)
        if read_loss:
            return loss.item()
        return loss

    graphed = None

    def run_step(batch, read_loss):
        if graphed is None:
            return train_step(batch, read_loss)
        # resident leg: static inputs already hold the batch; e2e leg: pinned host tensors are copied in (H2D) every step
        loss = graphed(None if batch is devb else batch)
        return loss.item() if read_loss else loss

    def timed(batch_fn, steps, read_loss):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(steps):
            last = run_step(batch_fn(), read_loss)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
            dist.barrier()
        return ms, last

    def host_batch():
        if graphed is not None:
            return host                    # GraphedTrainStep copies pinned host tensors into its static inputs
        return {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}

    # warm-up (also builds the arena / first-use attribute setup)
    for _ in range(max(args.warmup, 3)):
        l0 = ops.LAUNCHES[0]
        train_step(devb, False)
        launches = ops.LAUNCHES[0] - l0      # kernels of ours per step (the graph replays exactly these)
    torch.cuda.synchronize()
    trace("warm-up done (%d launches / step)" % launches)
    graph_note = "eager launches"
    if not args.no_graph:
        try:
            from vilmedic_b200.graph import GraphedTrainStep
            graphed = GraphedTrainStep(model, opt, devb, warmup=1,
                                       step_fn=lambda b: (ops.rng_advance(ops.RNG_COUNTER[0], 4096), train_step(b, False))[1])
            graph_note = "whole step replayed as one CUDA graph"
            for _ in range(2):
                graphed(devb)
            torch.cuda.synchronize()
            trace("graph captured and replayed")
        except Exception as e:  # pragma: no cover - reported, never silent
            graphed = None
            graph_note = "eager launches (graph capture failed: %s)" % (str(e).splitlines()[0][:160],)
            torch.cuda.synchronize()

    if world > 1:                         # all ranks replay a graph, or none does
        ok = torch.tensor([1 if graphed is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and graphed is not None:
            graphed = None
            graph_note = "eager launches (graph capture failed on another rank)"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, last = timed(lambda: devb, args.steps, False)
    trace("resident leg timed: %.2f ms / step" % (ms / args.steps))
    e2e_note = "inputs copied host -> device at the start of every step"
    if args.quick:
        ms_e2e, last_loss = ms, None
        args.no_roofline = args.no_cpu_baseline = True
    else:
        prefetch_ok = graphed is not None and os.environ.get("VLM_BENCH_PREFETCH", "1") == "1"
        if prefetch_ok:
            try:                      # input pipeline with one batch of look-ahead: step n+1's H2D copy runs under step n
                graphed.prefetch(host)
                graphed.replay_prefetched()
                torch.cuda.synchronize()
            except Exception as e:    # pragma: no cover - reported in the JSON line, never silent
                prefetch_ok = False
                e2e_note += " (prefetch path failed: %s)" % (str(e).splitlines()[0][:120],)
                torch.cuda.synchronize()
        # VLM_BENCH_SYNC_LOSS=1: read the loss with loss.item() right after every replay (the host then idles the GPU for its own
        # launch latency once per step: 22.85 vs 22.28 ms measured in round 2)
        async_loss = os.environ.get("VLM_BENCH_SYNC_LOSS") != "1"
        if prefetch_ok:
            def e2e_leg(steps):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                graphed.prefetch(host)                       # first batch: its copy is inside the timed region too
                last = None
                if async_loss:
                    for _ in range(steps):
                        # replay on the prefetched batch, start the next batch's H2D copy, copy this step's loss D2H (pinned, event);
                        # the value collected here is the PREVIOUS step's, so the host stays one step ahead of the GPU
                        prev = graphed.step_prefetched_async(host)
                        last = prev if prev is not None else last
                    last = graphed.drain()                   # the last step's loss: inside the timed region as well
                else:
                    for _ in range(steps):
                        loss = graphed.replay_prefetched()
                        graphed.prefetch(host)               # next step's inputs travel while this step computes
                        last = loss.item()                   # D2H read of the step's result
                e1.record()
                torch.cuda.synchronize()
                t_ms = e0.elapsed_time(e1)
                if world > 1:
                    tt = torch.tensor([t_ms], device=dev)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    t_ms = tt.item()
                    dist.barrier()
                return t_ms, last
            try:
                ms_e2e, last_loss = e2e_leg(args.steps)
                e2e_note = ("every step: pinned host -> device copy of the NEXT batch on a copy stream under the current step "
                            "(one batch of look-ahead), device-to-device hand-over, " +
                            ("D2H copy of the step's loss into pinned memory behind the step, collected by the host one step later "
                             "(GraphedTrainStep.step_prefetched_async)" if async_loss else "loss.item()"))
            except Exception as e:    # pragma: no cover - reported, never silent
                torch.cuda.synchronize()
                e2e_note += " (prefetch leg failed: %s)" % (str(e).splitlines()[0][:120],)
                ms_e2e, last_loss = timed(host_batch, args.steps, True)
        else:
            ms_e2e, last_loss = timed(host_batch, args.steps, True)
    clocks = sampler.stop() if rank == 0 else None
    trace("e2e leg timed")

    value = world * B * args.steps / (ms / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    roof = None
    if rank == 0 and not args.no_roofline:
        roof = gemm_roofline(lambda b, r: train_step(b, r, exchange=False), devb, ops)   # rank-local: no collectives
    trace("roofline pass done")
    if world > 1:
        dist.barrier()

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:      # reported at N=1 only (the contract)
        cores = host_threads()
        v, mean = cpu_reference_step_rate(args.cpu_batch, T, 2, 1, args.dropout)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle RRG train step (HF ViTModel + BertGenerationDecoder, fp32 eager, AdamW), B=%d, T=%d, 1 warm-up + 2 timed steps (%.1f s/step)" % (
                   args.cpu_batch, T, mean)}

    decode = gpu_base = None
    if rank == 0 and world == 1 and not args.quick:
        if not args.no_decode:
            try:
                decode = decode_metric(dev)
            except Exception as e:      # pragma: no cover - reported, never silent
                decode = {"error": str(e).splitlines()[0][:200]}
                torch.cuda.synchronize()
        if not args.no_gpu_baseline:
            try:
                v_g, ms_g, _ = gpu_reference_step_rate(B, T, 5, 3, args.dropout, dev)
                gpu_base = {"value": v_g, "unit": UNIT, "ms_per_step": ms_g, "kind": "reference path on this GPU (informational)",
                            "how": "oracle composition (HF ViTModel + BertGenerationDecoder), torch bf16 autocast, SDPA, fused AdamW, "
                                   "B=%d, T=%d, eager; library kernels only" % (B, T), "speedup_of_this_repo": value / v_g}
            except Exception as e:      # pragma: no cover
                gpu_base = {"error": str(e).splitlines()[0][:200]}
                torch.cuda.synchronize()

    # peer-memory transport: the device-side error flag of every rank (bounded waits, csrc/p2p.cu) must be clear
    px_err = False
    if world > 1 and sync.px is not None:
        t = torch.tensor([int(sync.px.err.item())], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        px_err = int(t.item()) != 0
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "RRG train step (fwd+bwd+grad all-reduce+AdamW): ViT-B/16 -> 12-layer BERT-base decoder, "
                                   "V=%d, B=%d/GPU, 224x224 images, T=%d, decoder dropout %.2f (BASELINE configs[1])" % (VOCAB, B, T, args.dropout),
                       "global_batch": world * B, "seq_len": T, "parallelism": "dp%d" % world,
                       "l2": "per-step working set (activations ~4 GB + 1.3 GB weights/grads) >> 126 MB L2; no explicit flush",
                       "launch": graph_note,
                       "grad_exchange": ("none (1 GPU)" if world == 1 else
                                         ("peer memory: every rank's optimizer kernel reads all ranks' bf16 gradient buckets over NVLink "
                                          "(csrc/p2p.cu), no collective in the step" if sync.transport == "p2p" else
                                          "NCCL all-reduce, %s payload, per-layer buckets under the backward pass" % sync.payload)),
                       "loss_last": float(last_loss) if last_loss is not None else None},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "how": e2e_note},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "decode": decode,
            "gpu_baseline": gpu_base,
            "mfu_vs_sustained_peak": value / world * FLOP_PER_PAIR_TRAIN / (peaks()["bf16_tflops_sustained"] * 1e12),
        }
        if no_exchange:
            line["invalid"] = "VLM_BENCH_NO_EXCHANGE=1: gradient all-reduce skipped (diagnostic run, not a bench value)"
        if px_err:
            line["invalid"] = "peer-memory exchange: a wait on a peer flag timed out on some rank (csrc/p2p.cu) — not a bench value"
        print(json.dumps(line))
    if world > 1:
        # destroy_process_group() blocks forever here (both ranks, observed on 2 x B200 with torch 2.11 / NCCL 2.28): the
        # captured CUDA graph still holds NCCL work.  Everything is flushed and every rank has passed a final barrier, so
        # leave without the communicator teardown.
        dist.barrier()
        torch.cuda.synchronize()
        trace("done")
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ------------------------------------------------------------------------------------------------ other BASELINE configs
OTHER = {
    # workload: (config file, metric, unit, train FLOP per sample, per-GPU batch rule)
    "convirt": ("config/SELFSUP/synthetic-convirt-resnet50.yml", "image-text pairs/sec (ConVIRT ResNet-50 + BERT-base train, BASELINE configs[2])",
                "pairs/s", 91.7e9),
    "mvqa": ("config/MVQA/synthetic-vit-b16.yml", "images/sec (MVQA ViT-B/16 + BERT encoder + classifier train, BASELINE configs[3])",
             "images/s", 210.8e9),
}


def run_other_workload(args):
    """configs[2] (ConVIRT, 64 pairs / GPU -> global 512 at 8 GPUs, rank-local negatives, weak scaling) and configs[3] (MVQA, global
    batch 256 split over the ranks, strong scaling) through the same harness: model built from the synthetic YAML config by
    executors.create_model (the reference's construction path), the optimizer the config names on the fused kernel, per-layer
    gradient buckets, whole step replayed as a CUDA graph when capture succeeds.  No roofline object: the headline kernel analysis
    belongs to the rrg workload."""
    import torch.distributed as dist

    from vilmedic_b200 import executors, ops
    from vilmedic_b200.arena import get_arena
    from vilmedic_b200.ddp import GradSync
    path, metric, unit, flop = OTHER[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    config = executors.load_config(os.path.join(ROOT, path))
    tcfg = executors.utils.get(config, "trainor")
    if args.workload == "convirt":
        per_gpu, scaling = 64, "weak"
    else:
        assert 256 % world == 0
        per_gpu, scaling = 256 // world, "strong"
    torch.manual_seed(0)
    dl = executors.SyntheticLoader(tcfg, per_gpu, n_batches=1, seed=1234 + rank)
    model = executors.create_model(tcfg, dl).train()
    opt = executors.create_optimizer(tcfg, None, model)
    arena = get_arena(model)
    sync = GradSync(arena, optimizer=opt).attach()
    host = next(iter(dl))
    host = {k: v for k, v in host.items() if v is not None}
    devb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values() if isinstance(v, torch.Tensor))

    def train_step(batch):
        out = model(**batch)
        out["loss"].backward()
        sync.step(opt)
        return out["loss"]

    # eager warm-up on the side stream the graph will be captured on (autograd ties a leaf's gradient accumulation to the stream of
    # its first use; these models have leaves that receive gradients through autograd — see graph.py)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(args.warmup, 3)):
            l0 = ops.LAUNCHES[0]
            train_step(devb)
            launches = ops.LAUNCHES[0] - l0
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graphed, note = None, "eager launches"
    if not args.no_graph:
        try:
            from vilmedic_b200.graph import GraphedTrainStep
            graphed = GraphedTrainStep(model, opt, devb, warmup=1, stream=side,
                                       step_fn=lambda b: (ops.rng_advance(ops.RNG_COUNTER[0], 4096), train_step(b))[1])
            note = "whole step replayed as one CUDA graph"
        except Exception as e:      # pragma: no cover - reported, never silent
            graphed = None
            note = "eager launches (graph capture failed: %s)" % (str(e).splitlines()[0][:160],)
            torch.cuda.synchronize()
            import gc
            gc.collect()
            torch.cuda.empty_cache()
    if world > 1:
        ok = torch.tensor([1 if graphed is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            graphed = None

    def timed(e2e):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(args.steps):
            if graphed is not None:
                loss = graphed(host if e2e else None)
            else:
                loss = train_step({k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in host.items()} if e2e else devb)
            last = loss.item() if e2e else loss
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, last

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, _ = timed(False)
    ms_e2e, last = timed(True)
    clocks = sampler.stop() if rank == 0 else None
    loss_kernels = None
    if rank == 0 and args.workload == "convirt":
        loss_kernels = contrastive_loss_timing(dev)
    if rank == 0:
        value = world * per_gpu * args.steps / (ms / 1e3)
        print(json.dumps({
            "loss_kernels": loss_kernels,
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "%s train step from %s (model built by executors.create_model), per-GPU batch %d, optimizer %s" % (
                args.workload, path, per_gpu, tcfg.optimizer), "global_batch": world * per_gpu, "parallelism": "dp%d" % world,
                "launch": note, "loss_last": float(last)},
            "e2e": {"value": world * per_gpu * args.steps / (ms_e2e / 1e3), "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "how": "pinned host batch copied to the device at the start of every step, loss.item()"},
            "gpu_launches": launches, "clocks": clocks, "roofline": None, "cpu_baseline": None,
            "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9,
            "mfu_vs_sustained_peak": value / world * flop / (peaks()["bf16_tflops_sustained"] * 1e12)}))
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)



def contrastive_loss_timing(dev, N=512, D=768, iters=20):
    """The contrastive loss on its own at the GLOBAL batch of BASELINE configs[2] (N = 512 pairs, D = 768): forward + backward of
    ConVIRTLoss (row-normalise + hi/lo split, tcgen05 similarity GEMM, symmetric LSE / CE kernels and their backward) captured into a
    CUDA graph and replayed; us per loss evaluation and GB/s on the algorithmic bytes of SURVEY.md §8d (2*N*D*4 read + 3*N*4 written,
    forward; the backward reads them again and writes 2*N*D*4)."""
    from vilmedic_b200.blocks.losses import ConVIRTLoss
    try:
        g = torch.Generator(device="cpu").manual_seed(0)
        l = torch.randn(N, D, generator=g).to(dev).requires_grad_(True)
        v = torch.randn(N, D, generator=g).to(dev).requires_grad_(True)
        crit = ConVIRTLoss(tau=0.1, lambda_=0.75)

        def once():
            l.grad = v.grad = None
            loss = crit(l, v)[0]
            loss.backward()
            return loss
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                once()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            loss = once()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        alg = 2 * (2 * N * D * 4) + 3 * N * 4 + 2 * N * D * 4
        return {"what": "ConVIRTLoss forward + backward, N=%d, D=%d, one CUDA-graph replay" % (N, D), "us": us,
                "algorithmic_bytes": alg, "gbs_on_algorithmic_bytes": alg / us / 1e3, "loss": float(loss.item())}
    except Exception as e:      # pragma: no cover - reported, never silent
        torch.cuda.synchronize()
        return {"error": str(e).splitlines()[0][:200]}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def gemm_roofline(train_step, batch, ops):
    """Dominant-kernel roofline: every tcgen05 GEMM launch of one training step, timed with CUDA events on the launching
    stream.  The step is first run once with recording on (exact arguments + operands of all its GEMM launches are kept);
    the recorded launches are then replayed in step order, back to back, each bracketed by an event pair.  (Timing them
    inside the eager step itself is not robust: the ~1.9 k stream entries of a step make the host the bottleneck on some
    boxes, and the event pairs then also measure launch gaps.)  The replay runs behind a short spin kernel, so the queue
    is never empty; operands are the real tensors of the step, > 126 MB apart from launch to launch (no L2 reuse beyond what
    the step itself has)."""
    torch.cuda.synchronize()
    ops.GEMM_TIMING = []
    train_step(batch, False)
    torch.cuda.synchronize()
    recs = ops.GEMM_TIMING
    ops.GEMM_TIMING = None
    for r in recs[:8]:                       # warm the replay path (ctypes thunks, event pool)
        r[5]()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in recs]
    torch.cuda._sleep(int(0.02 * 1.9e9))
    for r, (e0, e1) in zip(recs, evs):
        e0.record()
        r[5]()
        e1.record()
    torch.cuda.synchronize()
    tot_ms, tot_flop, tot_bytes = 0.0, 0.0, 0.0
    by_shape = {}
    for (M, N, K, nb, nbytes, _call, _keep), (e0, e1) in zip(recs, evs):
        ms = e0.elapsed_time(e1)
        fl = 2.0 * M * N * K * nb
        tot_ms += ms
        tot_flop += fl
        tot_bytes += nbytes
        k = (M, N, K, nb)
        a = by_shape.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ms
        a[2] += fl
    # The family's time back to back: the same launches, same order, same operands, captured into ONE CUDA graph (as the step itself
    # is replayed) and timed with a single event pair around each replay.  The per-launch event pairs above add ~2-3 us of record /
    # launch latency to every one of the ~400 launches (round 1: 17.7 ms against 16.3 ms of ncu kernel time); they are kept for the
    # per-shape table only.
    pair_ms, how = tot_ms, "all GEMM launches of one step recorded, then replayed in order under CUDA event pairs"
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for r in recs:
                r[5]()
        g.replay()
        torch.cuda.synchronize()
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        tot_ms = e0.elapsed_time(e1) / reps
        how = ("all GEMM launches of one step recorded with their operands, captured in step order into one CUDA graph, the replay "
               "timed with one CUDA event pair (mean of %d replays); per-shape figures from per-launch event pairs" % reps)
    except Exception as e:      # pragma: no cover - reported in the method string
        torch.cuda.synchronize()
        how += " (graph replay of the family failed: %s)" % (str(e).splitlines()[0][:120],)
    pk = peaks()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")     # ncu dram__bytes_{read,write}.sum over the same launches
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    achieved = tot_flop / (tot_ms / 1e3) / 1e12 if tot_ms > 0 else 0.0
    ranked = sorted(by_shape.items(), key=lambda kv: -kv[1][1])
    top = ranked[:6]
    if os.environ.get("VLM_BENCH_SHAPES"):      # full per-shape table (builder diagnostics, not part of the JSON line)
        with open(os.environ["VLM_BENCH_SHAPES"], "w") as f:
            for k, v in ranked:
                f.write("%-28s n=%3d  %8.3f ms  %7.1f us/launch  %7.1f TF/s\n" % (
                    "x".join(map(str, k)), v[0], v[1], v[1] * 1e3 / v[0], v[2] / (v[1] / 1e3) / 1e12))
    return {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (all launches of one training step)",
            "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
            "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step)",
            "traffic": traffic.get("dram_bytes_per_step"), "traffic_source": traffic.get("source"),
            "algorithmic_bytes": tot_bytes, "launches": len(recs),
            "method": how, "gemm_ms_per_step": tot_ms, "gemm_ms_event_pairs": pair_ms, "gemm_flop_per_step": tot_flop,
            "top_shapes": [{"MNKb": list(k), "n": v[0], "ms": round(v[1], 3), "tflops": round(v[2] / (v[1] / 1e3) / 1e12, 1)} for k, v in top]}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "reference-gpu":
        run_reference_gpu(a)
    elif a.workload != "rrg":
        run_other_workload(a)
    else:
        run_ours(a)
