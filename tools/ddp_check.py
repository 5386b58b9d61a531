"""2-rank data-parallel parity check on real GPUs: each rank steps RRG on its half of a batch with the
NCCL span all-reduce (vilmedic_b200.ddp.GradSync); rank 0 also steps an identically initialised replica on the WHOLE
batch.  After one AdamW step the parameters of the 2-rank run must match (a) each other bit-for-bit and (b) the
full-batch replica within bf16 gradient noise.  Dropout 0 so that both runs see the same function.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from vilmedic_b200 import synth
    from vilmedic_b200.arena import get_arena
    from vilmedic_b200.ddp import GradSync
    from vilmedic_b200.models import RRG
    from vilmedic_b200.optim import FusedAdamW

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    V, T, b = 1000, 32, 4
    dec = synth.bert_base_decoder(V, layers=2, dropout=0.0)
    cnn = dict(proto="VisualEncoder", backbone="vit", permute="no_permute", **synth.vit_b16())
    cnn["num_hidden_layers"] = 2

    def make():
        torch.manual_seed(0)
        m = RRG(dict(dec), dict(cnn)).cuda().train()
        return m, get_arena(m), FusedAdamW(m, lr=1e-3, weight_decay=0.01)

    full = synth.rrg_batch(world * b, T, V, seed=99)
    shard = {k: (v[rank * b:(rank + 1) * b].to(dev) if isinstance(v, torch.Tensor) else v) for k, v in full.items()}

    model, arena, opt = make()
    # per-layer buckets launched from the backward pass; VLM_DDP_PIPELINE=1: optimizer update pipelined behind each bucket
    # VLM_DDP_TRANSPORT=p2p: no NCCL in the step — the optimizer kernel reads all ranks' gradient buckets through peer memory
    pipe = os.environ.get("VLM_DDP_PIPELINE") == "1" or os.environ.get("VLM_DDP_TRANSPORT") == "p2p"
    nsteps = int(os.environ.get("VLM_DDP_STEPS", "1"))
    sync = GradSync(arena, bucket_bytes=8 << 20, optimizer=opt if pipe else None).attach()
    if os.environ.get("VLM_DDP_TRANSPORT") == "p2p":
        assert sync.transport == "p2p", "peer-memory transport was requested but is not active"
    for _ in range(nsteps):
        sync.launches = 0
        out = model(**shard)
        out["loss"].backward()
        early = sync.launches
        sync.step(opt)
    sync.detach()
    torch.cuda.synchronize()
    if sync.px is not None:
        sync.px.check()
    assert early >= 2, "no gradient bucket was launched during the backward pass"

    # (a) replicas identical after the step
    mine = arena.flat.clone()
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    same = bool((mine == ref0).all().item())
    flags = [None] * world
    dist.all_gather_object(flags, same)
    # (b) rank 0: full-batch replica
    ok_b, rel = True, 0.0
    if rank == 0:
        m2, a2, o2 = make()
        fb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in full.items()}
        for _ in range(nsteps):
            m2(**fb)["loss"].backward()
            o2.step(grad_scale=1.0)
        torch.cuda.synchronize()
        torch.manual_seed(0)
        init = get_arena(RRG(dict(dec), dict(cnn)).cuda()).flat
        d_ddp, d_full = mine - init, a2.flat - init
        rel = ((d_ddp - d_full).norm() / d_full.norm()).item()
        ok_b = rel < 5e-2
        print("ddp_check[%s, %d step(s)]: replicas identical=%s  update rel. diff vs full-batch step=%.3e  (|update|=%.3e)" % (
            sync.transport, nsteps, all(flags), rel, d_full.norm().item()), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not (all(flags) and ok_b):
        sys.exit(1)


if __name__ == "__main__":
    main()
