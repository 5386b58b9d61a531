"""Micro-benchmark of the attention kernels on the three shapes of the RRG training step (B=64, H=12, dh=64):
ViT self (197 x 197), decoder causal self (128 x 128, dropout 0.1), decoder cross (128 x 197, dropout 0.1).
CUDA-event timed back-to-back launches.  Usage: python tools/attn_bench.py [--iters N] [--only vit|self|cross]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vilmedic_b200 import ops  # noqa: E402

B, H, DH = 64, 12, 64
D = H * DH
SHAPES = [  # name, Tq, Sk, causal, p_drop, packed-qkv?
    ("vit", 197, 197, False, 0.0),
    ("self", 128, 128, True, 0.1),
    ("cross", 128, 197, False, 0.1),
]


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(2e7))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 20
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    dev = torch.device("cuda:0")
    for name, Tq, Sk, causal, p in SHAPES:
        if only and only != name:
            continue
        if Tq == Sk:
            qkv = (torch.randn(B, Tq, 3 * D, device=dev) * 0.5).to(torch.bfloat16)
            q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
            dqkv = torch.empty_like(qkv)
            dq, dk, dv = dqkv[:, :, :D], dqkv[:, :, D:2 * D], dqkv[:, :, 2 * D:]
        else:
            q = (torch.randn(B, Tq, D, device=dev) * 0.5).to(torch.bfloat16)
            kv = (torch.randn(B, Sk, 2 * D, device=dev) * 0.5).to(torch.bfloat16)
            k, v = kv[:, :, :D], kv[:, :, D:]
            dq = torch.empty_like(q)
            dkv = torch.empty_like(kv)
            dk, dv = dkv[:, :, :D], dkv[:, :, D:]
        kmask = torch.ones(B, Sk, device=dev, dtype=torch.uint8) if name != "vit" else None
        do = torch.randn(B, Tq, D, device=dev).to(torch.bfloat16)
        o, lse = ops.attention_fwd(q, k, v, H, DH, kmask=kmask, causal=causal, p_drop=p, seed=1, offset=1)
        fwd = timed(lambda: ops.attention_fwd(q, k, v, H, DH, kmask=kmask, causal=causal, p_drop=p, seed=1, offset=1), iters)
        bwd = timed(lambda: ops.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, H, DH, kmask=kmask, causal=causal, p_drop=p,
                                              seed=1, offset=1), iters)
        fl = 4.0 * B * H * Tq * Sk * DH
        print("%-6s Tq=%d Sk=%d causal=%d p=%.1f  fwd %7.1f us (%5.0f TF/s)   bwd(+delta) %7.1f us (%5.0f TF/s)" % (
            name, Tq, Sk, causal, p, fwd, fl / fwd / 1e6, bwd, 2.5 * fl / bwd / 1e6))


if __name__ == "__main__":
    main()
