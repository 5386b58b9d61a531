"""Micro-benchmark of the HBM-bound "tail" kernels of the RRG training step at the step's own shapes: LayerNorm forward /
backward (old warp-per-row kernels vs the round-2 multi-warp-per-row ones), bias-gradient column sums (foreground /
background launch shape), the attention delta pre-pass and the shifted softmax-CE.

Every case rotates over enough buffer sets to exceed the 126 MB L2, the launches of one rotation are captured into a CUDA graph
(the Python / ctypes launch path is slower than these 5-20 us kernels) and the graph replay is timed with CUDA events.
Prints us per launch and GB/s on the ALGORITHMIC bytes (each tensor read or written once).

    python tools/tail_bench.py [--iters N] [--only ln_fwd|ln_bwd|colsum|delta|ce]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vilmedic_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")


def graph_time(fns, iters):
    """fns: list of closures (one per buffer set).  Returns us per launch."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * len(fns))


def nsets(bytes_per_set):
    return max(3, int(400e6 // bytes_per_set) + 1)


def report(name, us, nbytes):
    print("%-44s %8.2f us  %7.0f GB/s" % (name, us, nbytes / us / 1e3), flush=True)


def bench_ln_fwd(iters):
    for M in (12608, 8192):
        D = 768
        n = nsets(M * D * 4)
        xs = [torch.randn(M, D, device=DEV).to(torch.bfloat16) for _ in range(n)]
        g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
        for v2 in ("0", "1"):
            os.environ["VLM_LN_V2"] = v2
            us = graph_time([(lambda x=x: ops.layernorm_fwd(x, g, b, 1e-12)) for x in xs], iters)
            report("ln_fwd  M=%d v2=%s" % (M, v2), us, M * D * 4 + M * 8)


def bench_ln_bwd(iters):
    D = 768
    for M, drop, res in ((12608, None, True), (8192, (0.1, 1, 1), False), (8192, None, False)):
        n = nsets(M * D * 2 * (4 + (drop is not None)))
        xs = [torch.randn(M, D, device=DEV).to(torch.bfloat16) for _ in range(n)]
        dys = [torch.randn(M, D, device=DEV).to(torch.bfloat16) for _ in range(n)]
        rs = [torch.randn(M, D, device=DEV).to(torch.bfloat16) for _ in range(n)] if res else [None] * n
        mean, rstd = torch.randn(M, device=DEV), torch.rand(M, device=DEV) + 0.5
        g = torch.randn(D, device=DEV)
        dg, db, cs = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
        nb = M * D * 2 * (3 + int(res) + int(drop is not None)) + M * 8
        for v2 in ("0", "1"):
            os.environ["VLM_LN_V2"] = v2
            us = graph_time([(lambda x=x, dy=dy, r=r: ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, dres=r, drop=drop, colsum=cs))
                             for x, dy, r in zip(xs, dys, rs)], iters)
            report("ln_bwd  M=%d drop=%d dres=%d colsum v2=%s" % (M, drop is not None, res, v2), us, nb)


def bench_colsum(iters):
    for M, N in ((12608, 2304), (12608, 3072), (8192, 2304), (8192, 768), (8192, 3072), (8192, 30528)):
        n = nsets(M * N * 2)
        xs = [torch.randn(M, N, device=DEV).to(torch.bfloat16) for _ in range(n)]
        out = torch.zeros(N, device=DEV)
        us = graph_time([(lambda x=x: ops.colsum(x, out)) for x in xs], iters)
        report("colsum  %dx%d foreground" % (M, N), us, M * N * 2)

        def bg(x):
            with ops.background():
                ops.colsum(x, out)
        us = graph_time([(lambda x=x: bg(x)) for x in xs], iters)
        report("colsum  %dx%d background CTAs" % (M, N), us, M * N * 2)


def bench_delta(iters):
    B, H, DH = 64, 12, 64
    for T in (197, 128):
        n = nsets(B * T * H * DH * 4)
        os_ = [torch.randn(B, T, H * DH, device=DEV).to(torch.bfloat16) for _ in range(n)]
        dos = [torch.randn(B, T, H * DH, device=DEV).to(torch.bfloat16) for _ in range(n)]
        if not hasattr(ops, "attention_delta"):
            print("delta: ops.attention_delta not exposed")
            return
        us = graph_time([(lambda o=o, d=d: ops.attention_delta(o, d, H, DH)) for o, d in zip(os_, dos)], iters)
        report("attn_delta T=%d" % T, us, B * T * H * DH * 4 + B * H * T * 4)


def bench_ce(iters):
    M, V, Vp, T = 8192, 30522, 30528, 128
    n = 3
    bufs = [torch.randn(M, Vp, device=DEV).to(torch.bfloat16) for _ in range(n)]
    ids = torch.randint(0, V, (M // T, T), device=DEV)
    us = graph_time([(lambda b=b: ops.softmax_ce(b, ids, V, shift_T=T, grad_scale=1.0 / M, dlogits=b)) for b in bufs], iters)
    report("softmax_ce 8192x30522 (in-place dlogits)", us, M * V * 4)


def main():
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 10
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    for name, fn in (("ln_fwd", bench_ln_fwd), ("ln_bwd", bench_ln_bwd), ("colsum", bench_colsum), ("delta", bench_delta), ("ce", bench_ce)):
        if only and only != name:
            continue
        try:
            fn(iters)
        except Exception as e:  # keep going: one broken case must not hide the others
            print("%s: FAILED %r" % (name, e), flush=True)


if __name__ == "__main__":
    main()
