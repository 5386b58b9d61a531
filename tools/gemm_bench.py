"""Micro-benchmark of the tcgen05 GEMM on the shapes of the RRG training step: every tile configuration x operand
layout, CUDA-event timed back-to-back launches (GPU-bound).  Usage: python tools/gemm_bench.py [--json out.json] [--only "<substring of shape name>"] [--cfg <force_bn>] [--epi none|bias|gelu|res|gelugrad|acc]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vilmedic_b200 import ops  # noqa: E402

SHAPES = [  # (name, M, N, K, layout, epilogue)
    ("vit ffn-up fwd", 12608, 3072, 768, "nt", "gelu"),
    ("vit ffn-down fwd", 12608, 768, 3072, "nt", "res"),
    ("vit qkv fwd", 12608, 2304, 768, "nt", "bias"),
    ("vit out fwd", 12608, 768, 768, "nt", "res"),
    ("dec ffn-up fwd", 8192, 3072, 768, "nt", "gelu"),
    ("dec out fwd", 8192, 768, 768, "nt", "res"),
    ("vit ffn-down dgrad", 12608, 3072, 768, "nn", "gelugrad"),
    ("vit ffn-up dgrad", 12608, 768, 3072, "nn", "none"),
    ("vit ffn wgrad", 3072, 768, 12608, "tn", "acc"),
    ("vit out wgrad", 768, 768, 12608, "tn", "acc"),
    ("dec qkv wgrad", 2304, 768, 8192, "tn", "acc"),
    ("lm head fwd", 8192, 30522, 768, "nt", "bias"),
    ("lm head wgrad", 30522, 768, 8192, "tn", "acc"),
]
CONFIGS = [0, 128, 192, 256, 1128, 1256]


def run(name, M, N, K, layout, epi, cfg, dev, iters=10):
    Np = (N + 7) // 8 * 8
    if layout == "nt":
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        b = torch.randn(N, K, device=dev).to(torch.bfloat16)
        kw = {}
    elif layout == "nn":      # dgrad: dY [M, K'] x W [K', N] (b MN-major)
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        b = torch.randn(K, Np, device=dev).to(torch.bfloat16)[:, :N]
        kw = dict(b_mn_major=True)
    else:                      # wgrad: dY [K', M] , X [K', N]
        Mp = (M + 7) // 8 * 8
        a = torch.randn(K, Mp, device=dev).to(torch.bfloat16)[:, :M]
        b = torch.randn(K, Np, device=dev).to(torch.bfloat16)[:, :N]
        kw = dict(a_mn_major=True, b_mn_major=True)
    out_dtype = torch.float32 if epi == "acc" else torch.bfloat16
    out = torch.zeros(M, Np, device=dev, dtype=out_dtype)[:, :N]
    bias = torch.randn(Np, device=dev)
    if epi == "gelu":
        kw.update(bias=bias, act=ops.ACT_GELU, aux_out=torch.empty(M, Np, device=dev, dtype=torch.bfloat16)[:, :N])
    elif epi == "res":
        kw.update(bias=bias, residual=torch.randn(M, Np, device=dev).to(torch.bfloat16)[:, :N])
    elif epi == "bias":
        kw.update(bias=bias)
    elif epi == "gelugrad":
        kw.update(act=ops.ACT_GELU_GRAD, aux_in=torch.randn(M, Np, device=dev).to(torch.bfloat16)[:, :N])
    elif epi == "acc":
        kw.update(accumulate=True)
    if cfg >= 1000 and epi == "acc" and False:
        return None
    try:
        for _ in range(3):
            ops.gemm(a, b, out=out, force_bn=cfg, **kw)
        torch.cuda.synchronize()
        torch.cuda._sleep(int(2e7))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.gemm(a, b, out=out, force_bn=cfg, **kw)
        e1.record()
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        return {"err": str(e)[:80]}
    us = e0.elapsed_time(e1) * 1e3 / iters
    return {"us": round(us, 1), "tflops": round(2.0 * M * N * K / us / 1e6, 1)}


def main():
    dev = torch.device("cuda:0")
    rows = []
    print("%-22s %-18s " % ("shape", "MxNxK") + " ".join("%12s" % ("cfg%d" % c) for c in CONFIGS))
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    cfgs = [int(sys.argv[sys.argv.index("--cfg") + 1])] if "--cfg" in sys.argv else CONFIGS
    epi_override = sys.argv[sys.argv.index("--epi") + 1] if "--epi" in sys.argv else None
    for (name, M, N, K, layout, epi) in SHAPES:
        if only and only not in name:
            continue
        if epi_override:
            epi = epi_override
        res = [run(name, M, N, K, layout, epi, c, dev) if c in cfgs else None for c in CONFIGS]
        rows.append({"name": name, "M": M, "N": N, "K": K, "layout": layout, "epi": epi, "results": dict(zip(map(str, CONFIGS), res))})
        cells = []
        for r in res:
            cells.append("%12s" % ("-" if r is None else ("ERR" if "err" in r else "%5.0fus %4.0fT" % (r["us"], r["tflops"]))))
        print("%-22s %-18s " % (name, "%dx%dx%d" % (M, N, K)) + " ".join(cells))
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
