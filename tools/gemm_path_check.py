"""Bitwise comparison of the staged TMA-store epilogue against the direct-store epilogue (VLM_GEMM_TMA_STORE=0 in a child
process) over small / ragged shapes and every epilogue mode.  Usage: python tools/gemm_path_check.py"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def run_all():
    from vilmedic_b200 import ops
    dev = torch.device("cuda:0")
    outs = {}
    g = torch.Generator(device="cpu").manual_seed(0)
    for M in (1, 3, 12, 33, 128, 130, 300):
        for N in (64, 304, 768, 2304):
            for K in (64, 768):
                a = torch.randn(M, K, generator=g).to(dev).to(torch.bfloat16)
                b = (torch.randn(N, K, generator=g) * 0.05).to(dev).to(torch.bfloat16)
                bias = torch.randn(N, generator=g).to(dev)
                res = torch.randn(M, N, generator=g).to(dev).to(torch.bfloat16)
                aux = torch.randn(M, N, generator=g).to(dev).to(torch.bfloat16)
                for mode in ("none", "bias", "gelu", "res", "resdrop", "gelugrad"):
                    kw = {}
                    if mode == "bias":
                        kw = dict(bias=bias)
                    elif mode == "gelu":
                        kw = dict(bias=bias, act=ops.ACT_GELU, aux_out=torch.zeros(M, N, device=dev, dtype=torch.bfloat16))
                    elif mode == "res":
                        kw = dict(bias=bias, residual=res)
                    elif mode == "resdrop":
                        kw = dict(bias=bias, residual=res, p_drop=0.1, seed=5, offset=3)
                    elif mode == "gelugrad":
                        kw = dict(act=ops.ACT_GELU_GRAD, aux_in=aux)
                    out = torch.full((M, N), 7.0, device=dev, dtype=torch.bfloat16)
                    if mode == "gelugrad":
                        bt = (torch.randn(K, N, generator=g) * 0.05).to(dev).to(torch.bfloat16)
                        ops.gemm(a, bt, b_mn_major=True, out=out, **kw)
                    else:
                        ops.gemm(a, b, out=out, **kw)
                    outs["%d_%d_%d_%s" % (M, N, K, mode)] = out.float().cpu()
                    if mode == "gelu":
                        outs["%d_%d_%d_gelu_aux" % (M, N, K)] = kw["aux_out"].float().cpu()
    torch.cuda.synchronize()
    return outs


if __name__ == "__main__":
    if len(sys.argv) > 1:
        torch.save(run_all(), sys.argv[1])
        sys.exit(0)
    env = dict(os.environ, VLM_GEMM_TMA_STORE="0")
    subprocess.check_call([sys.executable, __file__, "/tmp/gemm_direct.pt"], env=env)
    direct = torch.load("/tmp/gemm_direct.pt")
    staged = run_all()
    bad = 0
    for k in sorted(direct):
        if not torch.equal(direct[k], staged[k]):
            d = (direct[k] - staged[k]).abs()
            bad += 1
            if bad <= 20:
                idx = torch.nonzero(d > 0)
                print("MISMATCH %-28s max %.4g  n=%d first=%s" % (k, d.max().item(), idx.shape[0], idx[0].tolist()))
    print("gemm_path_check: %d / %d cases differ" % (bad, len(direct)))
    sys.exit(1 if bad else 0)
