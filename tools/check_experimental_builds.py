"""Compile (only) the translation units touched by the remaining build-time option (the scalar GELU epilogue kept for comparison
with the packed fp32x2 one) into a scratch directory, so that it stays compilable without touching the in-tree library.

  python tools/check_experimental_builds.py          (CPU box is enough: nvcc cross-compiles sm_100a)
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vilmedic_b200 import build as b  # noqa: E402

CASES = [
    ("scalar GELU epilogue", ["-DVLM_GELU_F32X2=0"], ["gemm_tcgen05_bn192.cu"]),
]


def main():
    base = [f for f in b.NVCC_FLAGS if not f.startswith("-DVLM_")]
    ok = True
    with tempfile.TemporaryDirectory() as tmp:
        for name, flags, srcs in CASES:
            for src in srcs:
                out = os.path.join(tmp, src[:-3] + ".o")
                r = subprocess.run([b.NVCC] + base + flags + ["-c", os.path.join(b.CSRC, src), "-o", out], capture_output=True, text=True)
                status = "ok" if r.returncode == 0 else "FAILED"
                print("%-52s %-28s %s" % (name, src, status))
                if r.returncode != 0:
                    ok = False
                    sys.stderr.write(r.stderr[-2000:])
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
