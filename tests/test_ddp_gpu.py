"""Data-parallel gradient exchange on REAL GPUs: launches tools/ddp_check.py under torchrun with 2 ranks — after one
fused optimizer step the two replicas must be bit-identical and their update must equal the full-batch single-GPU update within
bf16 gradient noise.  Skipped on single-GPU boxes (the gloo world-2 test in tests/test_cpu.py covers the host logic there)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("payload", ["bf16", "fp32", "p2p"])
def test_two_rank_step_equals_full_batch_step(payload):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(root, "tools", "ddp_check.py")]
    # bf16 = NCCL all-reduce of the bf16 payload, fp32 = the reference's DDP all-reduce
    env = dict(os.environ, VLM_DDP_PAYLOAD=payload, VLM_DDP_TRANSPORT="nccl")
    if payload == "p2p":                                        # peer-memory transport (csrc/p2p.cu): bf16 buckets read by the optimizer
        env.update(VLM_DDP_PAYLOAD="bf16", VLM_DDP_TRANSPORT="p2p", VLM_DDP_STEPS="3")      # kernel of every rank, 3 steps (epochs)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "replicas identical=True" in r.stdout
