#!/bin/bash
# Round-2 last 2-GPU job: the three data-parallel tests (NCCL bf16 / fp32 payload, peer memory over 3 steps) on the final tree.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2fin3_*
timeout 300 python -m pytest tests/test_ddp_gpu.py -m gpu -q > $O/r2fin3_ddp_tests.log 2>&1; echo "ddp tests rc=$?" >> $O/r2fin3_status.log
cat $O/r2fin3_status.log; tail -3 $O/r2fin3_ddp_tests.log | cut -c1-300
