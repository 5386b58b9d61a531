#!/bin/bash
# Round-2 GPU job P (1 GPU): forward attention — how many score chunks to keep in registers for the wide (Nk = 224) shapes.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2p_status.log
for k in 0 2 4 7; do
  VLM_ATTN_FWD_KEEP=$k timeout 300 python tools/attn_bench.py > $O/r2p_attn_keep$k.log 2>&1; echo "keep $k rc=$?" >> $O/r2p_status.log
done
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention or attn" > $O/r2p_attn_tests.log 2>&1; echo "attn tests rc=$?" >> $O/r2p_status.log
VLM_ATTN_FWD_KEEP=0 timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention or attn" > $O/r2p_attn_tests0.log 2>&1; echo "attn tests keep0 rc=$?" >> $O/r2p_status.log
cat $O/r2p_status.log; for k in 0 2 4 7; do echo "keep=$k"; cat $O/r2p_attn_keep$k.log | tail -3; done; tail -2 $O/r2p_attn_tests.log; tail -2 $O/r2p_attn_tests0.log
