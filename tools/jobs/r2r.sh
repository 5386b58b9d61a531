#!/bin/bash
# Round-2 GPU job R (1 GPU): forward attention with the nfull-dispatched straight-line softmax.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2r_status.log
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention or attn" > $O/r2r_attn_tests.log 2>&1; echo "attn tests rc=$?" >> $O/r2r_status.log
timeout 300 python tools/attn_bench.py > $O/r2r_attn_bench.log 2>&1; echo "attn bench rc=$?" >> $O/r2r_status.log
cat $O/r2r_status.log; tail -2 $O/r2r_attn_tests.log; cat $O/r2r_attn_bench.log
