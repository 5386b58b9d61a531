#!/bin/bash
# Round-2 GPU job Q (1 GPU): ncu captures of the rewritten forward attention (streamed chunks) on the ViT and causal shapes.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2q_status.log
NCU="ncu --set full --clock-control none --import-source on"
VLM_ATTN_FWD_KEEP=0 timeout 300 $NCU -k regex:attn_fwd_tc -s 1 -c 1 -o $O/r2q_attn_fwd_vit -f python tools/attn_bench.py --only vit --iters 2 > $O/r2q_ncu1.log 2>&1; echo "fwd vit rc=$?" >> $O/r2q_status.log
timeout 300 $NCU -k regex:attn_fwd_tc -s 1 -c 1 -o $O/r2q_attn_fwd_self -f python tools/attn_bench.py --only self --iters 2 > $O/r2q_ncu2.log 2>&1; echo "fwd self rc=$?" >> $O/r2q_status.log
cat $O/r2q_status.log
