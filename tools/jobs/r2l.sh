#!/bin/bash
# Round-2 GPU job L (1 GPU): ncu --set full captures of the 16-warp attention kernels and the LayerNorm backward.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2l_status.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:attn_fwd_tc -s 1 -c 1 -o $O/r2l_attn_fwd_self -f python tools/attn_bench.py --only self --iters 2 > $O/r2l_ncu1.log 2>&1; echo "fwd self rc=$?" >> $O/r2l_status.log
timeout 300 $NCU -k regex:attn_fwd_tc -s 1 -c 1 -o $O/r2l_attn_fwd_vit -f python tools/attn_bench.py --only vit --iters 2 > $O/r2l_ncu2.log 2>&1; echo "fwd vit rc=$?" >> $O/r2l_status.log
timeout 300 $NCU -k regex:attn_bwd_tc -s 1 -c 1 -o $O/r2l_attn_bwd_cross -f python tools/attn_bench.py --only cross --iters 2 > $O/r2l_ncu3.log 2>&1; echo "bwd cross rc=$?" >> $O/r2l_status.log
timeout 300 $NCU -k regex:layernorm_bwd -s 40 -c 1 -o $O/r2l_ln_bwd -f python bench.py --steps 1 --warmup 1 --no-graph --quick --no-cpu-baseline --no-decode --no-gpu-baseline --no-roofline > $O/r2l_ncu4.log 2>&1; echo "ln bwd rc=$?" >> $O/r2l_status.log
cat $O/r2l_status.log; ls -la $O/r2l_*
